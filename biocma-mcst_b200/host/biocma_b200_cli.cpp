// Command-line driver with the reference's flags (apps/cli/src/cli_parser.cpp:106-152): every flag
// takes a value, parsed pairwise (:40-70).
//   -np particles  -d final time [s]  -dt step [s] (<= 0: auto for multi-compartment cases)
//   -mn model (fixed_length | monod | simple_acetate; anything else: alert + default model, global_initaliser.cpp:261-271)
//   -f case: directory with an rcmtool `cma_case` (as far as apps/api/tests/data/0d shows the format) | directory of raw
//      flat arrays written by biocma_mcst_b200.synth.write_case | ring:<n> | rotate:<maps>:<n>:<seconds per map> | 0d
//   -fi initial liquid concentrations: raw f64 file, n_compartments x n_species values, species fastest (the reference
//      reads an HDF5 initialiser, scalar_factory.cpp:166-223; no HDF5 library in this image)
//   -er result stem  -nex number of exports  -serde checkpoint to resume the MC unit from (cli_parser.cpp:147-151)
//   -nt/-force/-r accepted (ignored here).  `<stem>_serde_0.raw` is written at the end (serde.cpp:64-70).
#include <cstring>
#include <iostream>
#include <memory>

#include "bmc_host.hpp"

namespace {
struct UserControlParameters {
  std::string results_file_name = "result", model_name = "monod", cma_case_path = "0d";
  uint64_t number_particle = 0, number_exported_result = 2;
  double delta_time = 0., final_time = 0.;
  double feed_flow = 0., feed_concentration = 0.;
  uint64_t seed = 2024;
  bool load_serde = false; std::string serde_file;
  std::string initialiser_path;
};
int model_id(const std::string& n) {
  if (n == "fixed_length") return BMC_MODEL_FIXED_LENGTH;
  if (n == "monod") return BMC_MODEL_MONOD;
  if (n == "simple_acetate") return BMC_MODEL_SIMPLE_ACETATE;
  // "Model not found, using Default model instead" (global_initaliser.cpp:261-271; DefaultModel is the fixed-length model)
  std::cerr << "biocma_b200: model '" << n << "' not found, using the default model (fixed_length) instead\n";
  return BMC_MODEL_FIXED_LENGTH;
}
void write_raw(const std::string& path, const void* p, std::size_t bytes) {
  std::ofstream f(path, std::ios::binary);
  f.write(static_cast<const char*>(p), static_cast<std::streamsize>(bytes));
}
}  // namespace

int main(int argc, char** argv) {
  UserControlParameters uc;
  try {
    for (int i = 1; i + 1 < argc; i += 2) {
      const std::string k = argv[i] + (argv[i][0] == '-' ? 1 : 0), v = argv[i + 1];
      if (k == "np") uc.number_particle = std::stoull(v);
      else if (k == "d") uc.final_time = std::stod(v);
      else if (k == "dt") uc.delta_time = std::stod(v);
      else if (k == "mn") uc.model_name = v;
      else if (k == "f") uc.cma_case_path = v;
      else if (k == "er") uc.results_file_name = v;
      else if (k == "nex") uc.number_exported_result = std::stoull(v);
      else if (k == "feed") uc.feed_flow = std::stod(v);      // extension: constant chemostat feed [m3/s]
      else if (k == "feedc") uc.feed_concentration = std::stod(v);
      else if (k == "seed") uc.seed = std::stoull(v);
      else if (k == "serde") { uc.load_serde = true; uc.serde_file = v; }
      else if (k == "fi") uc.initialiser_path = v;
      else if (k == "nt" || k == "force" || k == "r") {}
      else throw std::invalid_argument("bad argument -" + k);
    }
    // sanitise_check_cli (cli_parser.cpp:267-289)
    if (uc.delta_time < 0) throw std::invalid_argument("Wrongtime step (d_t<0)");
    if (uc.number_particle == 0) throw std::invalid_argument("Missing number of particles");
    if (uc.final_time <= 0) throw std::invalid_argument("Final time must be positive");

    CmaUtils::FlowMap fm;
    std::unique_ptr<CmaUtils::Transitioner> transitioner;
    if (uc.cma_case_path == "0d") fm = CmaUtils::FlowMap::zero_d();
    else if (uc.cma_case_path.rfind("ring:", 0) == 0) fm = CmaUtils::FlowMap::ring(std::stoull(uc.cma_case_path.substr(5)), 0.02, 0.1);
    else if (uc.cma_case_path.rfind("rotate:", 0) == 0) {  // rotate:<maps>:<n>:<t_per_flow_map>: ring maps of increasing exchange rate
      std::size_t maps = 0, n = 0; double t_per = 0.;
      if (std::sscanf(uc.cma_case_path.c_str(), "rotate:%zu:%zu:%lf", &maps, &n, &t_per) != 3 || maps == 0 || n < 2) throw std::invalid_argument("bad rotate: case");
      std::vector<CmaUtils::FlowMap> ms;
      for (std::size_t k = 0; k < maps; ++k) ms.push_back(CmaUtils::FlowMap::ring(n, 0.02, 0.05 * static_cast<double>(k + 1)));
      fm = ms[0];
      transitioner = std::make_unique<CmaUtils::Transitioner>(std::move(ms), t_per);
    }
    else if (std::ifstream(uc.cma_case_path + "/cma_case").good()) fm = CmaUtils::FlowMap::load_cma_case(uc.cma_case_path);
    else fm = CmaUtils::FlowMap::load(uc.cma_case_path);
    if (uc.delta_time <= 0) uc.delta_time = fm.n > 1 ? CmaUtils::get_time_step(transitioner ? transitioner->advance(0., 0.) : fm) : 1e-2;  // global_initaliser.cpp:551-559

    const int model = model_id(uc.model_name);
    const std::size_t n_species = model == BMC_MODEL_SIMPLE_ACETATE ? 2 : 1;
    auto unit = std::make_unique<MC::MonteCarloUnit>(model, n_species, fm.n, uc.seed);
    Core::SimulationParameters params;
    params.d_t = uc.delta_time; params.final_time = uc.final_time; params.number_particle = uc.number_particle;
    params.number_exported_result = uc.number_exported_result;
    const double m_tot = unit->init(params.number_particle, params.uniform_mc_init, params.biomass_initial_concentration, fm.total_volume());
    std::vector<double> c0(n_species * fm.n, 1.0);  // uniform 1.0 without an initialiser file (global_initaliser.cpp:42-49)
    if (!uc.initialiser_path.empty()) {
      c0 = CmaUtils::FlowMap::read_raw<double>(uc.initialiser_path);
      if (c0.size() != n_species * fm.n) throw std::invalid_argument("-fi: expected n_compartments x n_species doubles");
    }
    Simulation::SimulationUnit simulation(std::move(unit), fm, n_species, c0);
    if (uc.load_serde) simulation.load_serde(uc.serde_file);  // particles, tallies, step counter and concentrations of the saved run
    if (uc.feed_flow > 0) simulation.add_feed(Simulation::Feed::FeedFactory::constant(uc.feed_flow, uc.feed_concentration, 0, 0));

    const Core::Records rec = Core::main_loop(params, simulation, transitioner.get());
    const auto c = simulation.mc_unit->counters();
    // particle balance (post_process.cpp:92-117)
    uint64_t total = 0;
    for (std::size_t k = rec.number_particle.size() - fm.n; k < rec.number_particle.size(); ++k) total += rec.number_particle[k];
    const bool balance = total == c.total_new - c.total_out + uc.number_particle;
    const std::string stem = uc.results_file_name;
    write_raw(stem + "_time.raw", rec.time.data(), rec.time.size() * 8);
    write_raw(stem + "_concentration_liquid.raw", rec.concentration_liquid.data(), rec.concentration_liquid.size() * 8);
    write_raw(stem + "_number_particle.raw", rec.number_particle.data(), rec.number_particle.size() * 8);
    write_raw(stem + "_tallies.raw", rec.tallies.data(), rec.tallies.size() * 8);
    simulation.save_serde(stem + "_serde_0.raw");
    std::printf("{\"n_compartments\": %zu, \"n_species\": %zu, \"d_t\": %.17g, \"n_records\": %zu, \"initial_mass\": %.17g, "
                "\"n_particles\": %llu, \"new\": %llu, \"out\": %llu, \"compactions\": %llu, \"steps\": %llu, \"balance_ok\": %s}\n",
                fm.n, n_species, uc.delta_time, rec.time.size(), m_tot, (unsigned long long)total, (unsigned long long)c.total_new,
                (unsigned long long)c.total_out, (unsigned long long)c.n_compactions, (unsigned long long)c.step, balance ? "true" : "false");
    return balance ? 0 : 3;
  } catch (const std::exception& e) {
    std::cerr << "biocma_b200: " << e.what() << "\n";
    return 2;
  }
}
