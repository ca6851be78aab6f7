// C++ host side above the C ABI (include/bmc.h): the reference's host-facing classes for this
// path, with the same names, argument meaning and error behaviour, re-implemented on top of
// libbmc_b200.so.  Header-only, C++17, no Kokkos/Eigen.
//
//   MC::MonteCarloUnit      apps/libs/mc/public/mc/unit.hpp:33-74 (container + domain + events)
//   Simulation::Feed::*     apps/libs/simulation/public/simulation/feed_descriptor.hpp:38-161
//   ScalarSimulation        apps/libs/simulation/includes/scalar_simulation.hpp (liquid phase only)
//   SimulationUnit          apps/libs/simulation/public/simulation/simulation.hpp:44-239
//   main_loop               apps/core/src/host_specific.cpp:215-330 (step ordering = the contract)
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <vector>

#include "../../include/bmc.h"

namespace MC {
enum class Status : char { Idle = 0, Division, Exit, Dead };                                  // alias.hpp:124-130
enum class EventType : char { NewParticle = 0, Exit, Move, Death, Overflow, ChangeWeight, __COUNT__ };  // events.hpp:17-26
constexpr std::size_t number_event_type = static_cast<std::size_t>(EventType::__COUNT__);
}  // namespace MC

namespace CmaUtils {
// What rcmtool's IterationState hands the MC path and the scalar step: flat
// arrays.  `load` reads them as raw little-endian files written by biocma_mcst_b200.synth.
struct FlowMap {
  std::size_t n = 0, m = 0;
  std::vector<double> volumes, out_flows, cdf;
  std::vector<uint64_t> neighbors;
  std::vector<uint64_t> rows, cols;  // transition matrix, COO
  std::vector<double> vals;

  template <class T> static std::vector<T> read_raw(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamsize sz = f.tellg();
    f.seekg(0);
    std::vector<T> v(static_cast<std::size_t>(sz) / sizeof(T));
    f.read(reinterpret_cast<char*>(v.data()), sz);
    return v;
  }
  static FlowMap load(const std::string& dir) {
    FlowMap fm;
    fm.volumes = read_raw<double>(dir + "/volumes.raw");
    fm.out_flows = read_raw<double>(dir + "/out_flows.raw");
    fm.neighbors = read_raw<uint64_t>(dir + "/neighbors.raw");
    fm.cdf = read_raw<double>(dir + "/proba.raw");
    fm.rows = read_raw<uint64_t>(dir + "/transition_rows.raw");
    fm.cols = read_raw<uint64_t>(dir + "/transition_cols.raw");
    fm.vals = read_raw<double>(dir + "/transition_vals.raw");
    fm.n = fm.volumes.size();
    if (fm.n == 0 || fm.out_flows.size() != fm.n || fm.neighbors.size() % fm.n || fm.cdf.size() != fm.neighbors.size())
      throw std::invalid_argument("Neighbors and proba should have the same size");  // domain.cpp:52-56
    fm.m = fm.neighbors.size() / fm.n;
    return fm;
  }
  // Flow map from the triplets of a flow matrix (bmc_cma_build: what the reference gets from rcmtool's
  // get_transition_matrix / get_cumulative_probabilities / get_diag_transition, 03_cma.md:26-63)
  static FlowMap from_flows(const std::vector<double>& volumes, const std::vector<uint64_t>& from, const std::vector<uint64_t>& to,
                            const std::vector<double>& flow) {
    FlowMap fm; fm.n = volumes.size(); fm.volumes = volumes;
    uint64_t m = 0;
    if (bmc_cma_build(fm.n, flow.size(), from.data(), to.data(), flow.data(), 0, &m, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr) != BMC_OK)
      throw std::invalid_argument("invalid flow matrix");
    fm.m = m;
    std::size_t nz = fm.n;
    for (std::size_t e = 0; e < flow.size(); ++e) nz += (from[e] != to[e] && flow[e] > 0.) ? 1 : 0;
    fm.neighbors.resize(fm.n * m); fm.cdf.resize(fm.n * m); fm.out_flows.resize(fm.n);
    fm.rows.resize(nz); fm.cols.resize(nz); fm.vals.resize(nz);
    if (bmc_cma_build(fm.n, flow.size(), from.data(), to.data(), flow.data(), m, &m, fm.neighbors.data(), fm.cdf.data(), fm.out_flows.data(),
                      fm.rows.data(), fm.cols.data(), fm.vals.data()) != BMC_OK)
      throw std::invalid_argument("invalid flow matrix");
    return fm;
  }
  // One flow map of an rcmtool case directory, as far as the format can be read off the one case in the reference
  // tree (apps/api/tests/data/0d/{cma_case, vofL.raw, flowL.raw}; the format itself belongs to the un-vendored
  // rcmtool crate):  vofL.raw = u32 n, then n f64 liquid volumes;  flowL.raw = u32 rows, u32 cols, then triplets
  // {u64 row, u64 col, f64 flow} up to the end of the file.  `cma_case` names the two files (length-prefixed strings).
  struct CmaCase { std::vector<std::string> files; uint32_t header[3] = {0, 0, 0}; double t_per_flowmap = 0.; };
  static CmaCase read_cma_case(const std::string& dir) {
    const auto raw = read_raw<unsigned char>(dir + "/cma_case");
    CmaCase c;
    if (raw.size() < 20) throw std::runtime_error("cma_case: file too short");
    std::memcpy(c.header, raw.data(), 12);
    std::memcpy(&c.t_per_flowmap, raw.data() + 12, 8);
    for (std::size_t o = 20; o + 4 < raw.size(); ++o) {  // length-prefixed ASCII strings ending in ".raw"
      uint32_t len = 0; std::memcpy(&len, raw.data() + o, 4);
      if (len < 5 || len > 4096 || o + 4 + len > raw.size()) continue;
      const std::string name(reinterpret_cast<const char*>(raw.data() + o + 4), len);
      bool ascii = true; for (char ch : name) ascii = ascii && ch >= 0x20 && ch < 0x7f;
      if (ascii && name.size() > 4 && name.compare(name.size() - 4, 4, ".raw") == 0) { c.files.push_back(name); o += 3 + len; }
    }
    if (c.files.size() < 2) throw std::runtime_error("cma_case: volume / flow file names not found");
    return c;
  }
  static FlowMap load_cma_case(const std::string& dir) {
    const CmaCase c = read_cma_case(dir);
    auto find = [&](const char* key) { for (const auto& f : c.files) if (f.find(key) != std::string::npos) return dir + "/" + f; throw std::runtime_error(std::string("cma_case: no ") + key); };
    const auto vraw = read_raw<unsigned char>(find("vof"));
    uint32_t n = 0; std::memcpy(&n, vraw.data(), 4);
    if (vraw.size() != 4 + 8 * static_cast<std::size_t>(n) || n == 0) throw std::runtime_error("vofL.raw: size does not match its count");
    std::vector<double> vol(n); std::memcpy(vol.data(), vraw.data() + 4, 8 * static_cast<std::size_t>(n));
    const auto fraw = read_raw<unsigned char>(find("flow"));
    uint32_t nr = 0, ncol = 0; std::memcpy(&nr, fraw.data(), 4); std::memcpy(&ncol, fraw.data() + 4, 4);
    if (nr != n || ncol != n || (fraw.size() - 8) % 24) throw std::runtime_error("flowL.raw: shape does not match the volumes");
    std::vector<uint64_t> from, to; std::vector<double> flow;
    for (std::size_t o = 8; o + 24 <= fraw.size(); o += 24) {
      uint64_t r, cc; double f; std::memcpy(&r, fraw.data() + o, 8); std::memcpy(&cc, fraw.data() + o + 8, 8); std::memcpy(&f, fraw.data() + o + 16, 8);
      if (r != cc && f > 0.) { from.push_back(r); to.push_back(cc); flow.push_back(f); }
    }
    if (n == 1) return zero_d(vol[0]);
    return from_flows(vol, from, to, flow);
  }
  // 0D reactor (apps/api/tests/data/0d: one compartment of 0.02 m3)
  static FlowMap zero_d(double volume = 0.02) {
    FlowMap fm; fm.n = 1; fm.m = 1; fm.volumes = {volume}; fm.out_flows = {0.}; fm.cdf = {0.}; fm.neighbors = {0};
    return fm;
  }
  // stirred-tank ring of n compartments exchanging with both neighbours (self-contained cases for the CLI)
  static FlowMap ring(std::size_t n, double v_total, double exchange_rate /* 1/s */) {
    if (n == 1) return zero_d(v_total);
    FlowMap fm; fm.n = n; fm.m = 2;
    fm.volumes.assign(n, v_total / static_cast<double>(n)); fm.out_flows.assign(n, 0.);
    fm.neighbors.resize(2 * n); fm.cdf.resize(2 * n);
    const double f = 0.5 * exchange_rate * v_total / static_cast<double>(n);
    for (std::size_t i = 0; i < n; ++i) {
      fm.neighbors[2 * i] = (i + 1) % n; fm.neighbors[2 * i + 1] = (i + n - 1) % n;
      fm.cdf[2 * i] = 0.5; fm.cdf[2 * i + 1] = 1.0; fm.out_flows[i] = 2 * f;
      fm.rows.push_back(i); fm.cols.push_back((i + 1) % n); fm.vals.push_back(f);
      fm.rows.push_back(i); fm.cols.push_back((i + n - 1) % n); fm.vals.push_back(f);
      fm.rows.push_back(i); fm.cols.push_back(i); fm.vals.push_back(-2 * f);
    }
    return fm;
  }
  double total_volume() const { double s = 0; for (double v : volumes) s += v; return s; }
};
// Flow-map schedule of a run (CmaUtils::TransitionnerPtrType, used by host_specific.cpp:81-87, 263-266; the class itself
// lives in the un-vendored rcmtool crate): `size()` flow maps, each valid for t_per_flow_map seconds, visited in a
// loop (the reference's "n14" case rotates 14 maps, tools/cases.xml:21-46).
//   need_advance(t, d_t): true when the map that covers time t is not the current one
//   advance(t, d_t):      makes it current and returns it (main_loop then calls simulation.updateHydro)
class Transitioner {
 public:
  Transitioner(std::vector<FlowMap> maps, double t_per_flow_map) : maps_(std::move(maps)), t_per_(t_per_flow_map) {
    if (maps_.empty()) throw std::invalid_argument("Transitioner: no flow map");
    for (const auto& m : maps_) if (m.n != maps_[0].n) throw std::invalid_argument("Transitioner: flow maps of different size");
  }
  std::size_t size() const noexcept { return maps_.size(); }
  std::size_t index_at(double t) const noexcept {
    if (maps_.size() == 1 || !(t_per_ > 0.)) return 0;
    return static_cast<std::size_t>(std::floor(t / t_per_)) % maps_.size();
  }
  bool need_advance(double t, double /*d_t*/) const noexcept { return index_at(t) != current_; }
  const FlowMap& advance(double t, double /*d_t*/) noexcept { current_ = index_at(t); return maps_[current_]; }
  const FlowMap& get_current() const noexcept { return maps_[current_]; }
  std::size_t current_index() const noexcept { return current_; }
  double t_per_flow_map() const noexcept { return t_per_; }
  // compute_n_per_flowmap (global_initaliser.cpp:103-114)
  std::size_t n_per_flowmap(double d_t) const noexcept {
    return (t_per_ == 0. || maps_.size() == 1) ? 1 : static_cast<std::size_t>(t_per_ / d_t) + 1;
  }

 private:
  std::vector<FlowMap> maps_; double t_per_; std::size_t current_ = 0;
};
// get_time_step: min(F/V)/100 although named residence time (utils.cpp:47-74, global_initaliser.cpp:65-72; Q10)
inline double get_time_step(const FlowMap& fm) {
  double m = std::numeric_limits<double>::max();
  for (std::size_t i = 0; i < fm.n; ++i) if (fm.out_flows[i] > 0) m = std::min(m, fm.volumes[i] / fm.out_flows[i]);
  return m / 100.;
}
}  // namespace CmaUtils

namespace Simulation {
namespace Feed {  // feed_descriptor.hpp:38-101, feed_descriptor.cpp:19-28
struct Constant {};
struct Exponential { double f0, alpha; };
struct Linear { double f0, df; };
using FeedTypeVariant = std::variant<Constant, Linear, Exponential>;
struct FeedValue { double concentration; std::size_t species_index; };
struct FeedDescriptor {
  double flow{};
  std::vector<FeedValue> values;
  std::size_t input_position{};
  std::optional<std::size_t> output_position;
  FeedTypeVariant extra;
  bool use_relative_time = true;
  void update(double t, double /*d_t*/) noexcept {
    if (auto* e = std::get_if<Exponential>(&extra)) flow = e->f0 + std::exp(e->alpha * t);  // sic (Q11)
    else if (auto* l = std::get_if<Linear>(&extra)) flow = l->f0 + t * l->df;
  }
};
struct FeedFactory {  // feed_descriptor.cpp:33-120
  static FeedDescriptor constant(double flow, double concentration, std::size_t species_index, std::size_t input_position,
                                 std::optional<std::size_t> output_position = std::nullopt, bool set_output = true) {
    if (set_output && !output_position) output_position = input_position;  // chemostat: outlet = inlet compartment
    if (flow < 0.) throw std::invalid_argument("FeedException: NegativeFlow");
    if (concentration <= 0.) throw std::invalid_argument("FeedException: NegativeConcentration");
    return FeedDescriptor{flow, {{concentration, species_index}}, input_position, output_position, Constant{}, true};
  }
  static FeedDescriptor linear(double flow, double df, double concentration, std::size_t species_index, std::size_t input_position,
                               std::optional<std::size_t> output_position = std::nullopt, bool set_output = true) {
    auto fd = constant(flow, concentration, species_index, input_position, output_position, set_output);
    fd.extra = Linear{flow, df};
    return fd;
  }
};
}  // namespace Feed

// Liquid scalar field: dm/dt = C*M - C*sink + sources, explicit Euler, C = mass * V^-1
// (implScalar.cpp:251-266).  Concentrations are species-fastest like the kernel view
// (alias.hpp:169-173); `sources` uses the same layout here.
class ScalarSimulation {
 public:
  ScalarSimulation(std::size_t n_compartments, std::size_t n_species, const std::vector<double>& volumes)
      : n_r(n_species), n_c(n_compartments), vol(volumes), C(n_species * n_compartments, 0.), mass(C.size(), 0.),
        sources(C.size(), 0.), sink(n_compartments, 0.) {}
  std::size_t n_row() const noexcept { return n_r; }
  std::size_t n_col() const noexcept { return n_c; }
  void set_concentration(const std::vector<double>& c) {
    if (c.size() != C.size()) throw std::invalid_argument("bad concentration size");
    for (double v : c) if (v < 0) throw std::invalid_argument("initial concentrations must be >= 0");  // simulation.cpp:180-199
    C = c;
    for (std::size_t j = 0; j < n_c; ++j) for (std::size_t s = 0; s < n_r; ++s) mass[s + n_r * j] = C[s + n_r * j] * vol[j];
  }
  void setVolumes(const std::vector<double>& v) { vol = v; }
  void set_transition(const CmaUtils::FlowMap& fm) { rows = fm.rows; cols = fm.cols; vals = fm.vals; }
  void set_feed(std::size_t i_r, std::size_t i_c, double val) { sources[i_r + n_r * i_c] += val; }  // scalar_simulation.hpp:178-182
  void set_sink(std::size_t i_compartment, double val) { sink[i_compartment] += val; }               // :184-188
  void set_zero_contribs() { std::fill(sources.begin(), sources.end(), 0.); std::fill(sink.begin(), sink.end(), 0.); }  // :162-170
  double volume_at(std::size_t i) const { return vol[i]; }
  void performStep(double d_t) {  // implScalar.cpp:251-266
    std::vector<double> dm(C.size(), 0.);
    for (std::size_t e = 0; e < vals.size(); ++e)
      for (std::size_t s = 0; s < n_r; ++s) dm[s + n_r * cols[e]] += C[s + n_r * rows[e]] * vals[e];
    for (std::size_t j = 0; j < n_c; ++j)
      for (std::size_t s = 0; s < n_r; ++s) {
        const std::size_t k = s + n_r * j;
        dm[k] += -C[k] * sink[j] + sources[k];
        mass[k] += d_t * dm[k];
        C[k] = mass[k] * (1.0 / vol[j]);
      }
  }
  std::vector<double>& getConcentrationData() { return C; }
  std::vector<double>& getContributionData() { return sources; }
  const std::vector<double>& getVolume() const { return vol; }

 private:
  std::size_t n_r, n_c;
  std::vector<double> vol, C, mass, sources, sink;
  std::vector<uint64_t> rows, cols;
  std::vector<double> vals;
};
}  // namespace Simulation

namespace MC {
// Owner of the GPU-resident Monte-Carlo state: ParticlesContainer + ReactorDomain + EventContainer
// of the reference, behind one bmc_ctx.
class MonteCarloUnit {
 public:
  MonteCarloUnit(int model, std::size_t n_species, std::size_t n_compartments, uint64_t seed, uint32_t rank = 0, int device = 0,
                 int n_var_udf = 32) {
    bmc_config cfg{};
    cfg.device = device; cfg.model = model; cfg.n_var_udf = n_var_udf; cfg.n_species = n_species;
    cfg.n_compartments = n_compartments; cfg.seed = seed; cfg.rank = rank;
    // load_tuning_constant (mc/src/unit.cpp:302-343): a value outside (min, max] is ignored (read_env_valid_or, :28-41)
    auto valid_or = [](const char* name, double vdefault, double lo, double hi) {
      const char* e = std::getenv(name);
      if (!e) return vdefault;
      const double v = std::atof(e);
      return (v > lo && v <= hi) ? v : vdefault;
    };
    cfg.allocation_factor = valid_or("BIOMC_MC_ALLOC_FACTOR", 0.0, 0., 5.);   // 0 = the library's default
    cfg.buffer_ratio = valid_or("BIOMC_MC_BUFFER_RATIO", 0.0, 0., 1.);
    if (const char* e = std::getenv("BIOMC_MC_REMOVE_RATIO_THRESHOLD")) cfg.dead_particle_ratio_threshold = std::atof(e);
    if (const char* e = std::getenv("BIOMC_MC_MINIMUM_REMOVAL")) cfg.minimum_dead_particle_removal = std::strtoull(e, nullptr, 10);
    if (bmc_create(&ctx, &cfg) != BMC_OK) throw std::runtime_error("bmc_create failed");
    n_comp = n_compartments;
  }
  ~MonteCarloUnit() { if (ctx) bmc_destroy(&ctx); }
  MonteCarloUnit(const MonteCarloUnit&) = delete;
  MonteCarloUnit& operator=(const MonteCarloUnit&) = delete;

  void check(int rc) const { if (rc != BMC_OK) throw std::runtime_error(bmc_last_error(ctx)); }
  // MC::init<M> (mcinit.hpp:67-105) + post_init_weight (unit.cpp:232-257)
  double init(uint64_t n_particles, bool uniform_init, double x0, double total_volume) {
    double total_mass = 0;
    check(bmc_init_particles(ctx, n_particles, uniform_init ? 1 : 0, nullptr, &total_mass));
    init_weight = (x0 * total_volume) / total_mass;
    check(bmc_set_weight(ctx, init_weight));
    return total_mass;
  }
  // ParticlesContainer accessors (particles_container.hpp:445-500)
  uint64_t n_particles() const { return counters().n_used; }
  uint64_t capacity() const { return counters().capacity; }
  uint64_t get_inactive() const { return counters().n_inactive; }
  uint64_t n_particle() const { const auto c = counters(); return c.n_used - c.n_inactive; }  // unit.cpp:167-173
  void force_remove_dead() { check(bmc_compact(ctx)); }
  // EventContainer::get<event>() (events.hpp:149-264)
  template <EventType e> uint64_t get_event() const { return counters().events[static_cast<std::size_t>(e)]; }
  std::array<uint64_t, number_event_type> events() const {
    const auto c = counters(); std::array<uint64_t, number_event_type> a{};
    for (std::size_t i = 0; i < number_event_type; ++i) a[i] = c.events[i];
    return a;
  }
  std::vector<uint64_t> getRepartition() const {  // unit.cpp:190-230
    std::vector<uint64_t> r(n_comp);
    check(bmc_repartition(ctx, r.data()));
    return r;
  }
  bmc_counters counters() const { bmc_counters c{}; check(bmc_get_counters(ctx, &c)); return c; }
  // SerDe::save_simulation / load_simulation (apps/core/src/serde.cpp:64-219): `<results>_serde_<rank>.raw`
  void save(const std::string& path) const {
    uint64_t bytes = 0;
    check(bmc_checkpoint_size(ctx, &bytes));
    std::vector<char> buf(bytes);
    check(bmc_checkpoint_save(ctx, buf.data(), bytes));
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("Error opening file: " + path);
    f.write(buf.data(), static_cast<std::streamsize>(bytes));
  }
  void load(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot read file");  // serde.cpp:52
    std::vector<char> buf(static_cast<std::size_t>(f.tellg()));
    f.seekg(0);
    f.read(buf.data(), static_cast<std::streamsize>(buf.size()));
    check(bmc_checkpoint_load(ctx, buf.data(), buf.size()));
    bmc_counters c{}; check(bmc_get_counters(ctx, &c));
  }
  bmc_ctx* handle() const { return ctx; }
  double init_weight = 0;

 private:
  bmc_ctx* ctx = nullptr;
  std::size_t n_comp = 0;
};
}  // namespace MC

namespace Simulation {
struct Dimensions { std::size_t n_species, n_compartment; };

class SimulationUnit {
 public:
  SimulationUnit(std::unique_ptr<MC::MonteCarloUnit>&& unit, const CmaUtils::FlowMap& fm, std::size_t n_species,
                 const std::vector<double>& initial_concentration)
      : mc_unit(std::move(unit)), liquid_scalar(fm.n, n_species, fm.volumes), dims{n_species, fm.n} {
    liquid_scalar.set_concentration(initial_concentration);
    updateHydro(fm);
  }
  // simulation.cpp:95-140 updateHydro -> ReactorDomain::update + setVolumes/set_transition
  void updateHydro(const CmaUtils::FlowMap& fm) {
    mc_unit->check(bmc_domain_update(mc_unit->handle(), fm.volumes.data(), fm.n > 1 ? fm.neighbors.data() : nullptr,
                                     fm.out_flows.data(), fm.n > 1 ? fm.cdf.data() : nullptr, fm.n > 1 ? fm.m : 0));
    liquid_scalar.setVolumes(fm.volumes);
    liquid_scalar.set_transition(fm);
  }
  void add_feed(Feed::FeedDescriptor&& fd) { liquid_feeds.push_back(std::move(fd)); }
  // simulation.model.cpp:73-122
  void update_feed(double d_t, bool update_scalar = true) {
    std::vector<bmc_leaving_flow> outlets;
    for (auto& fd : liquid_feeds) {
      fd.update(fd.use_relative_time ? relative_time : absolute_time, d_t);
      if (update_scalar) {
        for (const auto& v : fd.values) liquid_scalar.set_feed(v.species_index, fd.input_position, fd.flow * v.concentration);
        if (fd.output_position) liquid_scalar.set_sink(*fd.output_position, fd.flow);
      }
      if (fd.output_position)  // only feeds with an outlet fill a slot (Q21)
        outlets.push_back(bmc_leaving_flow{*fd.output_position, fd.flow, liquid_scalar.volume_at(*fd.output_position)});
    }
    mc_unit->check(bmc_set_leaving_flows(mc_unit->handle(), outlets.size(), outlets.data()));
  }
  void ode_step(double d_t) { liquid_scalar.performStep(d_t); }                       // simulation.model.cpp:131-154
  double advance(double d_t) { absolute_time += d_t; relative_time += d_t; return absolute_time; }  // :125-129
  void clearContribution() { liquid_scalar.set_zero_contribs(); }                       // sync_prepare_next, sync.cpp:89-116
  // simulation.hpp:183-239: pre_cycle / launch_model / launch_move / post_cycle, sources back on the host
  void cycleProcess(double d_t) {
    if (mc_unit->n_particles() == 0) return;
    auto* h = mc_unit->handle();
    mc_unit->check(bmc_set_concentrations(h, liquid_scalar.getConcentrationData().data()));
    mc_unit->check(bmc_cycle(h, d_t));
    mc_unit->check(bmc_get_sources(h, liquid_scalar.getContributionData().data()));  // scatter_contribute + synchro_sources
  }
  // SerDe::load_simulation (serde.cpp:139-219): MC unit + liquid concentrations of the saved run
  void load_serde(const std::string& path) {
    mc_unit->load(path);
    std::vector<double> c(dims.n_species * dims.n_compartment);
    mc_unit->check(bmc_get_concentrations(mc_unit->handle(), c.data()));
    liquid_scalar.set_concentration(c);  // total mass = C * V is rebuilt from the archived concentrations, as the reference does
    // the source terms of the last cycle are still to be applied by the next ODE step (main_loop order): restored too
    // (the reference's archive drops them, so its resumed liquid lags the uninterrupted one by one step of uptake)
    mc_unit->check(bmc_get_sources(mc_unit->handle(), liquid_scalar.getContributionData().data()));
  }
  void save_serde(const std::string& path) {
    mc_unit->check(bmc_set_concentrations(mc_unit->handle(), liquid_scalar.getConcentrationData().data()));
    mc_unit->save(path);
  }
  Dimensions getDimensions() const { return dims; }
  double absolute() const { return absolute_time; }
  std::unique_ptr<MC::MonteCarloUnit> mc_unit;
  ScalarSimulation liquid_scalar;

 private:
  Dimensions dims;
  std::vector<Feed::FeedDescriptor> liquid_feeds;
  double absolute_time = 0., relative_time = 0.;
};
}  // namespace Simulation

namespace Core {
struct SimulationParameters {  // simulation_parameters.hpp
  double d_t = 0., final_time = 0.;
  uint64_t number_particle = 0, number_exported_result = 0;
  double biomass_initial_concentration = 0.5;  // X0, simulation_parameters.cpp:45
  bool uniform_mc_init = true;
};
struct Records {  // records/* of the result file (main_exporter.cpp:32-191, partial_exporter.cpp:26-161)
  std::vector<double> time, concentration_liquid;  // [t][n_comp][n_species]
  std::vector<uint64_t> number_particle;           // [t][n_comp]
  std::vector<uint64_t> tallies;                   // [t][6]
};
// get_n_interval, host_specific.cpp:106-125
inline void get_n_interval(const SimulationParameters& p, std::size_t& n_iter, std::size_t& dump_interval) {
  n_iter = static_cast<std::size_t>(p.final_time / p.d_t) + 1;
  const std::size_t dump_number = std::min<std::size_t>(n_iter, p.number_exported_result) - 1;
  dump_interval = (p.number_exported_result != 0 && dump_number != 0) ? n_iter / dump_number + 1 : n_iter + 1;
}
// main_loop, host_specific.cpp:215-330 (single rank; the ordering is the contract)
inline Records main_loop(const SimulationParameters& params, Simulation::SimulationUnit& simulation,
                         CmaUtils::Transitioner* d_transitionner = nullptr) {
  Records rec;
  std::size_t n_iter = 0, dump_interval = 0;
  get_n_interval(params, n_iter, dump_interval);
  const double d_t = params.d_t;
  auto dump = [&]() {
    rec.time.push_back(simulation.absolute());
    const auto& c = simulation.liquid_scalar.getConcentrationData();
    rec.concentration_liquid.insert(rec.concentration_liquid.end(), c.begin(), c.end());
    const auto rep = simulation.mc_unit->getRepartition();
    rec.number_particle.insert(rec.number_particle.end(), rep.begin(), rep.end());
    const auto ev = simulation.mc_unit->events();
    rec.tallies.insert(rec.tallies.end(), ev.begin(), ev.end());
  };
  simulation.update_feed(d_t);  // :241 (before the loop: the first step sees the feed twice, as in the reference)
  if (d_transitionner) simulation.updateHydro(d_transitionner->advance(simulation.absolute(), d_t));  // UPDATE_HYDRO_STEP, :257
  for (std::size_t it = 0; it < n_iter; ++it) {
    if (d_transitionner && d_transitionner->need_advance(simulation.absolute(), d_t))  // :263-266
      simulation.updateHydro(d_transitionner->advance(simulation.absolute(), d_t));
    if (params.number_exported_result != 0 && it % dump_interval == 0) dump();
    // sync_step: single rank -> nothing to reduce (multi-GPU: bmc_allreduce_sources inside cycleProcess)
    simulation.update_feed(d_t);
    simulation.ode_step(d_t);
    simulation.advance(d_t);
    simulation.clearContribution();
    simulation.cycleProcess(d_t);
  }
  simulation.mc_unit->force_remove_dead();  // :316
  dump();
  return rec;
}
}  // namespace Core
