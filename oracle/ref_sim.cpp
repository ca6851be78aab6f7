// =============================================================================
// oracle/ref_sim.cpp — TEST INFRASTRUCTURE, NOT THE PRODUCT.
//
// The reference's OWN Simulation::SimulationUnit (apps/libs/simulation/src/simulation.cpp, simulation.model.cpp,
// simulation.getset.cpp, scalar_init.cpp, feed_descriptor.cpp, implScalar.cpp, hydro/*.cpp + simulation.hpp), compiled
// where the sources lie over oracle/kokkos_shim, oracle/eigen_shim and oracle/rust_shim, and stepped by the body of
// the reference's main loop (apps/core/src/host_specific.cpp:281-291):
//     simulation.update_feed(d_t); simulation.ode_step(d_t); simulation.advance(d_t);
//     [sync_prepare_next ->] simulation.clearContribution();
//     simulation.cycleProcess(container, d_t, functors);
// Nothing of cycleProcess / post_cycle / scatter_contribute / update_feed / ode_step is restated here (ref_driver.cpp
// and ref_liquid.cpp restate those call sequences because they predate the Eigen stand-in); this file only builds the
// objects the way global_initaliser.cpp does and moves arrays in and out.  The particle container is a
// ParticlesContainer<Tap<M>> (ref_driver.cpp: Tap forwards every model hook after selecting the particle's random
// stream), passed to cycleProcess like the variant alternative the reference visits.
// =============================================================================
#include "ref_driver.cpp"  // shim hooks, Tap<M>, MonodQ1 (same translation unit: one definition of the hooks)

#include <common/eigen_diag.hpp>
#include <Eigen/Core>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <scalar_simulation.hpp>
#include <simulation/feed_descriptor.hpp>
#include <simulation/scalar_initializer.hpp>
#include <simulation/simulation.hpp>

namespace {
struct ISim {
  virtual ~ISim() = default;
  std::string err;
  virtual void set_particles(size_t n, const float* props, const uint64_t* pos, double weight) = 0;
  virtual void get_particles(size_t n, float* props, uint64_t* pos, uint8_t* st, float* ah, float* ad) = 0;
  virtual void update_hydro(const double* vol, const uint64_t* neigh, const double* proba, const double* out_flows, size_t m, size_t nnz,
                            const uint64_t* rows, const uint64_t* cols, const double* vals) = 0;
  // two-phase cases: the gas part of the NEXT update_hydro's iteration state (volumes, transition, energy dissipation)
  virtual void set_gas_hydro(const double* gvol, size_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals, const double* eps) = 0;
  virtual void get_gas(double* conc, double* mtr) = 0;
  virtual void step(double d_t) = 0;
  virtual void get_concentrations(double* out) = 0;
  virtual void get_sources(double* out) = 0;
  virtual void counters(unsigned long long* c) = 0;
};

template <class M> struct Sim final : ISim {
  using Container = MC::ParticlesContainer<M>;
  using Functors = Simulation::KernelInline::CycleFunctors<ComputeSpace, M>;
  size_t ns, nc;
  uint64_t seed; uint32_t rank = 0, step_id = 0;
  std::vector<double> vol, gvol;
  PhaseStateWrapper gas_state; std::vector<double> eps; bool two_phase = false;
  std::unique_ptr<Simulation::SimulationUnit> sim;
  Container container;
  std::unique_ptr<Functors> functors;
  MC::RuntimeParameters rt{0, 0.6, 1.5, 0.0, 0.01};
  KernelDispatchOptions opts{};
  unsigned long long total_out = 0, total_new = 0, n_compactions = 0, last_out = 0, last_waiting = 0;

  // gas_volumes != nullptr: two-phase flow (gas concentrations g0, gas feeds, mass transfer: kla_fixed per species, or the
  // turbulence correlation when kla_fixed == nullptr)
  Sim(size_t n_species, size_t n_comp, uint64_t seed_, const double* volumes, const double* c0, size_t n_feeds, const uint64_t* f_species,
      const uint64_t* f_in, const uint64_t* f_out, const double* f_flow, const double* f_conc, const double* gas_volumes = nullptr,
      const double* g0 = nullptr, size_t n_gas_feeds = 0, const uint64_t* gf_species = nullptr, const uint64_t* gf_in = nullptr,
      const uint64_t* gf_out = nullptr, const double* gf_flow = nullptr, const double* gf_conc = nullptr, const double* kla_fixed = nullptr)
      : ns(n_species), nc(n_comp), seed(seed_), vol(volumes, volumes + n_comp) {
    two_phase = gas_volumes != nullptr;
    if (two_phase) gvol.assign(gas_volumes, gas_volumes + n_comp);
    auto unit = std::make_unique<MC::MonteCarloUnit>();
    unit->domain = MC::ReactorDomain(std::span<double>(vol));  // mc/public/mc/mcinit.hpp:83
    // concentrations from a functor, like ScalarInitialiserType::Uniform / Local (scalar_factory.cpp)
    std::vector<double> init(c0, c0 + ns * nc);
    Simulation::ScalarInitializer si{};
    si.n_species = ns; si.volumesliq = std::span<double>(vol); si.type = Simulation::ScalarInitialiserType::Local;
    si.liquid_f_init = [init, n_species](std::size_t i, std::size_t j) { return init[i + n_species * j]; };
    si.gas_flow = two_phase;
    if (two_phase) {
      std::vector<double> ginit(g0, g0 + ns * nc);
      si.volumesgas = std::span<double>(gvol);
      si.gas_f_init = [ginit, n_species](std::size_t i, std::size_t j) { return ginit[i + n_species * j]; };
    }
    auto feed = Simulation::Feed::SimulationFeed::empty();
    for (size_t k = 0; k < n_gas_feeds; ++k)
      feed.add_gas(Simulation::Feed::FeedFactory::constant(gf_flow[k], gf_conc[k], gf_species[k], gf_in[k], std::optional<std::size_t>(gf_out[k])));
    for (size_t k = 0; k < n_feeds; ++k)  // one descriptor per entry: FeedFactory::constant (feed_descriptor.cpp:97-112)
      feed.add_liquid(Simulation::Feed::FeedFactory::constant(f_flow[k], f_conc[k], f_species[k], f_in[k], std::optional<std::size_t>(f_out[k])));
    sim = std::make_unique<Simulation::SimulationUnit>(std::move(unit), std::move(si), std::optional<Simulation::Feed::SimulationFeed>(std::move(feed)));
    if (two_phase) {  // global_initaliser.cpp: simulation->setMtrModel(...)
      if (kla_fixed) sim->setMtrModel(Simulation::MassTransfer::Type::MtrTypeVariant(Simulation::MassTransfer::Type::FixedKla{std::vector<double>(kla_fixed, kla_fixed + ns)}));
      else sim->setMtrModel(Simulation::MassTransfer::Type::MtrTypeVariant(Simulation::MassTransfer::Type::FlowmapTurbulence{}));
    }
  }
  void set_gas_hydro(const double* gv, size_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals, const double* e) override {
    gas_state.vol.assign(gv, gv + nc); gas_state.inv_vol.resize(nc);
    for (size_t j = 0; j < nc; ++j) gas_state.inv_vol[j] = 1.0 / gv[j];
    gas_state.out.assign(nc, 0.0);
    gas_state.coo.n = nc; gas_state.coo.r.assign(rows, rows + nnz); gas_state.coo.c.assign(cols, cols + nnz); gas_state.coo.v.assign(vals, vals + nnz);
    eps.assign(e, e + nc);
  }
  void get_gas(double* conc, double* mtr) override {
    const auto g = sim->getter().getCgasData();
    const auto m = sim->getter().getMTRData();
    if (!g || !m) throw std::runtime_error("no gas phase");
    for (size_t k = 0; k < ns * nc; ++k) { conc[k] = (*g)[k]; mtr[k] = (*m)[k]; }
  }

  void set_particles(size_t n, const float* props, const uint64_t* pos, double weight) override {
    container = Container(rt, n, 0);
    for (size_t i = 0; i < n; ++i) {
      for (size_t k = 0; k < M::n_var; ++k) container.model(i, k) = props[k * n + i];
      container.position(i) = pos ? pos[i] : 0;
      container.status(i) = MC::Status::Idle;
      container.ages(i, 0) = 0.f; container.ages(i, 1) = 0.f;
    }
    container.weights(0) = (typename M::FloatType)weight;
    functors.reset();
  }
  void get_particles(size_t n, float* props, uint64_t* pos, uint8_t* st, float* ah, float* ad) override {
    for (size_t i = 0; i < n; ++i) {
      for (size_t k = 0; k < M::n_var; ++k) props[k * n + i] = container.model(i, k);
      pos[i] = container.position(i); st[i] = (uint8_t)container.status(i);
      ah[i] = container.ages(i, 0); ad[i] = container.ages(i, 1);
    }
  }
  // SimulationUnit::updateHydro(state) (simulation.cpp:97-139): the MC domain and the liquid scalar from one iteration state
  void update_hydro(const double* v, const uint64_t* neigh, const double* proba, const double* out_flows, size_t m, size_t nnz,
                    const uint64_t* rows, const uint64_t* cols, const double* vals) override {
    auto* st = new IterationStateWrapper;
    st->liq.vol.assign(v, v + nc); st->liq.inv_vol.resize(nc);
    for (size_t j = 0; j < nc; ++j) st->liq.inv_vol[j] = 1.0 / v[j];
    st->liq.out.assign(out_flows, out_flows + nc);
    st->liq.coo.n = nc; st->liq.coo.r.assign(rows, rows + nnz); st->liq.coo.c.assign(cols, cols + nnz); st->liq.coo.v.assign(vals, vals + nnz);
    st->neighbors.assign(neigh, neigh + nc * m); st->probability_leaving.assign(proba, proba + nc * m);
    if (two_phase) { st->gas = gas_state; st->with_gas = true; st->energy_dissipation = eps; }
    CmaUtils::IterationStatePtrType state(st);
    sim->updateHydro(state);
    functors.reset();
  }
  // one iteration of the main loop (host_specific.cpp:281-291)
  void step(double d_t) override {
    set_streams(seed, rank, step_id);
    g_cfg.per_team = opts.m_p_p_team_move;
    if (!functors) functors = std::make_unique<Functors>(sim->template init_functors<ComputeSpace, M>(container, opts));  // host_specific.cpp:251-252
    const auto ev0 = events_now();
    const size_t n_before = container.n_particles(), inactive_before = container.get_inactive();
    sim->update_feed(d_t);
    sim->ode_step(d_t);
    sim->advance(d_t);
    sim->clearContribution();  // sync_prepare_next (sync.cpp:89-95)
    sim->cycleProcess(container, d_t, *functors);
    // bookkeeping for the counters the tests compare (derived from the reference's own tallies and container extents)
    const auto ev1 = events_now();
    last_out = ev1[1] - ev0[1]; last_waiting = ev1[4] - ev0[4];
    total_out += last_out;
    const bool compacted = container.get_inactive() < inactive_before + last_out;
    if (compacted) ++n_compactions;
    const size_t after_removal = compacted ? n_before - (inactive_before + last_out) : n_before;
    total_new += container.n_particles() - after_removal;
    ++step_id;
  }
  std::array<unsigned long long, 6> events_now() const {
    std::array<unsigned long long, 6> e{};
    const auto sp = sim->getter().mc_unit()->events.get_span();
    for (int k = 0; k < 6; ++k) e[k] = sp[k];
    return e;
  }
  void get_concentrations(double* out) override {
    const auto c = sim->getter().getCliqData();  // LayoutLeft (n_species, n_comp): species fastest
    for (size_t k = 0; k < ns * nc; ++k) out[k] = c[k];
  }
  void get_sources(double* out) override {  // `sources` is LayoutRight (n_species, n_comp): transpose to species fastest
    const auto s = sim->getter().getContributionData();
    for (size_t j = 0; j < nc; ++j) for (size_t i = 0; i < ns; ++i) out[i + ns * j] = s[i * nc + j];
  }
  void counters(unsigned long long* c) override {
    const auto e = events_now();
    for (int k = 0; k < 6; ++k) c[k] = e[k];
    c[6] = container.n_particles(); c[7] = container.get_inactive(); c[8] = last_out; c[9] = 0; c[10] = last_waiting;
    c[11] = 0; c[12] = container.capacity(); c[13] = total_out; c[14] = total_new; c[15] = n_compactions;
  }
};
}  // namespace

#define SIM_TRY(h, ...) try { __VA_ARGS__; return 0; } catch (const std::exception& e) { (h)->err = e.what(); return -1; }

extern "C" {
// model ids as in ref_create: 0 fixed_length, 1 monod.  Feeds: n_feeds constant feeds {species, input, output, flow, concentration}.
void* rsim_create(int model, uint64_t n_species, uint64_t n_comp, uint64_t seed, const double* volumes, const double* c0, uint64_t n_feeds,
                  const uint64_t* f_species, const uint64_t* f_in, const uint64_t* f_out, const double* f_flow, const double* f_conc) {
  try {
    if (model == 0) return new Sim<Tap<Models::FixedLength>>(n_species, n_comp, seed, volumes, c0, n_feeds, f_species, f_in, f_out, f_flow, f_conc);
    if (model == 1) return new Sim<Tap<MonodQ1>>(n_species, n_comp, seed, volumes, c0, n_feeds, f_species, f_in, f_out, f_flow, f_conc);
  } catch (const std::exception& e) { std::fprintf(stderr, "rsim_create: %s\n", e.what()); }
  return nullptr;
}
void rsim_destroy(void* h) { delete static_cast<ISim*>(h); }
const char* rsim_last_error(void* h) { return static_cast<ISim*>(h)->err.c_str(); }
int rsim_set_particles(void* h, uint64_t n, const float* props, const uint64_t* pos, double weight) {
  auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->set_particles(n, props, pos, weight));
}
int rsim_get_particles(void* h, uint64_t n, float* props, uint64_t* pos, uint8_t* st, float* ah, float* ad) {
  auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->get_particles(n, props, pos, st, ah, ad));
}
int rsim_update_hydro(void* h, const double* vol, const uint64_t* neigh, const double* proba, const double* out_flows, uint64_t m, uint64_t nnz,
                      const uint64_t* rows, const uint64_t* cols, const double* vals) {
  auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->update_hydro(vol, neigh, proba, out_flows, m, nnz, rows, cols, vals));
}
int rsim_step(void* h, double d_t) { auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->step(d_t)); }
int rsim_get_concentrations(void* h, double* out) { auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->get_concentrations(out)); }
int rsim_get_sources(void* h, double* out) { auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->get_sources(out)); }
int rsim_get_counters(void* h, unsigned long long* c) { auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->counters(c)); }
// two-phase SimulationUnit (model ids as above); kla_fixed == NULL selects Type::FlowmapTurbulence
void* rsim_create_two_phase(int model, uint64_t n_species, uint64_t n_comp, uint64_t seed, const double* volumes, const double* c0, uint64_t n_feeds,
                            const uint64_t* f_species, const uint64_t* f_in, const uint64_t* f_out, const double* f_flow, const double* f_conc,
                            const double* gas_volumes, const double* g0, uint64_t n_gas_feeds, const uint64_t* gf_species, const uint64_t* gf_in,
                            const uint64_t* gf_out, const double* gf_flow, const double* gf_conc, const double* kla_fixed) {
  try {
    if (model == 0) return new Sim<Tap<Models::FixedLength>>(n_species, n_comp, seed, volumes, c0, n_feeds, f_species, f_in, f_out, f_flow, f_conc,
                                                             gas_volumes, g0, n_gas_feeds, gf_species, gf_in, gf_out, gf_flow, gf_conc, kla_fixed);
    if (model == 2) return new Sim<Tap<Models::SimpleAcetate>>(n_species, n_comp, seed, volumes, c0, n_feeds, f_species, f_in, f_out, f_flow, f_conc,
                                                               gas_volumes, g0, n_gas_feeds, gf_species, gf_in, gf_out, gf_flow, gf_conc, kla_fixed);
  } catch (const std::exception& e) { std::fprintf(stderr, "rsim_create_two_phase: %s\n", e.what()); }
  return nullptr;
}
int rsim_set_gas_hydro(void* h, const double* gvol, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals, const double* eps) {
  auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->set_gas_hydro(gvol, nnz, rows, cols, vals, eps));
}
int rsim_get_gas(void* h, double* conc, double* mtr) { auto* s = static_cast<ISim*>(h); SIM_TRY(s, s->get_gas(conc, mtr)); }
}  // extern "C"
