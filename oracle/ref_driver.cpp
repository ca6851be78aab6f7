// =============================================================================
// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT THE PRODUCT.
//
// Drives the REFERENCE'S OWN hot-path sources, compiled from /root/reference where
// they lie (never copied), on the CPU through oracle/kokkos_shim (a serial stand-in
// for Kokkos).  Built by `make -C oracle ref` into oracle/_ref/libbmc_ref.so.  Used
// only by tests/ and by tools that generate tests/golden/ fixtures, to pin
// oracle/bmc_oracle.cpp (and through it the CUDA path) to the reference:
//
//   reference code that runs here, unmodified
//     Simulation::KernelInline::CycleFunctors<Space,Model>::update / launch_model / launch_move
//                                   simulation/kernels/kernels.hpp:49-224
//     CycleFunctor<M>               simulation/kernels/model_kernel.hpp:163-268
//     ContributionFunctor<M>        simulation/kernels/contribution_kernel.hpp:48-186
//     MoveFunctor (move + leave)    simulation/kernels/move_kernel.hpp:140-669
//     probability_leaving<>         simulation/probability_leaving.hpp:16-46
//     MC::ParticlesContainer<M>     mc/particles_container.hpp (handle_division, merge_buffer,
//                                   update_and_remove_inactive, CompactParticlesFunctor, InsertFunctor)
//     MC::ReactorDomain             mc/domain.hpp + mc/src/domain.cpp
//     MC::EventContainer            mc/events.hpp
//     Models::FixedLength, Models::Monod, Models::SimpleAcetate   models/*.hpp
//     Common::c_league_size         common/src/common.cpp
//     PostProcessing::get_properties core/post_process.hpp:173-250 (+ GetPropertiesFunctor :33-147)
//   what this file adds
//     the body of SimulationUnit::cycleProcess / post_cycle (simulation/simulation.hpp:183-239): the same calls in the
//     same order, on a bare container + domain + concentration view so that every input of a step can be set from
//     outside (the fixtures).  The reference's SimulationUnit ITSELF runs in oracle/ref_sim.cpp (Eigen and rcmtool
//     stand-ins), where nothing of the step is restated; tests/test_reference_simulation_unit.py ties the two together;
//     Tap<M>: forwards every model hook unchanged after telling the shim's random generator which particle
//     it serves (the reference's pool is not indexable by particle; DESIGN.md §4 defines the streams);
//     MonodQ1: SURVEY.md Q1 — monod.hpp predates the 7-argument hook concept and has no n_c; the wrapper
//     calls Models::Monod::{init,update,division,mass} as they are and mirrors phi_s_c into contribs(.,0).
//
// Not reproduced (shim limits, stated in DESIGN.md): parallel execution order, the XorShift1024 pool,
// ScatterView duplication (float sums are accumulated in particle order).
// =============================================================================
#include <Kokkos_Core.hpp>

#include <mc/macros.hpp>
// monod.hpp ends with CHECK_MODEL(Monod), which cannot hold: check the models we wrap ourselves
#undef CHECK_MODEL
#define CHECK_MODEL(name)
#include <models/monod.hpp>
#undef CHECK_MODEL
#define CHECK_MODEL(name) static_assert(ModelType<name>, #name);
#include <models/fixed_length.hpp>
#include <models/simple_acetate.hpp>
#include <simulation/kernels/kernels.hpp>
#include <core/post_process.hpp>  // PostProcessing::get_properties / GetPropertiesFunctor (apps/core/public/core/post_process.hpp:33-250)
#ifdef BMC_REF_WITH_UDF
#include <models/udf_model.hpp>  // hooks: apps/udf_model/minimal.cpp through UnsafeUDF::Loader (oracle/ref_udf.cpp)
#endif

#include <chrono>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ------------------------------------------------------------------ shim hooks
// Stream selection state: the configuration of the running kernel is shared (written by the launching thread
// before the parallel region), the position inside a stream is per thread.
namespace {
struct RngConfig { uint32_t key[2] = {0, 0}; uint32_t rank = 0, step = 0; Kokkos::shim::Mode mode = Kokkos::shim::Mode::Sequential; size_t per_team = 1024; };
RngConfig g_cfg;
thread_local Kokkos::shim::RngState t_rng;
int g_threads = 1;
bool g_init_kernel = false;
// BMC_REF_PROFILE=1: seconds per kernel label, printed by ref_destroy (tuning aid)
const bool g_profile = std::getenv("BMC_REF_PROFILE") != nullptr;
std::map<std::string, double> g_times; std::string g_label; std::chrono::steady_clock::time_point g_t0;
inline Kokkos::shim::RngState& synced() {
  Kokkos::shim::RngState& r = t_rng;
  r.key[0] = g_cfg.key[0]; r.key[1] = g_cfg.key[1]; r.rank = g_cfg.rank; r.step = g_cfg.step; r.mode = g_cfg.mode; r.per_team = g_cfg.per_team;
  return r;
}
inline void set_streams(uint64_t seed, uint32_t rank, uint32_t step) {
  g_cfg.key[0] = (uint32_t)seed; g_cfg.key[1] = (uint32_t)(seed >> 32); g_cfg.rank = rank; g_cfg.step = step;
}
}  // namespace
namespace Kokkos::shim {
RngState& rng() { return t_rng; }
int n_threads() { return g_threads; }
void set_threads(int n) { g_threads = n < 1 ? 1 : n; }
void kernel_begin(const std::string& label) {
  if (g_profile) { g_label = label; g_t0 = std::chrono::steady_clock::now(); }
  g_init_kernel = label == "mc_init_first";  // InitFunctor (mc/src/unit.cpp:102-144): one stream per visited particle
  if (label == "cycle_move") g_cfg.mode = Mode::MoveTape;
  else if (label == "cycle_move_leave") g_cfg.mode = Mode::Leave;
  else g_cfg.mode = Mode::Sequential;
}
void kernel_end() {
  g_cfg.mode = Mode::Sequential; g_init_kernel = false;
  if (g_profile) g_times[g_label] += std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count();
}
void team_begin(size_t league_rank) { RngState& r = synced(); r.league = league_rank; r.tape = 0; }
void range_index(size_t i) {
  RngState& r = synced(); r.index = i;
  if (g_init_kernel) r.start_sequence((uint32_t)i, 2u);  // M::init, then KPRNG::uniform_u, continue this stream
}
}  // namespace Kokkos::shim

// oracle/ref_unit.cpp: streams of the reference's own MC::init (step id 0xFFFFFFFF is reserved for initialisation)
void ref_unit_stream_setup(uint64_t seed, uint32_t rank) { set_streams(seed, rank, 0xFFFFFFFFu); }

// ------------------------------------------------------------------ model wrappers
namespace {
// draw blocks of the model hooks' generator: update/init 3.., division 0x40000001..
constexpr uint32_t kBaseUpdate = 2u, kBaseDivision = 0x40000000u;

template <class M> struct Tap : M {
  using Self = Tap;
  KOKKOS_INLINE_FUNCTION static MC::Status update(const MC::pool_type& pool, typename M::FloatType d_t, std::size_t idx,
                                                  const typename M::SelfParticle& arr, const typename M::SelfContribs& contribs,
                                                  std::size_t position, const MC::LocalConcentration& c) {
    synced().start_sequence((uint32_t)idx, kBaseUpdate);
    return M::update(pool, d_t, idx, arr, contribs, position, c);
  }
  KOKKOS_INLINE_FUNCTION static void division(const MC::pool_type& pool, std::size_t idx, std::size_t idx2,
                                              const typename M::SelfParticle& arr, const typename M::SelfParticle& buf) {
    synced().start_sequence((uint32_t)idx, kBaseDivision);
    M::division(pool, idx, idx2, arr, buf);
  }
};

struct MonodQ1 {
  using Base = Models::Monod;
  using Self = MonodQ1;
  using FloatType = Base::FloatType;
  using uniform_weight = std::true_type;
  using Config = std::nullopt_t;
  static constexpr std::size_t n_var = Base::n_var;
  static constexpr std::size_t n_c = 1;
  using SelfParticle = Base::SelfParticle;
  using SelfContribs = MC::ParticlesContribs<n_c, FloatType>;
  KOKKOS_INLINE_FUNCTION static void init(const MC::pool_type& pool, std::size_t idx, const SelfParticle& arr) { Base::init(pool, idx, arr); }
  KOKKOS_INLINE_FUNCTION static double mass(std::size_t idx, const SelfParticle& arr) { return Base::mass(idx, arr); }
  KOKKOS_INLINE_FUNCTION static MC::Status update(const MC::pool_type& pool, FloatType d_t, std::size_t idx, const SelfParticle& arr,
                                                  const SelfContribs& contribs, std::size_t position, const MC::LocalConcentration& c) {
    const MC::Status s = Base::update(pool, d_t, idx, arr, position, c);
    contribs(idx, 0) = arr(idx, INDEX_FROM_ENUM(Base::particle_var::phi_s_c));
    return s;
  }
  KOKKOS_INLINE_FUNCTION static void division(const MC::pool_type& pool, std::size_t idx, std::size_t idx2, const SelfParticle& arr,
                                              const SelfParticle& buf) { Base::division(pool, idx, idx2, arr, buf); }
  static std::vector<std::string_view> names() { return Base::names(); }        // partial export: length, mu, mu_eff
  static std::vector<std::size_t> get_number() { return Base::get_number(); }
};
static_assert(ModelType<MonodQ1>);
static_assert(ModelType<Tap<Models::FixedLength>>);
static_assert(ModelType<Tap<Models::SimpleAcetate>>);
static_assert(ModelType<Tap<MonodQ1>>);
static_assert(ConstWeightModelType<Tap<MonodQ1>>);
#ifdef BMC_REF_WITH_UDF
static_assert(ModelType<Tap<Models::UdfModel>>);
#endif

// ------------------------------------------------------------------ one simulation unit's worth of state
struct IRef {
  virtual ~IRef() = default;
  std::string err;
  uint64_t seed = 2024; uint32_t rank = 0, step = 0;
  size_t n_species = 1, n_comp = 1;
  MC::RuntimeParameters rt{0, 0.6, 1.5, 0.0, 0.01};
  KernelDispatchOptions opts{};
  // domain inputs (applied lazily: init_inner reallocates every view)
  std::vector<double> vol, out_flows, proba; std::vector<size_t> neigh; size_t m = 1;
  std::vector<size_t> lf_index; std::vector<double> lf_flow, lf_vol;
  bool domain_dirty = true;
  unsigned long long last_out = 0, last_dead = 0, last_waiting = 0, total_out = 0, total_new = 0, n_compactions = 0;
  virtual int n_var() const = 0;
  virtual int n_c() const = 0;
  virtual void set_particles(size_t n, const float* props, const uint64_t* pos, const uint8_t* st, const float* ah, const float* ad) = 0;
  virtual void get_particles(size_t n, float* props, uint64_t* pos, uint8_t* st, float* ah, float* ad) = 0;
  virtual void get_contribs(size_t n, float* out) = 0;
  virtual void get_properties(double* pv, double* sv, double* ages, uint64_t* n_p, uint64_t* n_rows) = 0;
  virtual double init_particles(size_t n, bool uniform_pos, const float* linit) = 0;
  virtual void set_weight(double w) = 0;
  virtual void set_conc(const double* c) = 0;
  virtual void cycle(double dt) = 0;
  virtual void get_sources(double* out) = 0;
  virtual void counters(unsigned long long* c) = 0;
};

template <class M> struct Ref final : IRef {
  using Container = MC::ParticlesContainer<M>;
  using Functors = Simulation::KernelInline::CycleFunctors<ComputeSpace, M>;
  Container container;
  MC::ReactorDomain domain;
  MC::EventContainer events;
  MC::pool_type pool;
  Kokkos::View<double**, Kokkos::LayoutLeft, ComputeSpace> conc;  // (n_species, n_comp): scalar_simulation.hpp:116-119
  MC::kernelContribution contribs;                                  // float (n_species, n_comp)
  MC::ContributionView scatter;
  Simulation::ProbeAutogeneratedBuffer probe_leave, probe_div;
  std::unique_ptr<Functors> functors;
  double weight = 1.0;

  int n_var() const override { return (int)M::n_var; }
  int n_c() const override { return (int)M::n_c; }

  void set_particles(size_t n, const float* props, const uint64_t* pos, const uint8_t* st, const float* ah, const float* ad) override {
    container = Container(rt, n, 0);  // particles_container.hpp:670-697
    for (size_t i = 0; i < n; ++i) {
      for (size_t k = 0; k < M::n_var; ++k) container.model(i, k) = props[k * n + i];
      container.position(i) = pos ? pos[i] : 0;
      container.status(i) = st ? static_cast<MC::Status>(st[i]) : MC::Status::Idle;
      container.ages(i, 0) = ah ? ah[i] : 0.f;
      container.ages(i, 1) = ad ? ad[i] : 0.f;
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
  // MC::init<M> + impl_init + initialize_model (mc/public/mc/mcinit.hpp:67-105, mc/src/unit.cpp:146-163,258-300).
  // InitFunctor itself (unit.cpp:102-144) sits in an anonymous namespace of a translation unit that also pulls in
  // the generated model loader, so its three statements are issued here over the reference's own Model::init,
  // MC::KPRNG::uniform_u and Model::mass; step id 0xFFFFFFFF is the stream reserved for initialisation.
  double init_particles(size_t n, bool uniform_pos, const float* linit) override {
    container = Container(rt, n, 0);
    functors.reset();
    set_streams(seed, rank, 0xFFFFFFFFu);
    MC::KPRNG kprng(seed ? seed : 1);
    const uint64_t min_c = 0, max_c = uniform_pos ? n_comp : 1;  // select_bounds_compartment
    Kokkos::View<float*, ComputeSpace> lin("linit", n);
    Kokkos::View<float**> lin2("linit2", n, 1);
    for (size_t i = 0; i < n; ++i) lin(i) = lin2(i, 0) = linit ? linit[i] : 1.5e-6f;
    double total_mass = 0.;
    for (size_t i = 0; i < n; ++i) {
      synced().start_sequence((uint32_t)i, kBaseUpdate);
      if constexpr (std::is_same_v<typename M::Config, Kokkos::View<float**>>) M::init(kprng.random_pool, i, container.model, lin2);  // UdfModel: config(idx, 0)
      else if constexpr (ConfigurableModel<M>) M::init(kprng.random_pool, i, container.model, typename M::Config(lin));
      else M::init(kprng.random_pool, i, container.model);
      container.position(i) = kprng.uniform_u(min_c, max_c);
      total_mass += M::mass(i, container.model);
    }
    container.weights(0) = (typename M::FloatType)weight;
    return total_mass;
  }
  // PostProcessing::get_properties (post_process.hpp:173-250): force_remove_dead, then exported properties + mass per
  // particle (rows x n_p), their per-compartment sums (rows x n_comp) and both ages (2 x n_p), all LayoutRight doubles
  void get_properties(double* pv, double* sv, double* ages, uint64_t* n_p, uint64_t* n_rows) override {
    auto res = PostProcessing::get_properties<M>(container, n_comp, true);
    if (!res.has_value() || !res->particle_values.has_value()) throw std::runtime_error("model has no export properties");
    const auto& P = *res->particle_values; const auto& S = *res->spatial_values; const auto& A = *res->ages;
    if (n_p) *n_p = P.extent(1);
    if (n_rows) *n_rows = P.extent(0);
    if (pv) for (size_t r = 0; r < P.extent(0); ++r) for (size_t i = 0; i < P.extent(1); ++i) pv[r * P.extent(1) + i] = P(r, i);
    if (sv) for (size_t r = 0; r < S.extent(0); ++r) for (size_t j = 0; j < S.extent(1); ++j) sv[r * S.extent(1) + j] = S(r, j);
    if (ages) for (size_t r = 0; r < 2; ++r) for (size_t i = 0; i < A.extent(1); ++i) ages[r * A.extent(1) + i] = A(r, i);
  }
  void get_contribs(size_t n, float* out) override {
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < M::n_c; ++j) out[j * n + i] = container.contribs(i, j);
  }
  void set_weight(double w) override { weight = w; if (container.weights.extent(0)) container.weights(0) = (typename M::FloatType)w; }
  void set_conc(const double* c) override {
    if (conc.extent(0) != n_species || conc.extent(1) != n_comp) { conc = decltype(conc)("conc", n_species, n_comp); functors.reset(); }
    for (size_t j = 0; j < n_comp; ++j) for (size_t s = 0; s < n_species; ++s) conc(s, j) = c[s + n_species * j];
  }
  void apply_domain() {
    if (!domain_dirty) return;
    domain = MC::ReactorDomain(std::span<double>(vol));          // mcinit.hpp:83
    domain.init_inner(lf_index.size());                          // simulation.cpp:74
    domain.update(vol, neigh, out_flows, proba);                 // simulation.cpp:112-120
    for (size_t i = 0; i < lf_index.size(); ++i)                 // simulation.model.cpp:101-110 (update_feed)
      if (lf_flow[i] > 0.) domain.set_leaving_flow(i, lf_index[i], lf_flow[i], lf_vol[i]);
    domain_dirty = false;
    functors.reset();
  }

  // SimulationUnit::cycleProcess + post_cycle (simulation/simulation.hpp:183-239), call for call
  void cycle(double d_t) override {
    const size_t n_particle = container.n_particles();
    if (n_particle == 0) return;
    apply_domain();
    if (conc.extent(0) != n_species) throw std::runtime_error("concentrations not set");
    if (contribs.extent(0) != n_species || contribs.extent(1) != n_comp) {
      contribs = MC::kernelContribution("contribs", n_species, n_comp);
      scatter = Kokkos::Experimental::create_scatter_view(contribs);  // simulation.cpp:76-77
      functors.reset();
    }
    set_streams(seed, rank, step);
    g_cfg.per_team = opts.m_p_p_team_move;
    if (!functors)  // init_functors (simulation.hpp:165-181)
      functors = std::make_unique<Functors>(opts, container, pool, MC::KernelConcentrationType(conc), scatter, events,
                                            domain.get_const_inner(), probe_leave, probe_div);
    functors->update(d_t, container, domain.get_const_inner());  // pre_cycle (simulation.hpp:154-162)
    // the concentrations view is captured by value in the cycle functor: same allocation, refreshed in place above
    scatter.reset();                                             // :201
    functors->launch_model(n_particle);                          // :202
    if (functors->move_kernel.need_launch()) functors->launch_move(n_particle);  // :205-208
    // post_cycle
    Kokkos::fence();
    Kokkos::deep_copy(contribs, 0.f);  // "scatter_contribute is called ... when contribs is empty" (simulation.cpp:147-150)
    Kokkos::Experimental::contribute(contribs, scatter);         // scatter_contribute, simulation.cpp:143-151
    const auto [host_red, host_out_counter] = functors->get_host_reduction();
    const size_t before = container.n_particles(), inactive_before = container.get_inactive();
    container.update_and_remove_inactive(host_out_counter, host_red.dead_total);
    if (container.n_particles() != before || (container.get_inactive() == 0 && inactive_before + host_out_counter > 0)) ++n_compactions;
    const size_t pre_merge = container.n_particles();
    container.merge_buffer();
    total_new += container.n_particles() - pre_merge;
    last_out = host_out_counter; last_dead = host_red.dead_total; last_waiting = host_red.waiting_allocation_particle;
    total_out += host_out_counter;
    ++step;
  }
  void get_sources(double* out) override {
    for (size_t j = 0; j < n_comp; ++j) for (size_t s = 0; s < n_species; ++s)
      out[s + n_species * j] = contribs.extent(0) ? (double)contribs(s, j) : 0.0;
  }
  void counters(unsigned long long* c) override {
    const auto ev = events.get_span();
    for (int k = 0; k < 6; ++k) c[k] = ev[k];
    c[6] = container.n_particles(); c[7] = container.get_inactive(); c[8] = last_out; c[9] = last_dead; c[10] = last_waiting;
    c[11] = 0; c[12] = container.capacity(); c[13] = total_out; c[14] = total_new; c[15] = n_compactions;
  }
};
}  // namespace

#define REF_TRY(h, ...) try { __VA_ARGS__; return 0; } catch (const std::exception& e) { (h)->err = e.what(); return -1; }

extern "C" {
// model ids as in oracle/oracle.py: 0 fixed_length, 1 monod, 2 simple_acetate, 4 udf_model (checker build only)
void* ref_create(int model, uint64_t n_species, uint64_t n_comp, uint64_t seed, uint32_t rank, uint64_t particles_per_team) {
  IRef* r = nullptr;
  try {
    if (model == 0) r = new Ref<Tap<Models::FixedLength>>();
    else if (model == 1) r = new Ref<Tap<MonodQ1>>();
    else if (model == 2) r = new Ref<Tap<Models::SimpleAcetate>>();
#ifdef BMC_REF_WITH_UDF
    else if (model == 4) { Models::UdfModel::set_nvar(); r = new Ref<Tap<Models::UdfModel>>(); }  // udfmodel_user.cpp:51-56
#endif
    else return nullptr;
  } catch (...) { return nullptr; }
  r->n_species = n_species; r->n_comp = n_comp; r->seed = seed; r->rank = rank;
  if (particles_per_team) {  // KernelDispatchOptions (execinfo.hpp:11-17); default = the generated 1024
    r->opts.m_p_p_team_model = r->opts.m_p_p_team_move = r->opts.m_p_p_team_contribs = particles_per_team;
  }
  r->vol.assign(n_comp, 1.0); r->out_flows.assign(n_comp, 0.0); r->neigh.assign(n_comp, 0); r->proba.assign(n_comp, 1.0);
  for (size_t i = 0; i < n_comp; ++i) r->neigh[i] = i;
  return r;
}
void ref_destroy(void* h) {
  if (g_profile) { for (auto& [k, v] : g_times) std::fprintf(stderr, "[ref] %-28s %.3f s\n", k.c_str(), v); g_times.clear(); }
  delete static_cast<IRef*>(h);
}
const char* ref_last_error(void* h) { return static_cast<IRef*>(h)->err.c_str(); }
int ref_n_var(void* h) { return static_cast<IRef*>(h)->n_var(); }
int ref_n_c(void* h) { return static_cast<IRef*>(h)->n_c(); }
void ref_set_runtime(void* h, uint64_t min_removal, double buffer_ratio, double allocation_factor, double shrink_ratio, double dead_ratio) {
  static_cast<IRef*>(h)->rt = MC::RuntimeParameters{min_removal, buffer_ratio, allocation_factor, shrink_ratio, dead_ratio};
}
// number of OpenMP threads the shim runs leagues / ranges on (1 = serial and deterministic, the default)
void ref_set_threads(int n) { Kokkos::shim::set_threads(n); }
int ref_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ref_set_step(void* h, uint32_t s) { static_cast<IRef*>(h)->step = s; }
int ref_set_particles(void* h, uint64_t n, const float* props, const uint64_t* pos, const uint8_t* st, const float* ah, const float* ad) {
  auto* r = static_cast<IRef*>(h); REF_TRY(r, r->set_particles(n, props, pos, st, ah, ad));
}
int ref_get_particles(void* h, uint64_t n, float* props, uint64_t* pos, uint8_t* st, float* ah, float* ad) {
  auto* r = static_cast<IRef*>(h); REF_TRY(r, r->get_particles(n, props, pos, st, ah, ad));
}
int ref_init_particles(void* h, uint64_t n, int uniform_pos, const float* linit, double* total_mass) {
  auto* r = static_cast<IRef*>(h);
  REF_TRY(r, { const double m = r->init_particles(n, uniform_pos != 0, linit); if (total_mass) *total_mass = m; });
}
// distributions of mc/prng/prng_extension.hpp on the streams orc_sample uses: counter {i, i>>32, 3.., 0}, key = seed
int ref_sample(int kind, uint64_t seed, uint64_t n, double p0, double p1, double p2, double p3, double* out) {
  using namespace MC::Distributions;
  MC::pool_type pool;
  for (uint64_t i = 0; i < n; ++i) {
    set_streams(seed, 0, (uint32_t)(i >> 32));
    synced().start_sequence((uint32_t)i, kBaseUpdate);
    auto gen = pool.get_state();
    switch (kind) {
      case 0: out[i] = gen.normal(p0, p1); break;
      case 1: out[i] = LogNormal<double>{p0, p1}.draw(gen); break;
      case 2: out[i] = TruncatedNormal<double>(p0, p1, p2, p3).draw(gen); break;
      case 3: out[i] = (double)TruncatedNormal<float>((float)p0, (float)p1, (float)p2, (float)p3).draw(gen); break;
      case 4: out[i] = (double)Exponential<float>{(float)p0}.draw(gen); break;
      case 5: out[i] = gen.drand(); break;
      case 6: out[i] = (double)gen.frand(); break;
      case 7: out[i] = norminv<double>(gen.drand(), p0, p1); break;
      default: return -1;
    }
    pool.free_state(gen);
  }
  return 0;
}
int ref_get_properties(void* h, double* pv, double* sv, double* ages, uint64_t* n_p, uint64_t* n_rows) {
  auto* r = static_cast<IRef*>(h); REF_TRY(r, r->get_properties(pv, sv, ages, n_p, n_rows));
}
int ref_get_contribs(void* h, uint64_t n, float* out) { auto* r = static_cast<IRef*>(h); REF_TRY(r, r->get_contribs(n, out)); }
void ref_set_weight(void* h, double w) { static_cast<IRef*>(h)->set_weight(w); }
int ref_domain_update(void* h, const double* vol, const uint64_t* neigh, const double* out_flows, const double* proba, uint64_t m) {
  auto* r = static_cast<IRef*>(h);
  REF_TRY(r, {
    r->vol.assign(vol, vol + r->n_comp); r->out_flows.assign(out_flows, out_flows + r->n_comp);
    r->neigh.assign(neigh, neigh + r->n_comp * m); r->proba.assign(proba, proba + r->n_comp * m); r->m = m; r->domain_dirty = true;
  });
}
int ref_set_leaving_flows(void* h, uint64_t k, const uint64_t* idx, const double* flow, const double* vol) {
  auto* r = static_cast<IRef*>(h);
  REF_TRY(r, { r->lf_index.assign(idx, idx + k); r->lf_flow.assign(flow, flow + k); r->lf_vol.assign(vol, vol + k); r->domain_dirty = true; });
}
int ref_set_concentrations(void* h, const double* c) { auto* r = static_cast<IRef*>(h); REF_TRY(r, r->set_conc(c)); }
int ref_cycle(void* h, double dt) { auto* r = static_cast<IRef*>(h); REF_TRY(r, r->cycle(dt)); }
int ref_get_sources(void* h, double* out) { auto* r = static_cast<IRef*>(h); REF_TRY(r, r->get_sources(out)); }
int ref_get_counters(void* h, unsigned long long* c) { auto* r = static_cast<IRef*>(h); REF_TRY(r, r->counters(c)); }
}
