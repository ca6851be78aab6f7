// =============================================================================
// oracle/ref_unit.cpp — TEST INFRASTRUCTURE, NOT THE PRODUCT (checker build only).
//
// The reference's Monte-Carlo unit initialisation and bookkeeping, run as they are:
//   MC::init<Model>                      mc/public/mc/mcinit.hpp:67-105
//   MC::impl_init / initialize_model / InitFunctor / select_bounds_compartment
//                                        mc/src/unit.cpp:44-57, 102-163, 258-300   (compiled from where it lies)
//   MC::post_init_weight                 mc/src/unit.cpp:232-257
//   MonteCarloUnit::getRepartition / NcellFunctor / n_particle
//                                        mc/src/unit.cpp:70-100, 167-230
//   MC::load_tuning_constant             mc/src/unit.cpp:302-343  (BIOMC_MC_* environment variables)
//   Models::FixedLength::get_config      models/src/config_loader.cpp:12-64 (its Metropolis sampler is stubbed:
//                                        kokkos_shim/Kokkos_sampling/metropolis.hpp hands back the given lengths)
// Random streams: kernel "mc_init_first" visits particle i with the generator restarted on the stream reserved for
// initialisation (counter {i, 0xFFFFFFFF, 3.., rank}), which M::init and KPRNG::uniform_u then consume in order.
// =============================================================================
#include <Kokkos_Core.hpp>
#include <Kokkos_sampling/metropolis.hpp>
#include <load_balancing/impl_lb.hpp>  // UniformLoadBalancer (apps/core/src/load_balancing/impl_lb.cpp, iload_balancer.cpp:26-49)
#include <mc/mcinit.hpp>
#include <models/fixed_length.hpp>
#include <models/simple_acetate.hpp>

#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace Kokkos::shim {
MetropolisSource& metropolis_source() { static MetropolisSource s; return s; }
}  // namespace Kokkos::shim

namespace {
struct UnitHandle {
  std::unique_ptr<MC::MonteCarloUnit> unit;
  std::vector<double> volumes;
  double total_mass = 0.;
  std::string err;
};
template <class M> void fetch(MC::MonteCarloUnit& u, uint64_t n, float* props, uint64_t* pos, float* weight) {
  auto& c = std::get<MC::ParticlesContainer<M>>(u.container);
  for (uint64_t i = 0; i < n; ++i) {
    for (size_t k = 0; k < M::n_var; ++k) props[k * n + i] = c.model(i, k);
    pos[i] = c.position(i);
  }
  *weight = (float)c.weights(0);
}
}  // namespace

void ref_unit_stream_setup(uint64_t seed, uint32_t rank);  // ref_driver.cpp: selects the initialisation streams

extern "C" {
// model: 0 fixed_length (lengths = linit, n values), 2 simple_acetate
void* ref_unit_init(int model, uint64_t n, const double* volumes, uint64_t n_comp, int uniform, uint64_t seed, uint32_t rank,
                    const float* linit, double x0) {
  auto* h = new UnitHandle();
  try {
    h->volumes.assign(volumes, volumes + n_comp);
    std::vector<size_t> neigh(n_comp, 0);
    Kokkos::shim::metropolis_source() = {linit, linit ? (size_t)n : 0};
    ref_unit_stream_setup(seed, rank);
    if (model == 0) h->unit = MC::init<Models::FixedLength>(nullptr, n, 0, std::span<double>(h->volumes), neigh, uniform != 0, h->total_mass);
    else if (model == 2) h->unit = MC::init<Models::SimpleAcetate>(nullptr, n, 0, std::span<double>(h->volumes), neigh, uniform != 0, h->total_mass);
    Kokkos::shim::metropolis_source() = {};
    if (!h->unit) { delete h; return nullptr; }
    MC::post_init_weight(h->unit, x0, h->total_mass);
  } catch (...) { delete h; return nullptr; }
  return h;
}
void ref_unit_destroy(void* p) { delete static_cast<UnitHandle*>(p); }
double ref_unit_total_mass(void* p) { return static_cast<UnitHandle*>(p)->total_mass; }
double ref_unit_weight(void* p) { return static_cast<UnitHandle*>(p)->unit->init_weight; }
uint64_t ref_unit_n_particle(void* p) { return static_cast<UnitHandle*>(p)->unit->n_particle(); }
int ref_unit_get(void* p, int model, uint64_t n, float* props, uint64_t* pos, float* weight) {
  auto* h = static_cast<UnitHandle*>(p);
  try {
    if (model == 0) fetch<Models::FixedLength>(*h->unit, n, props, pos, weight);
    else fetch<Models::SimpleAcetate>(*h->unit, n, props, pos, weight);
  } catch (const std::exception& e) { h->err = e.what(); return -1; }
  return 0;
}
int ref_unit_repartition(void* p, uint64_t* out, uint64_t n_comp) {
  auto* h = static_cast<UnitHandle*>(p);
  try {
    const auto r = h->unit->getRepartition();
    if (r.size() != n_comp) return -2;
    std::memcpy(out, r.data(), n_comp * 8);
  } catch (const std::exception& e) { h->err = e.what(); return -1; }
  return 0;
}
// ILoadBalancer::balance with the uniform strategy: particles of `rank` out of n over n_ranks (global_initaliser.cpp:275-279)
uint64_t ref_uniform_balance(uint32_t n_ranks, uint32_t rank, uint64_t n) {
  UniformLoadBalancer lb(n_ranks);
  return lb.balance(rank, n);
}
// {minimum_dead_particle_removal, buffer_ratio, allocation_factor, shrink_ratio, dead_particle_ratio_threshold}
void ref_load_tuning_constant(double* out5) {
  const MC::RuntimeParameters r = MC::load_tuning_constant();
  out5[0] = (double)r.minimum_dead_particle_removal; out5[1] = r.buffer_ratio; out5[2] = r.allocation_factor;
  out5[3] = r.shrink_ratio; out5[4] = r.dead_particle_ratio_threshold;
}
}
