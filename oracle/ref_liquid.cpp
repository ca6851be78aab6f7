// TEST INFRASTRUCTURE (oracle/): C entry points around the reference's OWN liquid / gas scalar solver, compiled from the
// sources where they lie — apps/libs/simulation/src/implScalar.cpp (ScalarSimulation: performStep, performStepGL,
// clearNegs, set_transition, set_mass ...), src/hydro/mass_transfer.cpp (MassTransferModel::gas_liquid_mass_transfer),
// includes/scalar_simulation.hpp, apps/libs/kokkos-eigen/public/kokkos_eigen.hpp — over oracle/eigen_shim (Eigen is a
// system package of the reference's build, absent here), oracle/kokkos_shim and oracle/rust_shim (rcmtool's cxx bridge
// header).  What this file adds is the call sequence of ONE time step as the reference's main loop runs it
// (apps/core/src/host_specific.cpp:281-291, apps/libs/simulation/src/simulation.model.cpp:55-154):
//     [sources hold the Monte-Carlo contributions of the previous cycle: synchro_sources, implScalar.cpp:194-205]
//     update_feed  -> set_scalar_feed: set_feed(species, input, flow * concentration), set_sink(output, flow)
//     ode_step     -> performStep, or gas_liquid_mass_transfer + performStepGL (gas, -1) + performStepGL (liquid, +1) + clearNegs
//     clearContribution -> set_zero_contribs
// SimulationUnit itself cannot be built (rcmtool states, MC unit, exporters), so those three calls are restated here.
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <span>
#include <vector>

#include <common/eigen_diag.hpp>
#include <Eigen/Core>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <hydro/impl_mass_transfer.hpp>
#include <scalar_simulation.hpp>
#include <simulation/mass_transfer.hpp>


namespace {
using Simulation::ScalarSimulation;
struct RefLiquid {
  size_t ns, nc;
  std::shared_ptr<ScalarSimulation> liq, gas;
  std::unique_ptr<Simulation::MassTransfer::MassTransferModel> mt;
};
ScalarSimulation* phase(RefLiquid* h, int gas) { return gas ? h->gas.get() : h->liq.get(); }
}  // namespace

extern "C" {
void* refl_create(uint64_t ns, uint64_t nc, const double* vol) {
  auto* h = new RefLiquid{(size_t)ns, (size_t)nc, nullptr, nullptr, nullptr};
  std::vector<double> v(vol, vol + nc);
  h->liq.reset(Simulation::makeScalarSimulation(nc, ns, std::span<double>(v)));
  return h;
}
void refl_destroy(void* p) { delete static_cast<RefLiquid*>(p); }
// gas phase + mass-transfer model (FixedKla with the given per-species value; the tests then overwrite kla / Henry of the
// proxy coefficient by coefficient to exercise arbitrary fields)
int refl_enable_gas(void* p, const double* gas_vol, const double* kla_per_species) {
  auto* h = static_cast<RefLiquid*>(p);
  std::vector<double> v(gas_vol, gas_vol + h->nc);
  h->gas.reset(Simulation::makeScalarSimulation(h->nc, h->ns, std::span<double>(v)));
  Simulation::MassTransfer::Type::FixedKla k{std::vector<double>(kla_per_species, kla_per_species + h->ns)};
  h->mt = std::make_unique<Simulation::MassTransfer::MassTransferModel>(Simulation::MassTransfer::Type::MtrTypeVariant(k), h->liq, h->gas);
  return 0;
}
int refl_set_kla_henry(void* p, const double* kla, const double* henry) {  // kla: species-fastest (ns x nc), henry: ns
  auto* h = static_cast<RefLiquid*>(p);
  if (!h->mt) return -1;
  auto& px = *h->mt->proxy();
  for (size_t j = 0; j < h->nc; ++j) for (size_t s = 0; s < h->ns; ++s) px.kla((Eigen::Index)s, (Eigen::Index)j) = kla[s + h->ns * j];
  for (size_t s = 0; s < h->ns; ++s) px.Henry((Eigen::Index)s) = henry[s];
  return 0;
}
// MassTransferModel of Type::FlowmapTurbulence: update(state) evaluates the kl / interfacial-area correlations of
// hydro/impl_mtr.cpp:22-149 from the state's volumes and energy dissipation and stores kla row 1 (oxygen) in the proxy
int refl_enable_gas_turbulence(void* p, const double* gas_vol) {
  auto* h = static_cast<RefLiquid*>(p);
  std::vector<double> v(gas_vol, gas_vol + h->nc);
  h->gas.reset(Simulation::makeScalarSimulation(h->nc, h->ns, std::span<double>(v)));
  h->mt = std::make_unique<Simulation::MassTransfer::MassTransferModel>(
      Simulation::MassTransfer::Type::MtrTypeVariant(Simulation::MassTransfer::Type::FlowmapTurbulence{}), h->liq, h->gas);
  return 0;
}
int refl_update_mass_transfer(void* p, const double* liq_vol, const double* gas_vol, const double* eps, double* kla_out) {
  auto* h = static_cast<RefLiquid*>(p);
  if (!h->mt) return -1;
  auto* st = new IterationStateWrapper;
  st->liq.vol.assign(liq_vol, liq_vol + h->nc); st->gas.vol.assign(gas_vol, gas_vol + h->nc); st->energy_dissipation.assign(eps, eps + h->nc);
  CmaUtils::IterationStatePtrType state(st);
  h->mt->update(state);
  const auto& px = *h->mt->proxy();
  for (size_t j = 0; j < h->nc; ++j) for (size_t s = 0; s < h->ns; ++s) kla_out[s + h->ns * j] = px.kla((Eigen::Index)s, (Eigen::Index)j);
  return 0;
}
int refl_get_henry(void* p, double* henry) {  // what the reference's constructor put there (mass_transfer.cpp:116-119)
  auto* h = static_cast<RefLiquid*>(p);
  if (!h->mt) return -1;
  for (size_t s = 0; s < h->ns; ++s) henry[s] = h->mt->proxy()->Henry((Eigen::Index)s);
  return 0;
}
// updateScalarHydro (simulation.cpp:119-138): setVolumes + set_transition
int refl_set_hydro(void* p, int gas, const double* vol, const double* inv_vol, uint64_t nnz, const uint64_t* rows, const uint64_t* cols,
                   const double* vals) {
  auto* h = static_cast<RefLiquid*>(p);
  ScalarSimulation* s = phase(h, gas);
  if (!s) return -1;
  s->setVolumes(std::span<const double>(vol, h->nc), std::span<const double>(inv_vol, h->nc));
  auto* coo = new CooMatrixWrap;
  coo->n = h->nc;
  coo->r.assign(rows, rows + nnz); coo->c.assign(cols, cols + nnz); coo->v.assign(vals, vals + nnz);
  s->set_transition(CmaUtils::StateCooMatrixType(coo));
  return 0;
}
// post_init_concentration (simulation.cpp:157-200): deep_copy_concentration + set_mass
int refl_set_concentration(void* p, int gas, const double* c) {
  auto* h = static_cast<RefLiquid*>(p);
  ScalarSimulation* s = phase(h, gas);
  if (!s) return -1;
  if (!s->deep_copy_concentration(std::vector<double>(c, c + h->ns * h->nc))) return -2;
  s->set_mass();
  return 0;
}
int refl_get_concentration(void* p, int gas, double* c) {
  auto* h = static_cast<RefLiquid*>(p);
  ScalarSimulation* s = phase(h, gas);
  if (!s) return -1;
  auto sp = s->getConcentrationData();
  std::memcpy(c, sp.data(), sp.size() * sizeof(double));
  return 0;
}
int refl_get_mtr(void* p, double* out) {
  auto* h = static_cast<RefLiquid*>(p);
  if (!h->mt) return -1;
  auto d = h->mt->mtr_data();
  if (!d) return -2;
  std::memcpy(out, d->data(), d->size() * sizeof(double));
  return 0;
}
// one time step.  mc_sources (species-fastest, may be null): what synchro_sources left in `sources` after the previous
// cycle; feeds: n_feed_values entries {species, input compartment} with value = flow * concentration; sinks: n_sinks
// entries {compartment} with the feed's flow — per phase (index 0 liquid, 1 gas).
int refl_step(void* p, double d_t, const double* mc_sources, const uint64_t* n_feed_values, const uint64_t* const* feed_species,
              const uint64_t* const* feed_input, const double* const* feed_value, const uint64_t* n_sinks, const uint64_t* const* sink_comp,
              const double* const* sink_flow) {
  auto* h = static_cast<RefLiquid*>(p);
  if (mc_sources) {  // the layout of `sources` is LayoutRight (n_species x n_comp): go through the (i, j) accessor
    for (size_t j = 0; j < h->nc; ++j) for (size_t s = 0; s < h->ns; ++s) h->liq->set_feed(s, j, mc_sources[s + h->ns * j]);  // sources(i, j) += v on a zeroed matrix
  }
  for (int g = 0; g < (h->gas ? 2 : 1); ++g) {  // update_feed -> set_scalar_feed (simulation.model.cpp:55-71)
    ScalarSimulation* s = phase(h, g);
    for (uint64_t k = 0; k < n_feed_values[g]; ++k) s->set_feed(feed_species[g][k], feed_input[g][k], feed_value[g][k]);
    for (uint64_t k = 0; k < n_sinks[g]; ++k) s->set_sink(sink_comp[g][k], sink_flow[g][k]);
  }
  if (h->gas) {  // SimulationUnit::ode_step (simulation.model.cpp:131-154)
    h->mt->gas_liquid_mass_transfer();
    const auto& mtr = h->mt->proxy()->mtr;
    h->gas->performStepGL(d_t, mtr, Simulation::MassTransfer::Sign::GasToLiquid);
    h->liq->performStepGL(d_t, mtr, Simulation::MassTransfer::Sign::LiquidToGas);
    h->liq->clearNegs();
  } else {
    h->liq->performStep(d_t);
  }
  // clearContribution (simulation.model.cpp:43-50)
  if (h->gas) h->gas->set_zero_contribs();
  h->liq->set_zero_contribs();
  return 0;
}
}  // extern "C"
