// TEST INFRASTRUCTURE (oracle/): stand-in for the cxx bridge header of rcmtool (`lib.rs.h`, generated from un-vendored
// Rust sources at configure time in the reference's build).  Only what apps/libs/cma_utils/public/cma_utils/alias.hpp,
// apps/libs/simulation/src/implScalar.cpp, src/simulation.cpp (updateHydro) and src/hydro/impl_mtr.cpp touch:
// rust::Box, the COO matrix wrapper handed to ScalarSimulation::set_transition, and an iteration state that is a plain
// container of the arrays the reference reads from it.
#pragma once
#include <cstddef>
#include <memory>
#include <span>
#include <vector>

namespace rust {
template <class T> class Box {
  std::unique_ptr<T> p_;
 public:
  explicit Box(T* p) : p_(p) {}
  Box(Box&&) noexcept = default;
  Box& operator=(Box&&) noexcept = default;
  T* operator->() const { return p_.get(); }
  T& operator*() const { return *p_; }
};
}  // namespace rust

struct TransitionerWrapper;
struct CooMatrixWrap {
  std::size_t n = 0;
  std::vector<std::size_t> r, c;
  std::vector<double> v;
  std::size_t nrows() const { return n; }
  std::span<const std::size_t> row_indices() const { return {r.data(), r.size()}; }
  std::span<const std::size_t> col_indices() const { return {c.data(), c.size()}; }
  std::span<const double> values() const { return {v.data(), v.size()}; }
};
// what SimulationUnit::updateHydro (simulation.cpp:97-139) and hydro/impl_mtr.cpp read from an iteration state: per
// phase the volumes, their inverses, the out-flows and the transition matrix; the flattened neighbour and
// cumulative-probability tables of the liquid; the named per-compartment field "energy_dissipation"
struct PhaseStateWrapper {
  std::vector<double> vol, inv_vol, out;
  CooMatrixWrap coo;
  std::span<const double> volume() const { return {vol.data(), vol.size()}; }
  std::span<const double> inverse_volume() const { return {inv_vol.data(), inv_vol.size()}; }
  std::span<const double> out_flows() const { return {out.data(), out.size()}; }
  rust::Box<CooMatrixWrap> transition() const { return rust::Box<CooMatrixWrap>(new CooMatrixWrap(coo)); }
};
struct IterationStateWrapper {
  PhaseStateWrapper liq, gas;
  bool with_gas = false;
  std::vector<std::size_t> neighbors;
  std::vector<double> probability_leaving;
  std::vector<double> energy_dissipation;
  const PhaseStateWrapper* get_liquid() const { return &liq; }
  const PhaseStateWrapper* get_gas() const { return &gas; }
  bool has_gas() const { return with_gas; }
  std::span<const std::size_t> flat_neighobrs() const { return {neighbors.data(), neighbors.size()}; }  // (sic)
  std::span<const double> flat_probability_leaving() const { return {probability_leaving.data(), probability_leaving.size()}; }
  std::span<const double> get_misc(const char*) const { return {energy_dissipation.data(), energy_dissipation.size()}; }
};
