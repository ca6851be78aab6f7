// TEST INFRASTRUCTURE (oracle/): stand-in for the cxx bridge header of rcmtool (`lib.rs.h`, generated from un-vendored
// Rust sources at configure time in the reference's build).  Only what apps/libs/cma_utils/public/cma_utils/alias.hpp and
// apps/libs/simulation/src/implScalar.cpp touch: rust::Box and the COO matrix wrapper handed to
// ScalarSimulation::set_transition (nrows, row_indices, col_indices, values).
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
// what hydro/impl_mtr.cpp reads from an iteration state: the volumes of both phases and the named per-compartment
// field "energy_dissipation"
struct PhaseStateWrapper {
  std::vector<double> vol;
  std::span<const double> volume() const { return {vol.data(), vol.size()}; }
};
struct IterationStateWrapper {
  PhaseStateWrapper liq, gas;
  std::vector<double> energy_dissipation;
  const PhaseStateWrapper* get_liquid() const { return &liq; }
  const PhaseStateWrapper* get_gas() const { return &gas; }
  std::span<const double> get_misc(const char*) const { return {energy_dissipation.data(), energy_dissipation.size()}; }
};
struct CooMatrixWrap {
  std::size_t n = 0;
  std::vector<std::size_t> r, c;
  std::vector<double> v;
  std::size_t nrows() const { return n; }
  std::span<const std::size_t> row_indices() const { return {r.data(), r.size()}; }
  std::span<const std::size_t> col_indices() const { return {c.data(), c.size()}; }
  std::span<const double> values() const { return {v.data(), v.size()}; }
};
