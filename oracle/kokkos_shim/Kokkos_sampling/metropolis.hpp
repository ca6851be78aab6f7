// oracle/kokkos_shim: stand-in for the un-vendored `mh_sampling` wrap (test infrastructure).  The reference samples the
// initial cell lengths of its configurable models with a Metropolis sampler (models/src/config_loader.cpp:34,54,
// apps/udf_model/minimal.cpp:139-140).  That sampler is not in the reference tree; the parity tests hand the lengths in,
// so the stand-in copies them from the array the driver registered (shim::metropolis_source) instead of sampling.
#pragma once
#include <cstddef>
#include <stdexcept>
namespace Kokkos::shim {
struct MetropolisSource { const float* values = nullptr; std::size_t n = 0; };
MetropolisSource& metropolis_source();  // defined by the driver (oracle/ref_unit.cpp)
}  // namespace Kokkos::shim
namespace Sampling {
template <class F, class V, class T> int metropolis(F&&, V&& samples, T, T) {
  const auto& src = Kokkos::shim::metropolis_source();
  if (!src.values || src.n < samples.extent(0)) return 1;  // "Error when sampling"
  for (std::size_t i = 0; i < samples.extent(0); ++i) samples(i) = src.values[i];
  return 0;
}
}  // namespace Sampling
