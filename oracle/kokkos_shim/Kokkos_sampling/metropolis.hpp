// oracle/kokkos_shim: stand-in for the un-vendored `mh_sampling` wrap (test infrastructure).  Only the example UDF's
// get_config hook uses it (apps/udf_model/minimal.cpp:139-140); the parity tests pass the initial lengths explicitly.
#pragma once
#include <stdexcept>
namespace Sampling {
template <class F, class V, class T> void metropolis(F&&, V&&, T, T) { throw std::runtime_error("mh_sampling is not available in the shim build"); }
}  // namespace Sampling
