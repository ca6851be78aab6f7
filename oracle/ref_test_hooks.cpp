// oracle/ref_test_hooks.cpp — TEST INFRASTRUCTURE.  Default definitions of the shim's driver hooks for executables that
// are NOT driven by ref_driver.cpp: the reference's own unit tests, compiled as they are over oracle/kokkos_shim
// (`make -C oracle ref_tests`).  One sequential Philox stream per thread, serial execution.
#include <Kokkos_Core.hpp>
#include <Kokkos_sampling/metropolis.hpp>
namespace Kokkos::shim {
RngState& rng() { thread_local RngState s; return s; }
int n_threads() { return 1; }
void set_threads(int) {}
void kernel_begin(const std::string&) {}
void kernel_end() {}
void team_begin(size_t) {}
void range_index(size_t) {}
MetropolisSource& metropolis_source() { static MetropolisSource s; return s; }
}  // namespace Kokkos::shim
