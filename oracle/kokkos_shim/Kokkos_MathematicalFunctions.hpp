// oracle/kokkos_shim: forwards to the single-header serial stand-in (test infrastructure, see Kokkos_Core.hpp)
#pragma once
#include <Kokkos_Core.hpp>
