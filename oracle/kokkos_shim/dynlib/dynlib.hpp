// oracle/kokkos_shim: stand-in for the un-vendored `dynlib` wrap (test infrastructure); see dyn_module.hpp
#pragma once
class DynamicLibrary;  // only named by UnsafeUDF::Loader::init_lib, which the checker build never calls
