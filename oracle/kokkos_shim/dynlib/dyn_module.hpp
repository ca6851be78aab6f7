// oracle/kokkos_shim: stand-in for the un-vendored `dynlib` wrap (test infrastructure).  The reference declares and
// exports its UDF function table with these macros (models/ext/udf_includes.hpp:111-122, apps/udf_model/minimal.cpp:
// 146-155) and loads it with dlopen.  The checker build includes the example UDF's translation unit directly
// (oracle/ref_udf.cpp), so the table is never needed: the macros expand to nothing.
#pragma once
#define DEFINE_MODULE(...)
#define MODULE_ITEM(x)
#define EXPORT_MODULE(...)
