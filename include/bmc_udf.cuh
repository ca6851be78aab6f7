// include/bmc_udf.cuh — source-level contract of a user-defined model (UDF).
//
// The reference loads a UDF as a host shared object whose hooks are called
// through function pointers over Kokkos views (apps/libs/models/ext/
// udf_includes.hpp:27-122, apps/libs/models/src/udfmodel_user.cpp:28-56) and
// refuses to do so in CUDA builds (meson.build:30-34).  Here the same hooks are
// device code: the model SOURCE named by `-mn udf_model` + BIOMC_LIB_UDF
// (apps/api/src/udf_handle.cpp:22-39) or bmc_config::udf_source_path is compiled
// at bmc_create() with NVRTC for sm_100a, inlined into the fused cycle kernel
// and loaded as a CUDA library.  What is kept is the reference's source-level
// contract (apps/udf_model/minimal.cpp:59-155):
//
//   * hook names, argument order and meaning
//       init    (random_pool, idx, arr, config)
//       update  (random_pool, d_t, idx, arr, arr_contribs, position_index, c) -> MC::Status
//       division(random_pool, idx, idx2, arr, buffer_arr)
//       mass    (idx, arr) -> double
//       set_nvar() / set_nc()          (must be constexpr here: sizes are
//                                       template parameters of the kernel)
//   * the type names the hooks are written against (MC::pool_type, MC::Status,
//     MC::LocalConcentration, MC::DynParticlesModel<float>,
//     Models::UdfModel::{FloatType,SelfParticle,SelfContribs,Config})
//   * the access macros of mc/macros.hpp:18-50 (GET_PROPERTY, GET_PROPERTY_FROM,
//     GET_CONTRIBS, GET_CONCENTRATION, INDEX_FROM_ENUM, COPY_PROPERTY_TO)
//   * the helpers of models/utils.hpp (check_div, c_linear_density, get_phi_s_max)
//   * EXPORT_MODULE(module, &init, &update, &division, &mass, names, get_number,
//                   &set_nvar, &set_nc, get_config[, species]) with the reference's
//     argument positions.  names / get_number / get_config / species are host-side
//     in the reference (std::vector, Kokkos host views): pass BMC_UDF_NONE.
//
// The translation unit is compiled with NVRTC's -default-device, so hooks need no
// __device__ annotation; it has no host headers (<vector>, <cstdio>, Kokkos).
// The random pool is a counter-based Philox generator with the subset of the
// Kokkos generator interface the reference's models use (frand/drand/urand64/
// normal), obtained as in the reference: `auto gen = random_pool.get_state();
// ... random_pool.free_state(gen);`.
//
// Optional hints (extensions, #define before including this header):
//   BMC_UDF_WRITE_ONLY_MASK      bit k = property k is never read before being written, so its
//                                column is not loaded
//   BMC_UDF_ALWAYS_WRITTEN_MASK  bit k = update assigns property k on every call, so the column
//                                is stored without an old/new comparison
#pragma once
#include "bmc_kernels.cuh"

namespace std { using ::size_t; }

namespace MC {
enum class Status : char { Idle = 0, Division, Exit, Dead };  // mc/alias.hpp:124-130

// handle to the generator of the particle in flight (MC::pool_type, mc/alias.hpp:98-102)
struct pool_type {
  bmc::Gen* g;
  using generator_type = bmc::Gen;
  __device__ __forceinline__ bmc::Gen get_state() const { return *g; }
  __device__ __forceinline__ void free_state(const bmc::Gen& s) const { *g = s; }
};
using generator_type = pool_type::generator_type;

// arr(idx, k): the registers of the particle in flight (idx ignored) or rows of the
// division buffer in global memory (MC::DynParticlesModel, mc/alias.hpp:60-85)
template <class F> struct DynParticlesModel {
  F* base;
  size_t stride;   // elements between properties
  size_t idx_mul;  // 0 = register row, 1 = SoA columns indexed by idx
  __device__ __forceinline__ F& operator()(size_t idx, size_t k) const { return base[k * stride + idx * idx_mul]; }
};
template <class F> using DynParticlesContribs = DynParticlesModel<F>;
using LocalConcentration = bmc::ConcView;  // c(species, position), mc/alias.hpp:169-173
}  // namespace MC

namespace Models {
struct UdfModel {  // models/public/models/udf_model.hpp:10-64
  using FloatType = float;
  using SelfParticle = MC::DynParticlesModel<FloatType>;
  using SelfContribs = MC::DynParticlesContribs<FloatType>;
  struct Config {  // Kokkos::View<float**>: config(idx, col); one column at this boundary (bmc_init_particles linit)
    const float* base;
    __device__ __forceinline__ float operator()(size_t idx, size_t) const { return base ? base[idx] : 1.5e-6f; }
  };
};
template <class F> __host__ __device__ constexpr F c_linear_density(F rho, F d) {  // models/utils.hpp:92-97
  return rho * (F)3.14159265358979323846 * d * d / (F)4.0;
}
template <class F> __host__ __device__ constexpr F get_phi_s_max(F density, F dl, F y = (F)0.5) {  // utils.hpp:45-52
  return (dl * density) / y;
}
template <class F> __device__ __forceinline__ MC::Status check_div(F l, F lc) {  // utils.hpp:62-67
  return (l >= lc) ? MC::Status::Division : MC::Status::Idle;
}
}  // namespace Models

// ---- mc/macros.hpp:18-50 -----------------------------------------------------
#define MODEL_CONSTANT static constexpr
#define INDEX_FROM_ENUM(e) static_cast<std::size_t>((e))
#define GET_PROPERTY_FROM_IDX(__index__, __array_name__, __idx__) __array_name__(__index__, __idx__)
#define GET_PROPERTY_FROM(__index__, __array_name__, enum_name) \
  GET_PROPERTY_FROM_IDX(__index__, __array_name__, INDEX_FROM_ENUM(enum_name))
#define GET_PROPERTY(enum_name) GET_PROPERTY_FROM(idx, arr, enum_name)
#define COPY_PROPERTY_TO(enum_name, __index__, __array_name__) \
  GET_PROPERTY_FROM(__index__, __array_name__, enum_name) = GET_PROPERTY(enum_name);
#define GET_CONCENTRATION(__species_index__) c((__species_index__), position_index)
#define GET_CONTRIBS_FROM_IDX(__index__, __array_name__, __idx__) __array_name__(__index__, __idx__)
#define GET_CONTRIBS_FROM(__index__, __array_name__, enum_name) \
  GET_PROPERTY_FROM_IDX(__index__, __array_name__, INDEX_FROM_ENUM(enum_name))
#define GET_CONTRIBS(enum_name) GET_PROPERTY_FROM(idx, arr_contribs, enum_name)

#define BMC_UDF_NONE 0
#ifndef BMC_UDF_WRITE_ONLY_MASK
#define BMC_UDF_WRITE_ONLY_MASK 0u
#endif
#ifndef BMC_UDF_ALWAYS_WRITTEN_MASK
#define BMC_UDF_ALWAYS_WRITTEN_MASK 0u
#endif

// EXPORT_MODULE (apps/libs/dynlib macro used at apps/udf_model/minimal.cpp:146-155):
// binds the free hooks to the static model concept the cycle kernel is instantiated on.
#define EXPORT_MODULE(module, init_f, update_f, division_f, mass_f, names_f, number_f, nvar_f, nc_f, config_f, ...)      \
  namespace bmc_udf_export {                                                                                          \
  constexpr auto f_init = (init_f);                                                                                   \
  constexpr auto f_update = (update_f);                                                                               \
  constexpr auto f_division = (division_f);                                                                           \
  constexpr auto f_mass = (mass_f);                                                                                   \
  struct Model {                                                                                                      \
    static constexpr int n_var = (int)(*(nvar_f))();                                                                  \
    static constexpr int n_c = (int)(*(nc_f))();                                                                      \
    static constexpr int n_pre = 0;                                                                                   \
    static constexpr uint64_t write_only_mask = (BMC_UDF_WRITE_ONLY_MASK);                                            \
    static constexpr uint64_t always_written_mask = (BMC_UDF_ALWAYS_WRITTEN_MASK);                                    \
    using Row = MC::DynParticlesModel<float>;                                                                         \
    template <class A, class Cfg> __device__ static void init(bmc::Gen& g, size_t idx, const A& arr, const Cfg& cfg) { \
      (*f_init)(MC::pool_type{&g}, idx, Row{arr.v, 1, 0}, Models::UdfModel::Config{cfg.base});                      \
    }                                                                                                                 \
    template <class A> __device__ static double mass(size_t idx, const A& arr) {                                      \
      return (*f_mass)(idx, Row{arr.v, 1, 0});                                                                      \
    }                                                                                                                 \
    template <class A, class C, class Conc>                                                                           \
    __device__ static bmc::Status update(bmc::Gen& g, float d_t, size_t idx, const A& arr, const C& arr_contribs,     \
                                         size_t position_index, const Conc& c) {                                      \
      const MC::Status s = (*f_update)(MC::pool_type{&g}, d_t, idx, Row{arr.v, 1, 0}, Row{arr_contribs.v, 1, 0},    \
                                         position_index, c);                                                          \
      return (bmc::Status)(int)s;                                                                                     \
    }                                                                                                                 \
    template <class A, class B>                                                                                       \
    __device__ static void division(bmc::Gen& g, size_t idx, size_t idx2, const A& arr, const B& buffer_arr) {        \
      (*f_division)(MC::pool_type{&g}, idx, idx2, Row{arr.v, 1, 0}, Row{buffer_arr.base, buffer_arr.stride, 1});    \
    }                                                                                                                 \
  };                                                                                                                  \
  }
