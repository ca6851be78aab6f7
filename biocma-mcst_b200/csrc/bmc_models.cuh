// Device-side model hooks for the fused cycle kernel.
//
// The hook surface is the reference's ModelType concept
// (apps/libs/mc/public/mc/traits.hpp:66-133), kept at source level:
//   n_var, n_c
//   init    (random_pool, idx, arr[, config])
//   mass    (idx, arr) -> double
//   update  (random_pool, d_t, idx, arr, arr_contribs, position_index, c) -> Status
//   division(random_pool, idx, idx2, arr, buffer_arr)
// with the reference's access idiom arr(idx,k), arr_contribs(idx,k),
// buffer_arr(idx2,k), c(species, position_index) (mc/macros.hpp:18-50).
// What changes is what stands behind the accessors: `arr` and `arr_contribs`
// are the registers of the particle in flight (the kernel loaded the SoA columns
// with 128-bit accesses and writes back only what changed), `buffer_arr` is the
// division buffer in global memory, `c` gathers from the replicated
// concentration table, and `random_pool` is a counter-based Philox generator.
//
// Optional per-model hints (extensions; a model that omits them is still correct):
//   write_only_mask : bit k set => property k is never read by update/division
//                     before being written; the kernel then skips loading that
//                     column (SURVEY.md §8d algorithmic bytes).
//   always_written_mask : bit k set => update assigns property k on every call; the
//                     kernel then stores the column without comparing old and new
//                     values (fewer registers and instructions).  0 is always valid.
//   n_pre + compartment_terms(c, position, out[n_pre]) : sub-expressions of update
//                     that depend on the local concentration only; evaluated once
//                     per compartment per step and handed back through c.term(k).
//
// All arithmetic is compiled with -fmad=false: IEEE single/double without
// contraction, so deterministic models are bit-reproducible against the oracle.
#pragma once
#include "bmc_rng.cuh"

namespace bmc {

enum Status : uint8_t { Idle = 0, Division = 1, Exit = 2, Dead = 3 };  // alias.hpp:124-130

// arr(idx,k) / arr_contribs(idx,k): registers of the particle in flight
struct RegRow {
  float* v;
  __device__ __forceinline__ float& operator()(size_t, int k) const { return v[k]; }
};
// buffer_arr(idx2,k): division buffer, SoA columns in global memory
struct BufRows {
  float* base;
  size_t stride;
  __device__ __forceinline__ float& operator()(size_t idx2, int k) const { return base[(size_t)k * stride + idx2]; }
};
// c(species, position): MC::LocalConcentration, species fastest (alias.hpp:169-173)
struct ConcView {
  const double* base;
  uint32_t n_species;
  const float* pre;  // this particle's compartment_terms (optional model hook), else nullptr
  __device__ __forceinline__ double operator()(size_t s, size_t pos) const {
    return __ldg(base + s + (size_t)n_species * pos);
  }
  __device__ __forceinline__ float term(int k) const { return pre[k]; }
  // a double stored by compartment_terms in words k (low) and k + 1 (high) with put_term_d
  __device__ __forceinline__ double term_d(int k) const { return __hiloint2double(__float_as_int(pre[k + 1]), __float_as_int(pre[k])); }
  __device__ static __forceinline__ void put_term_d(float* out, int k, double v) {
    out[k] = __int_as_float(__double2loint(v)); out[k + 1] = __int_as_float(__double2hiint(v));
  }
};
// Config view of configurable models (fixed_length.hpp:25): config(idx)
struct ConfigView {
  const float* base;
  __device__ __forceinline__ float operator()(size_t idx) const { return base ? base[idx] : 1.5e-6f; }
};

__host__ __device__ constexpr float c_linear_density(float rho, float d) {  // models/utils.hpp:92-97
  return rho * 3.14159265358979323846f * d * d / 4.0f;
}
__host__ __device__ constexpr float get_phi_s_max(float density, float dl, float y = 0.5f) {  // utils.hpp:45-52
  return (dl * density) / y;
}
__device__ __forceinline__ Status check_div(float l, float lc) {  // utils.hpp:62-67
  return (l >= lc) ? Division : Idle;
}

// ---- distributions: apps/libs/mc/public/mc/prng/prng_extension.hpp -----------
template <typename F> __device__ __forceinline__ F erfinv_w(F x) {  // :80-93
  const F a = (F)0.147;
  const F inv_a = (F)(1. / 0.147);
  const F tmp = (F)(2 / (3.14159265358979323846 * 0.147));
  const double ln1mx2 = log((1. - x) * (1. + x));
  const F term1 = (F)(tmp + (0.5 * ln1mx2));
  const F term2 = (F)(inv_a * ln1mx2);
  (void)a;
  return copysign(sqrt(sqrt(term1 * term1 - term2) - term1), x);
}
template <typename F> __device__ __forceinline__ F norminv(F p, F mean, F stddev) {  // :117-128
  const F e = erfinv_w<F>(2 * p - 1);
  const F c = fmin(fmax(e, (F)-5), (F)5);
  return (F)(mean + stddev * 1.41421356237309504880 * c);
}
template <typename F> __device__ __forceinline__ F truncated_normal(Gen& g, F mu, F sigma, F lower, F upper) {  // :373-404
  const F r = (F)g.drand();
  const F zl = fmin(fmax((lower - mu) / sigma, (F)-5e3), (F)0);
  const F zu = fmin(fmax((upper - mu) / sigma, (F)0), (F)5e3);
  const F pl = (F)(0.5 * erfc(-zl / 1.41421356237309504880));
  const F pu = (F)(0.5 * erfc(-zu / 1.41421356237309504880));
  const F p = r * (pu - pl) + pl;
  return norminv<F>(p, mu, sigma);
}
__device__ __forceinline__ double lognormal(Gen& g, double mu, double sigma) { return exp(g.normal(mu, sigma)); }  // :524-533

// =============================================================================
// fixed_length — apps/libs/models/public/models/fixed_length.hpp:19-161
// =============================================================================
struct FixedLength {
  static constexpr int n_var = 2, n_c = 1, n_pre = 1;
  enum particle_var { length = 0, l_max = 1 };
  static constexpr uint64_t write_only_mask = 0u, always_written_mask = 1u << length;
  static constexpr float l_dot_max = (float)(2e-6 / 3600.);
  static constexpr float l_max_m = (float)2e-6;
  static constexpr float k = (float)1e-3;
  static constexpr float d_m = (float)0.6e-6;
  static constexpr float lin_density = c_linear_density(1000.0f, d_m);
  static constexpr float phi_s_max = get_phi_s_max(lin_density, l_dot_max);

  template <class A, class Cfg>
  __device__ static void init(Gen&, size_t idx, const A& arr, const Cfg& config) {  // :109-120
    arr(idx, length) = config(idx);
    arr(idx, l_max) = l_max_m;
  }
  template <class A> __device__ static double mass(size_t idx, const A& arr) { return arr(idx, length) * lin_density; }
  template <class A, class C, class Conc>
  __device__ static Status update(Gen&, float d_t, size_t idx, const A& arr, const C& arr_contribs,
                                  size_t position_index, const Conc& c) {  // :122-142
    float& l = arr(idx, length);
    const float lmax = arr(idx, l_max);
    float& c_phi_s = arr_contribs(idx, 0);
    const float g = c.term(0);  // s / (k + s), hoisted to compartment_terms
    const float phi_s = phi_s_max * g;
    const float ldot = l_dot_max * g;
    l += d_t * ldot;
    c_phi_s = -phi_s;
    return check_div(l, lmax);
  }
  template <class Conc> __device__ static void compartment_terms(const Conc& c, size_t position_index, float* out) {
    const float s = (float)c(0, position_index);  // :133 (no clamp, Q17)
    out[0] = s / (k + s);
  }
  template <class A, class B>
  __device__ static void division(Gen&, size_t idx, size_t idx2, const A& arr, const B& buffer_arr) {  // :144-160
    const float new_current_length = arr(idx, length) / 2.0f;
    arr(idx, length) = new_current_length;
    arr(idx, l_max) = l_max_m;
    buffer_arr(idx2, length) = new_current_length;
    buffer_arr(idx2, l_max) = l_max_m;
  }
};

// =============================================================================
// monod — apps/libs/models/public/models/monod.hpp:26-215, re-expressed on the
// current 7-argument hook API with n_c = 1 (the shipped struct is stale, SURVEY
// Q1): contribs(idx,0) = phi_s_c, which is also kept as property 5.
// =============================================================================
struct Monod {
  static constexpr int n_var = 6, n_c = 1, n_pre = 1;
  enum particle_var { l = 0, l_max, mu_p, mue, cell_lenghtening, phi_s_c };
  static constexpr uint64_t write_only_mask = (1u << mue) | (1u << phi_s_c);
  static constexpr uint64_t always_written_mask = (1u << l) | (1u << mu_p) | write_only_mask;
  static constexpr float y_s_x = 2.0f;
  static constexpr float mu_max = (float)(0.77 / 3600.);
  static constexpr float tau_meta = (float)(1. / mu_max);
  static constexpr float l_max_m = (float)2e-6;
  static constexpr float l_min_m = (float)(l_max_m / 2.);
  static constexpr float k_s = (float)1e-3;
  static constexpr float d_m = (float)0.6e-6;
  static constexpr float lin_density = c_linear_density(1000.0f, d_m);

  template <class A, class Cfg> __device__ static void init(Gen& g, size_t idx, const A& arr, const Cfg&) {  // :80-113
    const float l0 = truncated_normal<float>(g, (float)(l_max_m * 0.75), (float)(l_max_m * 0.75 / 4), l_min_m, l_max_m);
    arr(idx, l) = l0;
    arr(idx, l_max) = l_max_m;
    arr(idx, mu_p) = mu_max;
    arr(idx, mue) = 0.0f;
    arr(idx, phi_s_c) = 0.0f;
    arr(idx, cell_lenghtening) = (float)((l_max_m / 2.) / 0.693147180559945309417232121458176568);
  }
  template <class A> __device__ static double mass(size_t idx, const A& arr) { return arr(idx, l) * lin_density; }
  template <class A, class C, class Conc>
  __device__ static Status update(Gen&, float d_t, size_t idx, const A& arr, const C& arr_contribs,
                                  size_t position_index, const Conc& c) {  // :122-160
    const float mu = c.term(0);  // mu_max * s / (k_s + s), hoisted to compartment_terms
    const float mu_eff = fminf(arr(idx, mu_p), mu);
    arr(idx, l) += d_t * (mu_eff * arr(idx, cell_lenghtening));
    // `d_t * (1.0 / tau_meta) * (mu - mu_p)` promotes to double (:144-146)
    arr(idx, mu_p) = (float)((double)arr(idx, mu_p) +
                             ((double)d_t * (1.0 / (double)tau_meta)) * (double)(mu - arr(idx, mu_p)));
    arr(idx, mue) = mu_eff;
    const float ph = -mu_eff * y_s_x * (float)mass(idx, arr);
    arr(idx, phi_s_c) = ph;
    arr_contribs(idx, 0) = ph;
    return check_div(arr(idx, l), arr(idx, l_max));
  }
  template <class Conc> __device__ static void compartment_terms(const Conc& c, size_t position_index, float* out) {
    const float s = (float)fmax(0., c(0, position_index));  // :130-131 bounded
    out[0] = mu_max * s / (k_s + s);                        // :132 instantaneous mu from Monod
  }
  template <class A, class B>
  __device__ static void division(Gen&, size_t idx, size_t idx2, const A& arr, const B& buffer_arr) {  // :162-192
    const float new_current_length = arr(idx, l) / 2.0f;
    arr(idx, l) = new_current_length;
    buffer_arr(idx2, l) = new_current_length;
    buffer_arr(idx2, l_max) = l_max_m;
    buffer_arr(idx2, mu_p) = arr(idx, mu_p);
    buffer_arr(idx2, cell_lenghtening) = arr(idx, cell_lenghtening);
    // mue / phi_s_c of the newborn are export-only and rewritten by its first update
    buffer_arr(idx2, mue) = 0.0f;
    buffer_arr(idx2, phi_s_c) = 0.0f;
  }
};

// Layout hint for the compartment table in shared memory (cycle_body): a model whose row is 8 words with the LAST one
// unused may ask for word planes instead of 32-byte rows (gathers of 32 random compartments then spread over all
// banks).  Off unless a model specialises it; user models keep the row-major table.
template <class M> struct PlanarTable { static constexpr bool value = false; };

// =============================================================================
// simple_acetate — apps/libs/models/public/models/simple_acetate.hpp:26-248
// =============================================================================
struct SimpleAcetate {
  // compartment terms: 1/(c0 + k_s), 1/(c1 + k_a) as floats, c0 and c1 as doubles (two words each), one word of
  // padding: a row of the compartment table is 8 words, fetched with two 128-bit gathers
  static constexpr int n_var = 9, n_c = 2, n_pre = 7;
  enum particle_var { length = 0, l_max, a_p, a_max, a_e, a_e_s, a_e_a, phi_s, phi_a };
  // a_e is read by division AFTER update wrote it in the same cycle -> still write-only for loading
  static constexpr uint64_t write_only_mask = (1u << a_e) | (1u << a_e_s) | (1u << a_e_a) | (1u << phi_s) | (1u << phi_a);
  static constexpr uint64_t always_written_mask = (1u << length) | write_only_mask;
  static constexpr float a_max_m = (float)(2e-6 / 3600.);
  static constexpr float l_max_m = (float)2e-6;
  static constexpr float l_min_m = (float)(l_max_m / 2.);
  static constexpr float d_m = (float)0.6e-6;
  static constexpr float lin_density = c_linear_density(1000.0f, d_m);
  static constexpr float k_s = (float)1e-3, k_a = (float)1e-4;
  static constexpr float y_s = 2.0f, y_a = 3.0f;

  __device__ static float tn_mean(float mu, float sigma, float lower, float upper) {  // prng_extension.hpp:406-413
    const float alpha = (lower - mu) / sigma, beta = (upper - mu) / sigma;
    const float ca = (float)(0.5 * (1 + erf(alpha / 1.41421356237309504880)));
    const float cb = (float)(0.5 * (1 + erf(beta / 1.41421356237309504880)));
    const float pa = (float)(0.3989422804014327 * exp(-0.5 * alpha * alpha));
    const float pb = (float)(0.3989422804014327 * exp(-0.5 * beta * beta));
    const float Z = cb - ca;
    return mu + sigma * (pa - pb) / Z;
  }
  template <class A, class Cfg> __device__ static void init(Gen& g, size_t idx, const A& arr, const Cfg&) {  // :132-152
    const float mu = (float)(l_max_m * 0.75), sg = (float)(l_max_m / 10.);
    const float lo = (float)(0.7 * l_min_m), hi = (float)(l_max_m * 1.3);
    arr(idx, length) = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, l_max) = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, a_p) = (float)(a_max_m / 2.);
    arr(idx, a_max) = tn_mean(a_max_m, (float)(a_max_m / 2.), (float)(0.5 * a_max_m), (float)(a_max_m * 1.5));
    arr(idx, a_e) = 0; arr(idx, a_e_s) = 0; arr(idx, a_e_a) = 0; arr(idx, phi_s) = 0; arr(idx, phi_a) = 0;
  }
  template <class A> __device__ static double mass(size_t idx, const A& arr) { return arr(idx, length) * lin_density; }
  template <class A, class C, class Conc>
  __device__ static Status update(Gen&, float d_t, size_t idx, const A& arr, const C& arr_contribs,
                                  size_t position_index, const Conc& c) {  // :154-204
    const float adm0 = arr(idx, a_max), adm1 = arr(idx, a_max) / 3;
    // c0, c1 and the two reciprocals depend on the compartment only (two fp64 divisions and two 8-byte gathers per
    // particle otherwise): hoisted to compartment_terms, same expressions, same bits
    const double c0 = c.term_d(2), c1 = c.term_d(4);
    float inv = c.term(0);  // (float)(1. / (c0 + k_s))
    const float D0 = (float)(adm0 * c0 * inv);
    inv = c.term(1);        // (float)(1. / (c1 + k_a))
    const float D1 = (float)(adm1 * c1 * inv);
    arr(idx, a_e) = 0.0f;
    const float U0 = fminf(D0, arr(idx, a_p));
    arr(idx, a_e) += U0;
    const float pa = D0 - arr(idx, a_p);
    const float mask_pa = (float)(pa < 0.0f);
    const float U1 = mask_pa * fminf(D1, -pa) + (1 - mask_pa) * 0.0f;
    arr(idx, a_e) += U1;
    arr(idx, length) += d_t * arr(idx, a_e);
    arr(idx, a_e_s) = U0;
    arr(idx, a_e_a) = U1;
    const float ps = -1 * D0 * lin_density * y_s;
    const float pA = mask_pa * (-U1 * lin_density * y_a) + (1.0f - mask_pa) * (pa * lin_density * y_s / y_a);
    arr(idx, phi_s) = ps;
    arr(idx, phi_a) = pA;
    arr_contribs(idx, 0) = ps;
    arr_contribs(idx, 1) = pA;
    return check_div(arr(idx, length), arr(idx, l_max));
  }
  template <class Conc> __device__ static void compartment_terms(const Conc& c, size_t position_index, float* out) {  // :160-166
    const double c0 = c(0, position_index), c1 = c(1, position_index);
    out[0] = (float)(1. / (c0 + k_s));
    out[1] = (float)(1. / (c1 + k_a));
    Conc::put_term_d(out, 2, c0);
    Conc::put_term_d(out, 4, c1);
    out[6] = 0.0f;
  }
  template <class A, class B>
  __device__ static void division(Gen& g, size_t idx, size_t idx2, const A& arr, const B& buffer_arr) {  // :206-246
    const float new_current_length = arr(idx, length) / 2.0f;
    arr(idx, length) = new_current_length;
    for (int i = length; i < a_e; ++i) buffer_arr(idx2, i) = arr(idx, i);
    const float current_a_e = arr(idx, a_e);
    const double sigma = 0.2;
    const double average = (double)logf(current_a_e) - sigma * sigma / 2;
    const float gen1 = (float)lognormal(g, average, sigma);
    const float gen2 = (float)lognormal(g, average, sigma);
    const float mu = l_max_m, sg = (float)(l_max_m / 10.);
    const float lo = (float)(l_max_m * 0.7), hi = (float)(1.3 * l_max_m);
    const float lmax1 = truncated_normal<float>(g, mu, sg, lo, hi);
    const float lmax2 = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, a_p) = gen1;
    arr(idx, l_max) = lmax1;
    buffer_arr(idx2, a_p) = gen2;
    buffer_arr(idx2, l_max) = lmax2;
    for (int i = a_e; i < n_var; ++i) buffer_arr(idx2, i) = 0.0f;  // export-only, rewritten by first update
  }
};

template <> struct PlanarTable<SimpleAcetate> { static constexpr bool value = true; };  // words 1..6 used, word 7 padding
static_assert(SimpleAcetate::n_pre == 7, "PlanarTable: 8-word rows");

// =============================================================================
// Wide UDF — synthetic multi-metabolite user model (BASELINE.json configs[4]),
// written against the UDF hook surface (apps/udf_model/minimal.cpp:59-155):
// P float properties all read and written every step, n_c = 4 contributions.
// The test oracle carries an independent restatement of the same definition.
// =============================================================================
template <int P> struct WideUdf {
  static constexpr int n_var = P, n_c = 4, n_pre = 4;
  static constexpr uint64_t write_only_mask = 0u;
  static constexpr uint64_t always_written_mask = (P >= 64 ? ~0ull : ((1ull << P) - 1ull)) & ~2ull;  // all but l_max
  enum { length = 0, l_max = 1, first_pool = 2 };
  static constexpr float l_dot_max = (float)(2e-6 / 3600.);
  static constexpr float lin_density = c_linear_density(1000.0f, (float)0.6e-6);
  static constexpr float phi_max = get_phi_s_max(lin_density, l_dot_max);
  template <class A, class Cfg> __device__ static void init(Gen&, size_t idx, const A& arr, const Cfg& config) {
    arr(idx, length) = config(idx);
    arr(idx, l_max) = (float)2e-6;
#pragma unroll
    for (int k = first_pool; k < P; ++k) arr(idx, k) = 0.5f;
  }
  template <class A> __device__ static double mass(size_t idx, const A& arr) { return arr(idx, length) * lin_density; }
  template <class A, class C, class Conc>
  __device__ static Status update(Gen&, float d_t, size_t idx, const A& arr, const C& arr_contribs,
                                  size_t position_index, const Conc& c) {
    float sat[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) sat[j] = c.term(j);  // hoisted to compartment_terms
    float acc = 0.0f;
#pragma unroll
    for (int k = first_pool; k < P; ++k) {
      const float tau = 50.0f + 10.0f * (float)(k & 7);
      float x = arr(idx, k);
      x += d_t * ((sat[k & 3] - x) / tau);
      arr(idx, k) = x;
      acc += x;
    }
    const float act = (P > first_pool) ? acc / (float)(P - first_pool) : 1.0f;
    arr(idx, length) += d_t * (l_dot_max * act);
#pragma unroll
    for (int j = 0; j < 4; ++j) arr_contribs(idx, j) = -phi_max * sat[j] * act * (1.0f / (float)(j + 1));
    return check_div(arr(idx, length), arr(idx, l_max));
  }
  template <class Conc> __device__ static void compartment_terms(const Conc& c, size_t position_index, float* out) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = (float)fmax(0., c((size_t)j % c.n_species, position_index));
      out[j] = s / ((float)1e-3 * (float)(j + 1) + s);
    }
  }
  template <class A, class B>
  __device__ static void division(Gen&, size_t idx, size_t idx2, const A& arr, const B& buffer_arr) {
    const float nl = arr(idx, length) / 2.0f;
    arr(idx, length) = nl;
    buffer_arr(idx2, length) = nl;
    buffer_arr(idx2, l_max) = arr(idx, l_max);
#pragma unroll
    for (int k = first_pool; k < P; ++k) buffer_arr(idx2, k) = arr(idx, k);
  }
};

}  // namespace bmc
