// =============================================================================
// oracle/bmc_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A from-scratch CPU restatement of BioCMA-MCST's per-timestep Monte-Carlo
// particle loop (`SimulationUnit::cycleProcess`).  It exists to CHECK the
// sm_100a CUDA path and to serve as the reported CPU baseline.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
// legs may load it.  The product (biocma-mcst_b200/) never links, imports or
// calls anything in this directory.
//
// PARITY STATUS: PINNED against outputs of the reference itself run here.  The reference's
// own hot-path sources (CycleFunctors, CycleFunctor, ContributionFunctor, MoveFunctor,
// ParticlesContainer, ReactorDomain, EventContainer, models/*.hpp, prng_extension.hpp) are
// compiled where they lie under /root/reference over oracle/kokkos_shim (a serial stand-in
// for Kokkos, which is not installed) into oracle/_ref/libbmc_ref.so (oracle/ref_driver.cpp,
// `make -C oracle ref`).  tests/test_reference_sources.py compares this restatement with
// it, live and through the committed fixtures tests/golden/ref*.npz: compartment indices,
// statuses, counters, event tallies, float properties, both ages, MC::init and every
// distribution BIT-EXACT; source terms to 2e-5 relative (the reference sums in float).
// What the shim cannot pin (not in the reference tree): the XorShift1024 pool's bit
// streams, Kokkos' parallel execution order, ScatterView summation order, Kokkos::log(float).
// In addition, what the reference's own tests pin for this path is restated in
// tests/test_oracle_reference_invariants.py:
//   * container counts       apps/libs/mc/tests/test_container.cpp:62-148
//   * distribution moments   apps/libs/mc/tests/test_rng_2.cpp:61-107,190-283
//   * CDF-row invariants     apps/libs/cma_utils/tests/test_transport.cpp:42-79
//   * particle balance       apps/core/src/post_process.cpp:92-117
//   * Philox4x32-10 known-answer vectors (Random123 kat_vectors)
//
// Every function cites the reference file:line (relative to /root/reference)
// it follows.  Third-party arithmetic that is not in the reference tree and is
// therefore re-specified here (documented in DESIGN.md):
//   * Kokkos::Random_XorShift1024_Pool  -> counter-based Philox4x32-10 keyed by
//     (seed, rank | slot, step, draw-block); stochastic parity with the
//     reference is distributional only (BASELINE.json north_star).
//   * Kokkos::log(float) (<=1 ulp, backend dependent) -> (float)log((double)x).
//   * Kokkos ScatterView<float> summation order -> fp64 accumulation.
//   * Kokkos parallel_scan final-pass order in the compaction functor -> the
//     order a serial execution of that functor produces.
//
// Arithmetic is IEEE-754 with NO fused contraction (build with
// -ffp-contract=off, never -ffast-math) so that float state is bit-reproducible
// against the CUDA path (compiled with -fmad=false).
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123 v1.14 philox.h).  Replaces
// MC::pool_type = Kokkos::Random_XorShift1024_Pool (apps/libs/mc/public/mc/
// alias.hpp:98-102).  Counter layout shared by oracle and CUDA path:
//   key = { seed_lo, seed_hi }
//   ctr = { index, step, draw_block, rank }
// draw_block 0: index = slot >> 2, word (slot & 3) = u1 of that slot (leave-
//               compartment test); one block serves four neighbouring slots.
// draw_block 1: index = slot >> 2, word (slot & 3) = u3 (outlet test).
// draw_block 2: index = slot, word 0 = u2 (neighbour pick, movers only).
// draw_block >= 3 (index = slot) feeds the model hooks' generator: update/init
// use 3.., division uses 0x40000001...
// ---------------------------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  static constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      const uint64_t p0 = (uint64_t)M0 * c0;
      const uint64_t p1 = (uint64_t)M1 * c2;
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
      const uint32_t n1 = (uint32_t)p1;
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
      const uint32_t n3 = (uint32_t)p0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// uniform float in [0,1): top 24 bits.  Stands in for gen.frand(0.,1.)
// (move_kernel.hpp:242-258, 614-616).
static inline float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// uniform double in [0,1): 53 bits.  Stands in for gen.drand().
static inline double u01d(uint32_t hi, uint32_t lo) {
  const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

// Generator handed to model hooks (the reference hands them the RNG pool,
// traits.hpp:116-118).  Sequential draws out of Philox blocks 1,2,...
struct Gen {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t buf[4];
  int have;
  Gen(uint64_t seed, uint32_t rank, uint32_t slot, uint32_t step) {
    key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
    ctr[0] = slot; ctr[1] = step; ctr[2] = 2; ctr[3] = rank;  // first block drawn is 3
    have = 0;
  }
  inline uint32_t next32() {
    if (have == 0) { ctr[2] += 1; Philox::block(ctr, key, buf); have = 4; }
    return buf[4 - (have--)];
  }
  inline float frand() { return u01f(next32()); }
  inline double drand() { const uint32_t hi = next32(); const uint32_t lo = next32(); return u01d(hi, lo); }
  inline uint64_t urand64(uint64_t lo, uint64_t hi) {  // [lo,hi)
    const uint64_t v = ((uint64_t)next32() << 32) | next32();
    return lo + v % (hi - lo);
  }
  // Kokkos 5.1.1 Kokkos_Random.hpp normal(): Marsaglia polar method on drand().
  inline double normal() {
    double S = 2.0, U = 0.0;
    while (S >= 1.0) {
      U = 2.0 * drand() - 1.0;
      const double V = 2.0 * drand() - 1.0;
      S = U * U + V * V;
    }
    return U * std::sqrt(-2.0 * std::log(S) / S);
  }
  inline double normal(double mu, double sigma) { return mu + sigma * normal(); }
};

// ---------------------------------------------------------------------------
// Distributions — apps/libs/mc/public/mc/prng/prng_extension.hpp
// ---------------------------------------------------------------------------
// erfinv (Winitzki) prng_extension.hpp:80-93
template <typename F> static inline F erfinv_w(F x) {
  const F a = (F)0.147;
  const F inv_a = (F)(1. / a);
  const F tmp = (F)(2 / (M_PI * a));
  const double ln1mx2 = std::log((1. - x) * (1. + x));
  const F term1 = (F)(tmp + (0.5 * ln1mx2));
  const F term2 = (F)(inv_a * ln1mx2);
  return std::copysign(std::sqrt(std::sqrt(term1 * term1 - term2) - term1), x);
}
// norminv prng_extension.hpp:117-128
template <typename F> static inline F norminv(F p, F mean, F stddev) {
  const F lb = -5, ub = 5;
  const F e = erfinv_w<F>(2 * p - 1);
  const F c = std::min(std::max(e, lb), ub);
  return (F)(mean + stddev * 1.41421356237309504880 * c);
}
// TruncatedNormal<F>::draw_from prng_extension.hpp:373-404
template <typename F> static inline F truncated_normal(Gen& g, F mu, F sigma, F lower, F upper) {
  const F r = (F)g.drand();
  const F zl = std::min(std::max((lower - mu) / sigma, (F)-5e3), (F)0);
  const F zu = std::min(std::max((upper - mu) / sigma, (F)0), (F)5e3);
  const F pl = (F)(0.5 * std::erfc(-zl / 1.41421356237309504880));
  const F pu = (F)(0.5 * std::erfc(-zu / 1.41421356237309504880));
  const F p = r * (pu - pl) + pl;
  return norminv<F>(p, mu, sigma);
}
// LogNormal<double>::draw prng_extension.hpp:524-533
static inline double lognormal(Gen& g, double mu, double sigma) { return std::exp(g.normal(mu, sigma)); }
// Exponential<F>::draw prng_extension.hpp:616-622 with _ln = Kokkos::log(float)
static inline float ln_f32(float x) { return (float)std::log((double)x); }

// ---------------------------------------------------------------------------
// Status / events — alias.hpp:124-130, events.hpp:17-26
// ---------------------------------------------------------------------------
enum Status : uint8_t { Idle = 0, Division = 1, Exit = 2, Dead = 3 };
enum Event { NewParticle = 0, EvExit, Move, Death, Overflow, ChangeWeight, N_EVENTS };

// Property / contribution / concentration accessors with the reference's
// calling idiom arr(idx,k), contribs(idx,k), c(species,position)
// (mc/macros.hpp:18-50).  Layout: AoS rows (LayoutRight = Kokkos OpenMP
// default, alias.hpp:52-56).
struct Arr {
  float* base; int n_var;
  inline float& operator()(size_t idx, int k) const { return base[idx * (size_t)n_var + k]; }
};
struct Conc {
  const double* base; size_t n_species;  // species fastest (LayoutLeft, alias.hpp:169-173)
  inline double operator()(size_t s, size_t pos) const { return base[s + n_species * pos]; }
};

// consteval helpers (models/utils.hpp:45-52, 92-97)
static inline float c_linear_density(float rho, float d) { return rho * (float)M_PI * d * d / 4.0f; }
static inline float get_phi_s_max(float density, float dl, float y = 0.5f) { return (dl * density) / y; }
static inline Status check_div(float l, float lc) { return (l >= lc) ? Division : Idle; }

// ---------------------------------------------------------------------------
// Models.  Hook signatures mirror traits.hpp:66-133:
//   init(gen, idx, arr[, config]); mass(idx, arr);
//   update(gen, d_t, idx, arr, contribs, position, c) -> Status;
//   division(gen, idx, idx2, arr, buffer_arr)
// ---------------------------------------------------------------------------
struct FixedLength {  // apps/libs/models/public/models/fixed_length.hpp:19-161
  static constexpr int n_var = 2, n_c = 1;
  enum { length = 0, l_max = 1 };
  static float l_dot_max() { return (float)(2e-6 / 3600.); }
  static float l_max_m() { return (float)2e-6; }
  static float k() { return (float)1e-3; }
  static float lin_density() { return c_linear_density(1000.0f, (float)0.6e-6); }
  static float phi_s_max() { return get_phi_s_max(lin_density(), l_dot_max()); }
  static void init(Gen&, size_t idx, const Arr& arr, float linit) {  // :109-120
    arr(idx, length) = linit; arr(idx, l_max) = l_max_m();
  }
  static double mass(size_t idx, const Arr& arr) { return arr(idx, length) * lin_density(); }  // :81-84
  static Status update(Gen&, float d_t, size_t idx, const Arr& arr, const Arr& contribs, size_t pos,
                       const Conc& c) {  // :122-142
    float& l = arr(idx, length);
    const float lmax = arr(idx, l_max);
    const float s = (float)c(0, pos);
    const float g = s / (k() + s);
    const float phi_s = phi_s_max() * g;
    const float ldot = l_dot_max() * g;
    l += d_t * ldot;
    contribs(idx, 0) = -phi_s;
    return check_div(l, lmax);
  }
  static void division(Gen&, size_t idx, size_t idx2, const Arr& arr, const Arr& buf) {  // :144-160
    const float nl = arr(idx, length) / 2.0f;
    arr(idx, length) = nl; arr(idx, l_max) = l_max_m();
    buf(idx2, length) = nl; buf(idx2, l_max) = l_max_m();
  }
};

struct Monod {  // apps/libs/models/public/models/monod.hpp:26-215, re-expressed on
                // the current 7-argument hook API with n_c = 1:
                // contribs(idx,0) = phi_s_c, phi_s_c also kept as property 5.
  static constexpr int n_var = 6, n_c = 1;
  enum { l = 0, l_max, mu_p, mue, cell_len, phi_s_c };
  static float y_s_x() { return 2.0f; }
  static float mu_max() { return (float)(0.77 / 3600.); }
  static float tau_meta() { return (float)(1. / mu_max()); }
  static float l_max_m() { return (float)2e-6; }
  static float l_min_m() { return (float)(l_max_m() / 2.); }
  static float k_s() { return (float)1e-3; }
  static float lin_density() { return c_linear_density(1000.0f, (float)0.6e-6); }
  static void init(Gen& g, size_t idx, const Arr& arr) {  // :80-113
    const float l0 = truncated_normal<float>(g, (float)(l_max_m() * 0.75), (float)(l_max_m() * 0.75 / 4),
                                             l_min_m(), l_max_m());
    arr(idx, l) = l0; arr(idx, l_max) = l_max_m(); arr(idx, mu_p) = mu_max();
    arr(idx, mue) = 0.0f; arr(idx, phi_s_c) = 0.0f;
    const double dl = l_max_m() / 2.;
    arr(idx, cell_len) = (float)(dl / 0.693147180559945309417232121458176568);
  }
  static double mass(size_t idx, const Arr& arr) { return arr(idx, l) * lin_density(); }  // :115-120
  static Status update(Gen&, float d_t, size_t idx, const Arr& arr, const Arr& contribs, size_t pos,
                       const Conc& c) {  // :122-160
    const float s = (float)std::max(0., c(0, pos));
    const float mu = mu_max() * s / (k_s() + s);
    const float mu_eff = std::min(arr(idx, mu_p), mu);
    arr(idx, l) += d_t * (mu_eff * arr(idx, cell_len));
    // `d_t * (1.0 / tau_meta) * (mu - mu_p)` is evaluated in double (1.0 literal), :144-146
    arr(idx, mu_p) = (float)((double)arr(idx, mu_p) +
                             ((double)d_t * (1.0 / (double)tau_meta())) * (double)(mu - arr(idx, mu_p)));
    arr(idx, mue) = mu_eff;
    const float ph = -mu_eff * y_s_x() * (float)mass(idx, arr);
    arr(idx, phi_s_c) = ph;
    contribs(idx, 0) = ph;
    return check_div(arr(idx, l), arr(idx, l_max));
  }
  static void division(Gen&, size_t idx, size_t idx2, const Arr& arr, const Arr& buf) {  // :162-192
    const float nl = arr(idx, l) / 2.0f;
    arr(idx, l) = nl;
    buf(idx2, l) = nl; buf(idx2, l_max) = l_max_m();
    buf(idx2, mu_p) = arr(idx, mu_p); buf(idx2, cell_len) = arr(idx, cell_len);
    // The reference leaves the export-only columns of the buffer row untouched (zero at
    // allocation, stale after reuse); they are rewritten by the newborn's first update.
    // Pin them to 0 so the state is a function of the inputs only.
    buf(idx2, mue) = 0.0f; buf(idx2, phi_s_c) = 0.0f;
  }
};

struct SimpleAcetate {  // apps/libs/models/public/models/simple_acetate.hpp:26-248
  static constexpr int n_var = 9, n_c = 2;
  enum { length = 0, l_max, a_p, a_max, a_e, a_e_s, a_e_a, phi_s, phi_a };
  static float a_max_m() { return (float)(2e-6 / 3600.); }
  static float l_max_m() { return (float)2e-6; }
  static float l_min_m() { return (float)(l_max_m() / 2.); }
  static float lin_density() { return c_linear_density(1000.0f, (float)0.6e-6); }
  static float k(int i) { return i == 0 ? (float)1e-3 : (float)1e-4; }
  static float y(int i) { return i == 0 ? 2.0f : 3.0f; }
  static void init(Gen& g, size_t idx, const Arr& arr) {  // :132-152 (ld and lm are both l_dist)
    const float mu = (float)(l_max_m() * 0.75), sg = (float)(l_max_m() / 10.);
    const float lo = (float)(0.7 * l_min_m()), hi = (float)(l_max_m() * 1.3);
    arr(idx, length) = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, l_max) = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, a_p) = (float)(a_max_m() / 2.);
    arr(idx, a_max) = tn_mean(a_max_m(), (float)(a_max_m() / 2.), (float)(0.5 * a_max_m()), (float)(a_max_m() * 1.5));
    arr(idx, a_e) = 0; arr(idx, a_e_s) = 0; arr(idx, a_e_a) = 0; arr(idx, phi_s) = 0; arr(idx, phi_a) = 0;
  }
  // TruncatedNormal<float>::mean prng_extension.hpp:406-413 (+ std_normal_pdf/cdf :145-170)
  static float tn_mean(float mu, float sigma, float lower, float upper) {
    const float alpha = (lower - mu) / sigma, beta = (upper - mu) / sigma;
    auto cdf = [](float x) { return (float)(0.5 * (1 + std::erf(x / 1.41421356237309504880))); };
    auto pdf = [](float x) { return (float)(0.3989422804014327 * std::exp(-0.5 * x * x)); };
    const float Z = cdf(beta) - cdf(alpha);
    return mu + sigma * (pdf(alpha) - pdf(beta)) / Z;
  }
  static double mass(size_t idx, const Arr& arr) { return arr(idx, length) * lin_density(); }
  static Status update(Gen&, float d_t, size_t idx, const Arr& arr, const Arr& contribs, size_t pos,
                       const Conc& c) {  // :154-204
    const float adm0 = arr(idx, a_max), adm1 = arr(idx, a_max) / 3;
    const double c0 = c(0, pos), c1 = c(1, pos);
    float inv = (float)(1. / (c0 + k(0)));
    const float D0 = (float)(adm0 * c0 * inv);
    inv = (float)(1. / (c1 + k(1)));
    const float D1 = (float)(adm1 * c1 * inv);
    arr(idx, a_e) = 0.0f;
    const float U0 = std::min(D0, arr(idx, a_p));
    arr(idx, a_e) += U0;
    const float pa = D0 - arr(idx, a_p);
    const float mask_pa = (float)(pa < 0.0f);
    const float U1 = mask_pa * std::min(D1, -pa) + (1 - mask_pa) * 0.0f;
    arr(idx, a_e) += U1;
    arr(idx, length) += d_t * arr(idx, a_e);
    arr(idx, a_e_s) = U0; arr(idx, a_e_a) = U1;
    const float ps = -1 * D0 * lin_density() * y(0);
    const float pA = mask_pa * (-U1 * lin_density() * y(1)) + (1.0f - mask_pa) * (pa * lin_density() * y(0) / y(1));
    arr(idx, phi_s) = ps; arr(idx, phi_a) = pA;
    contribs(idx, 0) = ps; contribs(idx, 1) = pA;
    return check_div(arr(idx, length), arr(idx, l_max));
  }
  static void division(Gen& g, size_t idx, size_t idx2, const Arr& arr, const Arr& buf) {  // :206-246
    const float nl = arr(idx, length) / 2.0f;
    arr(idx, length) = nl;
    for (int i = length; i < a_e; ++i) buf(idx2, i) = arr(idx, i);
    const float cur_a_e = arr(idx, a_e);
    const double sigma = 0.2;
    const double average = std::log(cur_a_e) - sigma * sigma / 2;  // Kokkos::log(float) promoted
    const float gen1 = (float)lognormal(g, average, sigma);
    const float gen2 = (float)lognormal(g, average, sigma);
    const float mu = l_max_m(), sg = (float)(l_max_m() / 10.);
    const float lo = (float)(l_max_m() * 0.7), hi = (float)(1.3 * l_max_m());
    const float lmax1 = truncated_normal<float>(g, mu, sg, lo, hi);
    const float lmax2 = truncated_normal<float>(g, mu, sg, lo, hi);
    arr(idx, a_p) = gen1; arr(idx, l_max) = lmax1;
    buf(idx2, a_p) = gen2; buf(idx2, l_max) = lmax2;
    for (int i = a_e; i < n_var; ++i) buf(idx2, i) = 0.0f;  // export-only columns, see Monod::division
  }
};

// "Wide UDF": synthetic multi-metabolite user model standing in for BASELINE
// configs[4] ("many properties/particle").  Written against the UDF hook
// surface (apps/udf_model/minimal.cpp:59-155): P float properties all read and
// written each step, n_c = 4 contributions.  Not a reference model — the same
// source-level definition is compiled for the device in
// biocma-mcst_b200/csrc/models.cuh; parity = oracle vs CUDA on identical text.
struct WideUdf {
  static int n_var_rt;  // discovered at load time, like set_nvar_udf (udfmodel_user.cpp:51-56)
  static constexpr int n_c = 4;
  enum { length = 0, l_max = 1, first_pool = 2 };
  static float l_dot_max() { return (float)(2e-6 / 3600.); }
  static float lin_density() { return c_linear_density(1000.0f, (float)0.6e-6); }
  static float phi_max() { return get_phi_s_max(lin_density(), l_dot_max()); }
  static void init(Gen&, size_t idx, const Arr& arr, float linit) {
    arr(idx, length) = linit; arr(idx, l_max) = (float)2e-6;
    for (int k = first_pool; k < arr.n_var; ++k) arr(idx, k) = 0.5f;
  }
  static double mass(size_t idx, const Arr& arr) { return arr(idx, length) * lin_density(); }
  static Status update(Gen&, float d_t, size_t idx, const Arr& arr, const Arr& contribs, size_t pos, const Conc& c) {
    const int P = arr.n_var;
    float sat[4];
    for (int j = 0; j < 4; ++j) {
      const float s = (float)std::max(0., c((size_t)j % c.n_species, pos));
      sat[j] = s / ((float)1e-3 * (float)(j + 1) + s);
    }
    float acc = 0.0f;
    for (int k = first_pool; k < P; ++k) {
      const float tau = 50.0f + 10.0f * (float)(k & 7);
      float x = arr(idx, k);
      x += d_t * ((sat[k & 3] - x) / tau);
      arr(idx, k) = x;
      acc += x;
    }
    const float act = (P > first_pool) ? acc / (float)(P - first_pool) : 1.0f;
    arr(idx, length) += d_t * (l_dot_max() * act);
    for (int j = 0; j < 4; ++j) contribs(idx, j) = -phi_max() * sat[j] * act * (1.0f / (float)(j + 1));
    return check_div(arr(idx, length), arr(idx, l_max));
  }
  static void division(Gen&, size_t idx, size_t idx2, const Arr& arr, const Arr& buf) {
    const float nl = arr(idx, length) / 2.0f;
    arr(idx, length) = nl; buf(idx2, length) = nl; buf(idx2, l_max) = arr(idx, l_max);
    for (int k = first_pool; k < arr.n_var; ++k) buf(idx2, k) = arr(idx, k);
  }
};
int WideUdf::n_var_rt = 32;

// The reference's example user-defined model, apps/udf_model/minimal.cpp:24-121 (loaded through
// `-mn udf_model` + BIOMC_LIB_UDF).  The CUDA path compiles its own source-level version of it
// (examples/minimal_udf.cu) with NVRTC; this is the independent CPU restatement.
struct UdfMinimal {
  static constexpr int n_var = 2, n_c = 1;
  enum { length = 0, l_max = 1 };
  static float l_dot_max() { return (float)(2e-6 / 3600.); }
  static float l_max_m() { return (float)2e-6; }
  static float k() { return (float)1e-3; }
  static float lin_density() { return c_linear_density(1000.0f, (float)0.6e-6); }
  static float phi_s_max() { return (l_dot_max() * lin_density()) / 0.5f; }  // :36-37
  static void init(Gen&, size_t idx, const Arr& arr, float linit) {  // :59-67
    arr(idx, length) = linit; arr(idx, l_max) = l_max_m();
  }
  static double mass(size_t idx, const Arr& arr) { return arr(idx, length) * lin_density(); }  // :110-113
  static Status update(Gen&, float d_t, size_t idx, const Arr& arr, const Arr& contribs, size_t pos,
                       const Conc& c) {  // :69-89
    const float s = (float)c(0, pos);
    const float g = s / (k() + s);
    const float phi_s = phi_s_max() * g;
    const float ldot = l_dot_max() * g;
    const float d_length = d_t * ldot;
    // `length += d_length / (1.0 + d_t * ldot)`: the 1.0 literal promotes the quotient and the sum to double (:82)
    arr(idx, length) = (float)((double)arr(idx, length) + (double)d_length / (1.0 + (double)(d_t * ldot)));
    contribs(idx, 0) = -phi_s;
    return check_div(arr(idx, length), arr(idx, l_max));
  }
  static void division(Gen&, size_t idx, size_t idx2, const Arr& arr, const Arr& buf) {  // :91-108
    const float nl = arr(idx, length) / 2.0f;
    buf(idx2, length) = nl; buf(idx2, l_max) = l_max_m();
    arr(idx, length) = nl; arr(idx, l_max) = l_max_m();
  }
};

enum ModelId { M_FIXED_LENGTH = 0, M_MONOD = 1, M_SIMPLE_ACETATE = 2, M_WIDE_UDF = 3, M_UDF_MINIMAL = 4 };

// ---------------------------------------------------------------------------
// LeavingFlow — domain.hpp:17-23
// ---------------------------------------------------------------------------
struct LeavingFlow { uint64_t index; double flow; double volume; };

// RuntimeParameters — particles_container.hpp:25-43; defaults = CPU build
// (apps/autogenerated/meson.build:7-50, biocma_cst_config.hpp.in:34-45)
struct Runtime {
  uint64_t minimum_dead_particle_removal = 0;
  double buffer_ratio = 1.0;
  double allocation_factor = 2.5;
  double shrink_ratio = 0.1;
  double dead_particle_ratio_threshold = 0.01;
};

struct Ctx {
  int model = 0, n_var = 0, n_c = 0;
  size_t n_species = 1, n_comp = 1;
  uint64_t seed = 2024; uint32_t rank = 0; uint32_t step = 0;
  int n_threads = 1;
  bool quirk_contrib_return = false;  // SURVEY.md Q2 (contribution_kernel.hpp:172-178)
  Runtime rt;
  // container (particles_container.hpp:82-88, 222-231)
  std::vector<float> model_v;    // (N_alloc, n_var) AoS
  std::vector<float> contribs;   // (N_alloc, n_c)
  std::vector<uint64_t> position;
  std::vector<uint8_t> status;
  std::vector<float> age_hyd, age_div;  // ages(i,0), ages(i,1)
  float weight = 1.0f;
  std::vector<float> buffer_model; std::vector<uint64_t> buffer_position;
  uint64_t buffer_index = 0, buffer_cap = 0;
  size_t n_allocated = 0; uint64_t n_used = 0; size_t inactive_counter = 0;
  // domain (domain.hpp:28-35)
  size_t n_cols = 0;
  std::vector<uint64_t> neighbors; std::vector<double> cumulative_probability, diag_transition, liquid_volume;
  std::vector<LeavingFlow> leaving_flow;
  // liquid
  std::vector<double> concentrations;  // species fastest
  std::vector<double> sources;         // species fastest, fp64 accumulate
  // counters
  uint64_t events[N_EVENTS] = {0, 0, 0, 0, 0, 0};
  uint64_t last_out = 0, last_dead = 0, last_waiting = 0;
  uint64_t total_out = 0, total_new = 0, n_compactions = 0;
  std::string err;
};

// particles_container.hpp:601-643 `_resize`
static void resize_container(Ctx& c, size_t new_size, bool force) {
  if (new_size > 0 && (new_size > c.n_allocated || force)) {
    const size_t na = (size_t)std::ceil((double)new_size * c.rt.allocation_factor);
    c.n_allocated = na;
    c.position.resize(na, 0); c.model_v.resize(na * c.n_var, 0.0f); c.contribs.resize(na * c.n_c, 0.0f);
    c.status.resize(na, Idle); c.age_hyd.resize(na, 0.0f); c.age_div.resize(na, 0.0f);
  }
}
// particles_container.hpp:669-685 `__allocate_buffer__`
static void allocate_buffer(Ctx& c) {
  const size_t req = (size_t)std::ceil(c.rt.buffer_ratio * (double)c.n_allocated);
  if (c.buffer_cap < req) {
    c.buffer_position.assign(req, 0); c.buffer_model.assign(req * c.n_var, 0.0f);
    c.buffer_cap = req; c.buffer_index = 0;
  }
}

template <class M> struct ModelOps {
  static inline Status update(Gen& g, float dt, size_t i, const Arr& a, const Arr& cb, size_t pos, const Conc& c) {
    return M::update(g, dt, i, a, cb, pos, c);
  }
  static inline void division(Gen& g, size_t i, size_t j, const Arr& a, const Arr& b) { M::division(g, i, j, a, b); }
};

// --- cycle_model: model_kernel.hpp:163-217 (team chunk loop) + :230-268
// (exec_per_particle) + particles_container.hpp:559-573 (handle_division).
// Parallel form: pass A runs update + flags Division per 1024-particle chunk
// (chunking = apps/autogenerated/meson.build:20-23); the division buffer slot
// of mother i is its rank among dividing mothers in ascending i — exactly the
// slot a serial sweep of the reference functor hands out — so the result is
// independent of the thread count.
template <class M> static void cycle_model(Ctx& c, double d_t_in) {
  const float d_t = (float)d_t_in;  // M::FloatType d_t  (model_kernel.hpp:156-161,270; Q18)
  const size_t n = c.n_used;
  const Arr arr{c.model_v.data(), c.n_var};
  const Arr cb{c.contribs.data(), c.n_c};
  const Arr buf{c.buffer_model.data(), c.n_var};
  const Conc conc{c.concentrations.data(), c.n_species};
  const size_t chunk = 1024, n_chunks = (n + chunk - 1) / chunk;
  std::vector<uint32_t> chunk_div(n_chunks + 1, 0);
  std::vector<uint8_t> flag(n, 0);
#pragma omp parallel for schedule(static) num_threads(c.n_threads)
  for (long ch = 0; ch < (long)n_chunks; ++ch) {
    const size_t p0 = ch * chunk, p1 = std::min(n, p0 + chunk);
    uint32_t nd = 0;
    for (size_t i = p0; i < p1; ++i) {
      if (c.status[i] != Idle) continue;
      c.age_div[i] += d_t;
      Gen g(c.seed, c.rank, (uint32_t)i, c.step);
      const Status s = ModelOps<M>::update(g, d_t, i, arr, cb, c.position[i], conc);
      if (s == Division) { flag[i] = 1; ++nd; }
    }
    chunk_div[ch + 1] = nd;
  }
  for (size_t ch = 0; ch < n_chunks; ++ch) chunk_div[ch + 1] += chunk_div[ch];
  const uint64_t n_div = chunk_div[n_chunks];
  const uint64_t base = c.buffer_index;
  uint64_t waiting = 0;
#pragma omp parallel for schedule(static) num_threads(c.n_threads) reduction(+ : waiting)
  for (long ch = 0; ch < (long)n_chunks; ++ch) {
    if (chunk_div[ch + 1] == chunk_div[ch]) continue;
    const size_t p0 = ch * chunk, p1 = std::min(n, p0 + chunk);
    uint64_t j = base + chunk_div[ch];
    for (size_t i = p0; i < p1; ++i) {
      if (!flag[i]) continue;
      if (j < c.buffer_cap) {  // handle_division :562-570
        Gen g(c.seed, c.rank, (uint32_t)i, c.step);
        g.ctr[2] = 0x40000000u;  // division draws: blocks 0x40000001.. (update draws use 3..)
        ModelOps<M>::division(g, i, j, arr, buf);
        c.buffer_position[j] = c.position[i];
        c.age_div[i] = 0.0f;
      } else {
        ++waiting;  // model_kernel.hpp:253-258
      }
      ++j;
    }
  }
  c.buffer_index = std::min<uint64_t>(base + n_div, c.buffer_cap);
  c.events[Overflow] += waiting;
  c.events[NewParticle] += n_div;  // incremented even on overflow (Q6)
  c.last_waiting = waiting;
  c.last_dead = 0;  // dead_total is never incremented (Q3)
}

// --- contributions: contribution_kernel.hpp:144-186 (Tag3D) / :48-102 (Tag0D).
// S(j,pos) += weight * contribs(p,j) over Idle particles at the PRE-move
// position (Q15).  fp64 accumulation (documented deviation Q7/Q19).
static void contributions(Ctx& c) {
  const size_t n = c.n_used, ns = c.n_species, nb = ns * c.n_comp;
  const int nc = c.n_c;
  const double w = (double)c.weight;  // const double weight = get_weight(p)  (:179)
  const int T = std::max(1, c.n_threads);
  // Q2 belongs to the Tag3D functor only; single-compartment cases run Tag0D (kernels.hpp:200-222), whose
  // lambda skips a non-idle particle and stays inside n_particle (contribution_kernel.hpp:70-88)
  const bool quirk = c.quirk_contrib_return && c.n_comp > 1;
  std::vector<double> part((size_t)T * nb, 0.0);
  const size_t chunk = 1024, n_chunks = (n + chunk - 1) / chunk;
#pragma omp parallel num_threads(T)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    double* acc = part.data() + (size_t)t * nb;
#pragma omp for schedule(static)
    for (long ch = 0; ch < (long)n_chunks; ++ch) {
      const size_t p0 = ch * chunk;
      size_t p1 = std::min(n, p0 + chunk);
      // Q2, second half: the reference bounds the 32-particle runs by `i >= n_particle` with i the RUN index, not
      // the particle (:172-175), so the last run reads up to 31 slots past n_used — stale rows left behind by a
      // compaction still count.  Reproduced only in quirk mode (used to compare with the reference sources).
      if (quirk) p1 = std::min<size_t>(c.n_allocated, p0 + ((p1 - p0 + 31) / 32) * 32);
      for (size_t i0 = p0; i0 < p1; i0 += 32) {  // work_per_thread = 32 (:156)
        for (size_t p = i0; p < std::min(p1, i0 + 32); ++p) {
          if (c.status[p] != Idle) { if (quirk) break; else continue; }
          const size_t pos = c.position[p];
          for (int j = 0; j < nc; ++j) acc[(size_t)j + ns * pos] += w * (double)c.contribs[p * nc + j];
        }
      }
    }
  }
  for (int t = 0; t < T; ++t)
    for (size_t k = 0; k < nb; ++k) c.sources[k] += part[(size_t)t * nb + k];
}

// --- __find_next_compartment: move_kernel.hpp:61-103
static inline size_t find_next_compartment(bool do_search, const Ctx& c, size_t ic, double rnd) {
  const int mask = (int)do_search;
  const int max_neighbor = (int)c.n_cols;
  int left = 0, right = mask * (max_neighbor - 1);
  while (left < right) {
    const int mid = (left + right) >> 1;
    const double pm = c.cumulative_probability[ic * c.n_cols + mid];
    const int m = (int)(rnd > pm);
    left = m * (mid + 1) + (1 - m) * left;
    right = m * right + (1 - m) * mid;
  }
  return ic * (1 - mask) + c.neighbors[ic * c.n_cols + left] * mask;
}
// --- probability_leaving<fast_tag>: probability_leaving.hpp:33-46
static inline bool p_leave_fast(float rnd, double volume, double flow, double dt) { return (dt * flow / volume) > rnd; }
// --- probability_leaving<precision_tag>: probability_leaving.hpp:16-30, _ln maths.hpp:26-33
static inline bool p_leave_precise(float rnd, double volume, double flow, double dt) {
  return (dt * flow) > (-ln_f32(rnd) * volume);
}

// --- cycle_move: move_kernel.hpp:209-273 (TagMove) + :392-437 (handle_move).
// Applies to every slot < n_used regardless of status (the functor has no
// status check).  u1 = block 0 of the slot's group of four, u2 = block 2 of the slot.
static uint64_t cycle_move(Ctx& c, double d_t) {
  const size_t n = c.n_used;
  const uint32_t key[2] = {(uint32_t)c.seed, (uint32_t)(c.seed >> 32)};
  uint64_t moved = 0;
#pragma omp parallel for schedule(static) num_threads(c.n_threads) reduction(+ : moved)
  for (long i = 0; i < (long)n; ++i) {
    const uint32_t ctr[4] = {(uint32_t)i >> 2, c.step, 0u, c.rank};
    uint32_t r[4];
    Philox::block(ctr, key, r);
    const float rng1 = u01f(r[i & 3]);
    const size_t ic = c.position[i];
    const bool mask_next = p_leave_fast(rng1, c.liquid_volume[ic], c.diag_transition[ic], d_t);
    float rng2 = 0.0f;
    if (mask_next) {  // the draw is a pure function of (slot, step): skipping it is harmless (Q13)
      const uint32_t ctr2[4] = {(uint32_t)i, c.step, 2u, c.rank};
      Philox::block(ctr2, key, r);
      rng2 = u01f(r[0]);
    }
    c.position[i] = find_next_compartment(mask_next, c, ic, rng2);
    if (mask_next) ++moved;
  }
  c.events[Move] += moved;
  return moved;
}

// --- cycle_move_leave: move_kernel.hpp:347-359 (TagLeave) + :586-648
// (handle_exit) + :105-127 (find_flow).  Uses the POST-move position (Q13).
static uint64_t cycle_leave(Ctx& c, double d_t) {
  const size_t n = c.n_used, nf = c.leaving_flow.size();
  const uint32_t key[2] = {(uint32_t)c.seed, (uint32_t)(c.seed >> 32)};
  uint64_t dead = 0;
#pragma omp parallel for schedule(static) num_threads(c.n_threads) reduction(+ : dead)
  for (long i = 0; i < (long)n; ++i) {
    if (c.status[i] != Idle) continue;
    c.age_hyd[i] = (float)((double)c.age_hyd[i] + d_t);  // ages(idx,0) += d_t (double)
    const uint64_t pos = c.position[i];
    double flow = 0., vol = 0.;
    size_t k = 0;
    do {  // find_flow: do-while, first match wins
      const LeavingFlow& lf = c.leaving_flow[k++];
      if (pos == lf.index) { flow = lf.flow; vol = lf.volume; break; }
    } while (k < nf);
    if (flow != 0.) {
      const uint32_t ctr[4] = {(uint32_t)i >> 2, c.step, 1u, c.rank};
      uint32_t r[4];
      Philox::block(ctr, key, r);
      const float rng1 = u01f(r[i & 3]);
      const int leave_mask = (int)p_leave_precise(rng1, vol, flow, d_t);
      dead += leave_mask;
      c.age_hyd[i] *= (float)(1 - leave_mask);
      c.status[i] = (uint8_t)((int)c.status[i] * (1 - leave_mask) + (int)Exit * leave_mask);
    }
  }
  c.events[EvExit] += dead;
  return dead;
}

// --- remove_inactive_particles / CompactParticlesFunctor:
// particles_container.hpp:735-796, 292-385.  Serial-order semantics of the
// scan functor: for inactive slot i ascending, pull replacements from the tail
// (last_used - offset++), skipping non-idle slots and i itself; copy position,
// model row, contribs row, both ages; mark slot i Idle.
static void remove_inactive(Ctx& c, size_t to_remove) {
  if (to_remove == 0) return;
  if (to_remove == c.n_used) {
    c.n_used = 0; c.inactive_counter = 0; c.n_compactions++; return;
  }
  if (to_remove > c.n_used) { c.err = "remove_inactive_particles: cannot remove more element than existing"; return; }
  const size_t last = c.n_used - 1;
  size_t offset = 0, scan = 0;
  for (size_t i = 0; i < c.n_used; ++i) {
    const bool inactive = c.status[i] != Idle;
    const size_t scan_index = scan;
    scan += inactive ? 1 : 0;
    if (inactive && scan_index < to_remove) {
      size_t r = last - offset++;
      while (c.status[r] != Idle || r == i) r = last - offset++;
      c.status[i] = Idle;
      c.position[i] = c.position[r];
      for (int k = 0; k < c.n_var; ++k) c.model_v[i * c.n_var + k] = c.model_v[r * c.n_var + k];
      for (int k = 0; k < c.n_c; ++k) c.contribs[i * c.n_c + k] = c.contribs[r * c.n_c + k];
      c.age_hyd[i] = c.age_hyd[r]; c.age_div[i] = c.age_div[r];
    }
  }
  c.n_used -= to_remove;
  // Slots >= n_used are outside the container from here on; the reference leaves
  // whatever status they had and relies on zero-initialised (Idle) storage for
  // appended newborns (particles_container.hpp:403-443).  Make that explicit.
  for (size_t i = c.n_used; i <= last; ++i) c.status[i] = Idle;
  if (c.n_used <= (size_t)(c.rt.shrink_ratio * (double)c.n_allocated))
    resize_container(c, (size_t)((double)c.n_used * c.rt.allocation_factor), true);
  c.inactive_counter -= to_remove;
  c.n_compactions++;
}
// --- update_and_remove_inactive: particles_container.hpp:539-557
static void update_and_remove_inactive(Ctx& c, size_t out, size_t dead) {
  c.inactive_counter += out; c.inactive_counter += dead;
  const uint64_t thr = std::max<uint64_t>(c.rt.minimum_dead_particle_removal,
                                          (uint64_t)((double)c.n_used * c.rt.dead_particle_ratio_threshold));
  if (c.inactive_counter > thr) remove_inactive(c, c.inactive_counter);
}
// --- merge_buffer + InsertFunctor: particles_container.hpp:575-599, 403-443
static void merge_buffer(Ctx& c) {
  const uint64_t orig = c.n_used, n_add = c.buffer_index;
  if (n_add == 0) return;
  resize_container(c, orig + n_add, false);
  for (uint64_t i = 0; i < n_add; ++i) {
    for (int k = 0; k < c.n_var; ++k) c.model_v[(orig + i) * c.n_var + k] = c.buffer_model[i * c.n_var + k];
    c.position[orig + i] = c.buffer_position[i];
    c.age_hyd[orig + i] = 0; c.age_div[orig + i] = 0;
    c.status[orig + i] = Idle;
  }
  c.buffer_index = 0; c.n_used += n_add; c.total_new += n_add;
  allocate_buffer(c);
}

// --- cycleProcess: simulation.hpp:183-239; launch order kernels.hpp:160-224
// (model -> contribs) then :123-158 (move -> leave); post_cycle :213-239.
template <class M> static void cycle_process(Ctx& c, double d_t) {
  if (c.n_used == 0) { c.step++; return; }
  const bool enable_move = c.n_comp > 1;              // kernels.hpp:53-55
  const bool enable_leave = !c.leaving_flow.empty();  // kernels.hpp:56
  std::fill(c.sources.begin(), c.sources.end(), 0.0);  // contribs_scatter.reset() + sync_prepare_next
  cycle_model<M>(c, d_t);
  contributions(c);
  if (enable_move) cycle_move(c, d_t);
  uint64_t out = 0;
  if (enable_leave) out = cycle_leave(c, d_t);
  c.last_out = out; c.total_out += out;
  update_and_remove_inactive(c, out, c.last_dead);
  merge_buffer(c);
  c.step++;
}

static void dispatch_cycle(Ctx& c, double d_t) {
  switch (c.model) {
    case M_FIXED_LENGTH: cycle_process<FixedLength>(c, d_t); break;
    case M_MONOD: cycle_process<Monod>(c, d_t); break;
    case M_SIMPLE_ACETATE: cycle_process<SimpleAcetate>(c, d_t); break;
    case M_WIDE_UDF: cycle_process<WideUdf>(c, d_t); break;
    case M_UDF_MINIMAL: cycle_process<UdfMinimal>(c, d_t); break;
  }
}

// --- Liquid ODE step ("next" row 1): implScalar.cpp:251-266 `performStep`
//   dm/dt = C*M - C*sink + sources ; mass += dt*dm ; C = mass * V^-1
// with M the (n_comp x n_comp) transition matrix given as COO (rows, cols,
// vals) and C (n_species x n_comp) species-fastest.
// sparse_from_coo (implScalar.cpp:79-129): the triplets become a compressed sparse matrix — duplicates summed, the rows
// of a column ascending — and `c * m_transition` adds the terms of an element in that order.  Returns (col, row, value)
// sorted accordingly.
struct CooEntry { uint64_t col, row; double val; };
static std::vector<CooEntry> compressed_transition(size_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals) {
  std::vector<CooEntry> t(nnz);
  for (size_t e = 0; e < nnz; ++e) t[e] = CooEntry{cols[e], rows[e], vals[e]};
  std::stable_sort(t.begin(), t.end(), [](const CooEntry& a, const CooEntry& b) { return a.col != b.col ? a.col < b.col : a.row < b.row; });
  std::vector<CooEntry> out;
  for (const CooEntry& e : t) {
    if (!out.empty() && out.back().col == e.col && out.back().row == e.row) out.back().val += e.val;
    else out.push_back(e);
  }
  return out;
}
static void ode_step(size_t ns, size_t ncomp, double dt, double* C, double* mass, const double* vol,
                     const double* sink, const double* sources, size_t nnz, const uint64_t* rows,
                     const uint64_t* cols, const double* vals) {
  std::vector<double> dm(ns * ncomp, 0.0);
  for (const CooEntry& e : compressed_transition(nnz, rows, cols, vals))
    for (size_t s = 0; s < ns; ++s) dm[s + ns * e.col] += C[s + ns * e.row] * e.val;
  for (size_t j = 0; j < ncomp; ++j)
    for (size_t s = 0; s < ns; ++s) {
      const size_t k = s + ns * j;
      dm[k] = (dm[k] - C[k] * sink[j]) + sources[k];  // Eigen evaluates `c*M - c*sink + sources` coefficient-wise, left to right
      mass[k] += dt * dm[k];
      C[k] = mass[k] * (1.0 / vol[j]);
    }
}

// Two-phase ode_step (TEST INFRASTRUCTURE, like everything in this file): SimulationUnit::ode_step with a gas phase
// (apps/libs/simulation/src/simulation.model.cpp:131-154):
//   mt_model.gas_liquid_mass_transfer():  mtr = (kla o (Cg o Henry - Cl)) * diag(V_liquid)   (hydro/mass_transfer.cpp:143-161)
//   gas_scalar->performStepGL(d_t, mtr, GasToLiquid = -1), liquid_scalar->performStepGL(d_t, mtr, LiquidToGas = +1):
//       dm/dt = C*M - C*sink + sources + sign*mtr ; mass += dt*dm ; C = mass * V^-1           (implScalar.cpp:229-249)
//   liquid_scalar->clearNegs(): values in (-1e-4*5e-3, 0) become 0                            (implScalar.cpp:270-296)
// Pinned against the reference's own implScalar.cpp / mass_transfer.cpp compiled over oracle/eigen_shim
// (oracle/ref_liquid.cpp, tests/test_reference_liquid.py: bit-identical trajectories).
static void ode_step_gl(size_t ns, size_t ncomp, double dt, double* Cl, double* ml, const double* vl, const double* sink_l,
                        const double* src_l, size_t nnz_l, const uint64_t* rl, const uint64_t* cl, const double* valsl, double* Cg,
                        double* mg, const double* vg, const double* sink_g, const double* src_g, size_t nnz_g, const uint64_t* rg,
                        const uint64_t* cg, const double* valsg, const double* kla, const double* henry, double* mtr) {
  const size_t nb = ns * ncomp;
  for (size_t j = 0; j < ncomp; ++j)
    for (size_t s = 0; s < ns; ++s) { const size_t k = s + ns * j; mtr[k] = kla[k] * (Cg[k] * henry[s] - Cl[k]) * vl[j]; }
  auto step = [&](double* C, double* mass, const double* vol, const double* sink, const double* src, size_t nnz, const uint64_t* rows,
                  const uint64_t* cols, const double* vals, double sign) {
    std::vector<double> dm(nb, 0.0);
    for (const CooEntry& e : compressed_transition(nnz, rows, cols, vals))
      for (size_t s = 0; s < ns; ++s) dm[s + ns * e.col] += C[s + ns * e.row] * e.val;
    for (size_t j = 0; j < ncomp; ++j)
      for (size_t s = 0; s < ns; ++s) {
        const size_t k = s + ns * j;
        dm[k] = (dm[k] - C[k] * sink[j]) + src[k];
        dm[k] += sign * mtr[k];
        mass[k] += dt * dm[k];
        C[k] = mass[k] * (1.0 / vol[j]);
      }
  };
  step(Cg, mg, vg, sink_g, src_g, nnz_g, rg, cg, valsg, -1.0);
  step(Cl, ml, vl, sink_l, src_l, nnz_l, rl, cl, valsl, 1.0);
  for (size_t k = 0; k < nb; ++k) if (Cl[k] < 0.0 && std::fabs(Cl[k]) < 1e-4 * 5e-3) Cl[k] = 0.0;
}

// kla of Type::FlowmapTurbulence (hydro/impl_mtr.cpp:22-83, 107-140): row 1 (oxygen) = kl * a with
//   kl = 0.3 * (eps * nu)^0.25 * Sc^-0.5,  a = 6 alpha_g / (db (1 - alpha_g)),  alpha_g = Vg / (Vl + Vg),
//   nu = c_kinematic_viscosity(20 C), Sc = nu / 1e-9; the other rows keep what FunctorKla left (0).
static double c_kinematic_viscosity(double temp) {
  if (temp < 85)
    return std::round((0.00000000000282244333 * std::pow(temp, 6) - 0.00000000126441088087 * std::pow(temp, 5) +
                       0.00000023336659710795 * std::pow(temp, 4) - 0.0000234079044336466 * std::pow(temp, 3) +
                       0.00144686943485654 * std::pow(temp, 2) - 0.0607310297913931 * temp + 1.79194000343777) * 0.000001 * 10000000000) / 10000000000;
  return std::round((0.00000000000000178038 * std::pow(temp, 6) - 0.00000000000277495333 * std::pow(temp, 5) +
                     0.00000000181964246491 * std::pow(temp, 4) - 0.00000064995487357883 * std::pow(temp, 3) +
                     0.000136367622445752 * std::pow(temp, 2) - 0.0166081298727911 * temp + 1.08486933174497) * 0.000001 * 10000000000) / 10000000000;
}
static void kla_flowmap_turbulence(size_t ns, size_t ncomp, double db, const double* eps, const double* vl, const double* vg, double* kla) {
  const double nu = c_kinematic_viscosity(20.0), sc = nu / 1e-9;
  for (size_t j = 0; j < ncomp; ++j) {
    const double alpha = vg[j] / (vl[j] + vg[j]);
    const double kl = 0.3 * std::pow(eps[j] * nu, 0.25) * std::pow(sc, -0.5);
    const double a = 6. * alpha / (db * (1 - alpha));
    if (ns > 1) kla[1 + ns * j] = kl * a;
  }
}

}  // namespace orc

// =============================================================================
// C API (ctypes).  0 = ok, negative = error (api_raw.cpp:217-231 conventions).
// =============================================================================
using namespace orc;
extern "C" {

void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox::block(ctr, key, out); }

void* orc_create(int model, int n_var_udf, uint64_t n_species, uint64_t n_comp, uint64_t seed, uint32_t rank,
                 int n_threads) {
  Ctx* c = new (std::nothrow) Ctx();
  if (!c) return nullptr;
  c->model = model;
  switch (model) {
    case M_FIXED_LENGTH: c->n_var = FixedLength::n_var; c->n_c = FixedLength::n_c; break;
    case M_MONOD: c->n_var = Monod::n_var; c->n_c = Monod::n_c; break;
    case M_SIMPLE_ACETATE: c->n_var = SimpleAcetate::n_var; c->n_c = SimpleAcetate::n_c; break;
    case M_WIDE_UDF: c->n_var = n_var_udf; c->n_c = WideUdf::n_c; WideUdf::n_var_rt = n_var_udf; break;
    case M_UDF_MINIMAL: c->n_var = UdfMinimal::n_var; c->n_c = UdfMinimal::n_c; break;
    default: delete c; return nullptr;
  }
  c->n_species = n_species; c->n_comp = n_comp; c->seed = seed; c->rank = rank;
  c->n_threads = n_threads > 0 ? n_threads : 1;
  c->concentrations.assign(n_species * n_comp, 0.0);
  c->sources.assign(n_species * n_comp, 0.0);
  c->liquid_volume.assign(n_comp, 1.0); c->diag_transition.assign(n_comp, 0.0);
  return c;
}
void orc_destroy(void* h) { delete (Ctx*)h; }
int orc_n_var(void* h) { return ((Ctx*)h)->n_var; }
int orc_n_c(void* h) { return ((Ctx*)h)->n_c; }
int orc_set_runtime(void* h, uint64_t min_removal, double buffer_ratio, double alloc_factor, double shrink_ratio,
                    double dead_ratio) {
  Ctx& c = *(Ctx*)h;
  c.rt.minimum_dead_particle_removal = min_removal; c.rt.buffer_ratio = buffer_ratio;
  c.rt.allocation_factor = alloc_factor; c.rt.shrink_ratio = shrink_ratio;
  c.rt.dead_particle_ratio_threshold = dead_ratio;
  return 0;
}
int orc_set_quirk_contrib_return(void* h, int on) { ((Ctx*)h)->quirk_contrib_return = on != 0; return 0; }
int orc_set_step(void* h, uint32_t step) { ((Ctx*)h)->step = step; return 0; }

// props: SoA columns [n_var][n] (the exchange layout of include/bmc.h)
int orc_set_particles(void* h, uint64_t n, const float* props, const uint64_t* pos, const uint8_t* status,
                      const float* age_hyd, const float* age_div) {
  Ctx& c = *(Ctx*)h;
  c.n_allocated = 0; c.n_used = n; c.inactive_counter = 0;
  c.position.clear(); c.model_v.clear(); c.contribs.clear(); c.status.clear(); c.age_hyd.clear(); c.age_div.clear();
  resize_container(c, n, false);  // ParticlesContainer ctor, particles_container.hpp:692-727
  c.buffer_cap = 0; allocate_buffer(c);
  for (uint64_t i = 0; i < n; ++i) {
    for (int k = 0; k < c.n_var; ++k) c.model_v[i * c.n_var + k] = props[(size_t)k * n + i];
    c.position[i] = pos ? pos[i] : 0;
    c.status[i] = status ? status[i] : (uint8_t)Idle;
    c.age_hyd[i] = age_hyd ? age_hyd[i] : 0.0f;
    c.age_div[i] = age_div ? age_div[i] : 0.0f;
    if (c.status[i] != Idle) c.inactive_counter++;
    if (c.position[i] >= c.n_comp) { c.err = "position out of range"; return -2; }
  }
  return 0;
}
uint64_t orc_n_used(void* h) { return ((Ctx*)h)->n_used; }
uint64_t orc_capacity(void* h) { return ((Ctx*)h)->n_allocated; }
uint64_t orc_buffer_capacity(void* h) { return ((Ctx*)h)->buffer_cap; }
uint64_t orc_inactive(void* h) { return ((Ctx*)h)->inactive_counter; }
int orc_get_particles(void* h, uint64_t n, float* props, uint64_t* pos, uint8_t* status, float* age_hyd,
                      float* age_div) {
  Ctx& c = *(Ctx*)h;
  if (n > c.n_used) return -1;
  for (uint64_t i = 0; i < n; ++i) {
    if (props) for (int k = 0; k < c.n_var; ++k) props[(size_t)k * n + i] = c.model_v[i * c.n_var + k];
    if (pos) pos[i] = c.position[i];
    if (status) status[i] = c.status[i];
    if (age_hyd) age_hyd[i] = c.age_hyd[i];
    if (age_div) age_div[i] = c.age_div[i];
  }
  return 0;
}
int orc_set_weight(void* h, double w) { ((Ctx*)h)->weight = (float)w; return 0; }  // deep_copy(weights, new_weight) unit.cpp:243
// ReactorDomain::update domain.cpp:43-74
int orc_domain_update(void* h, const double* volumes, const uint64_t* neighbors_flat, const double* out_flows,
                      const double* proba_flat, uint64_t n_cols) {
  Ctx& c = *(Ctx*)h;
  c.n_cols = n_cols;
  c.liquid_volume.assign(volumes, volumes + c.n_comp);
  c.diag_transition.assign(out_flows, out_flows + c.n_comp);
  c.neighbors.assign(neighbors_flat, neighbors_flat + c.n_comp * n_cols);
  c.cumulative_probability.assign(proba_flat, proba_flat + c.n_comp * n_cols);
  for (uint64_t v : c.neighbors) if (v >= c.n_comp) { c.err = "neighbor out of range"; return -2; }
  return 0;
}
// ReactorDomain::set_leaving_flow domain.cpp:96-108 (only populated outlets, Q21)
int orc_set_leaving_flows(void* h, uint64_t n, const uint64_t* index, const double* flow, const double* volume) {
  Ctx& c = *(Ctx*)h;
  c.leaving_flow.resize(n);
  for (uint64_t i = 0; i < n; ++i) c.leaving_flow[i] = LeavingFlow{index[i], flow[i], volume[i]};
  return 0;
}
int orc_set_concentrations(void* h, const double* conc) {
  Ctx& c = *(Ctx*)h;
  std::copy(conc, conc + c.n_species * c.n_comp, c.concentrations.begin());
  return 0;
}
int orc_cycle(void* h, double d_t) {
  Ctx& c = *(Ctx*)h;
  if (c.n_comp > 1 && c.n_cols == 0) { c.err = "domain not set"; return -3; }
  dispatch_cycle(c, d_t);
  return c.err.empty() ? 0 : -1;
}
int orc_get_sources(void* h, double* out) {
  Ctx& c = *(Ctx*)h;
  std::copy(c.sources.begin(), c.sources.end(), out);
  return 0;
}
// counters[0..5] = events (events.hpp:17-26 order), [6]=n_used, [7]=inactive,
// [8]=last out, [9]=last dead_total, [10]=last waiting_allocation, [11]=buffer_index,
// [12]=capacity, [13]=total_out, [14]=total_new, [15]=n_compactions
int orc_get_counters(void* h, uint64_t* out) {
  Ctx& c = *(Ctx*)h;
  for (int i = 0; i < N_EVENTS; ++i) out[i] = c.events[i];
  out[6] = c.n_used; out[7] = c.inactive_counter; out[8] = c.last_out; out[9] = c.last_dead;
  out[10] = c.last_waiting; out[11] = c.buffer_index; out[12] = c.n_allocated; out[13] = c.total_out;
  out[14] = c.total_new; out[15] = c.n_compactions;
  return 0;
}
// MonteCarloUnit::getRepartition unit.cpp:190-230 (Idle particles only)
int orc_repartition(void* h, uint64_t* out) {
  Ctx& c = *(Ctx*)h;
  std::fill(out, out + c.n_comp, 0);
  for (uint64_t i = 0; i < c.n_used; ++i) if (c.status[i] == Idle) out[c.position[i]]++;
  return 0;
}
// PostProcessing::get_properties post_process.hpp:33-118, 173-250: force_remove_dead, then per Idle particle the
// exported properties + mass as doubles, their per-compartment sums, and both ages.
static double model_mass(const Ctx& c, size_t i) {
  const float lin = c_linear_density(1000.0f, (float)0.6e-6);
  return c.model_v[i * (size_t)c.n_var + 0] * lin;  // every model here: mass = length * lin_density (property 0)
}
int orc_get_properties(void* h, const uint64_t* indices, uint64_t n_indices, double* pv, double* sv, double* ages, uint64_t* n_out) {
  Ctx& c = *(Ctx*)h;
  remove_inactive(c, c.inactive_counter);
  const uint64_t n_p = c.n_used;
  if (n_out) *n_out = n_p;
  if (!pv && !sv && !ages) return 0;
  const size_t n_exp = indices ? n_indices : (size_t)c.n_var;
  if (sv) std::fill(sv, sv + (n_exp + 1) * c.n_comp, 0.0);
  for (uint64_t i = 0; i < n_p; ++i) {
    if (c.status[i] != Idle) continue;
    for (size_t e = 0; e < n_exp; ++e) {
      const size_t k = indices ? indices[e] : e;
      const double cur = c.model_v[i * (size_t)c.n_var + k];
      if (sv) sv[e * c.n_comp + c.position[i]] += cur;
      if (pv) pv[e * n_p + i] = cur;
    }
    const double m = model_mass(c, i);
    if (sv) sv[n_exp * c.n_comp + c.position[i]] += m;
    if (pv) pv[n_exp * n_p + i] = m;
    if (ages) { ages[i] = c.age_hyd[i]; ages[n_p + i] = c.age_div[i]; }
  }
  return 0;
}
// force_remove_dead particles_container.hpp:463-468
int orc_compact(void* h) { Ctx& c = *(Ctx*)h; remove_inactive(c, c.inactive_counter); return c.err.empty() ? 0 : -1; }
// test_container.cpp:62-148 drives handle_division / merge_buffer directly
int orc_handle_division(void* h, uint64_t idx) {
  Ctx& c = *(Ctx*)h;
  if (c.buffer_index < c.buffer_cap) {
    const uint64_t j = c.buffer_index++;
    const Arr arr{c.model_v.data(), c.n_var}; const Arr buf{c.buffer_model.data(), c.n_var};
    Gen g(c.seed, c.rank, (uint32_t)idx, c.step);
    g.ctr[2] = 0x40000000u;
    switch (c.model) {
      case M_FIXED_LENGTH: FixedLength::division(g, idx, j, arr, buf); break;
      case M_MONOD: Monod::division(g, idx, j, arr, buf); break;
      case M_SIMPLE_ACETATE: SimpleAcetate::division(g, idx, j, arr, buf); break;
      case M_WIDE_UDF: WideUdf::division(g, idx, j, arr, buf); break;
      case M_UDF_MINIMAL: UdfMinimal::division(g, idx, j, arr, buf); break;
    }
    c.buffer_position[j] = c.position[idx]; c.age_div[idx] = 0;
    return 1;
  }
  return 0;
}
int orc_merge_buffer(void* h) { merge_buffer(*(Ctx*)h); return 0; }
int orc_set_status(void* h, uint64_t idx, uint8_t s) {
  Ctx& c = *(Ctx*)h;
  if (c.status[idx] == Idle && s != Idle) c.inactive_counter++;
  c.status[idx] = s; return 0;
}
// MC::init<M> / InitFunctor  mcinit.hpp:67-105, unit.cpp:102-163: M::init, position
// = urand64(min_c,max_c), total mass.  `linit` feeds configurable models
// (fixed_length's Config view, fixed_length.hpp:109-120).
int orc_init_particles(void* h, uint64_t n, int uniform_pos, const float* linit, double* total_mass) {
  Ctx& c = *(Ctx*)h;
  c.n_allocated = 0; c.n_used = n; c.inactive_counter = 0;
  c.position.clear(); c.model_v.clear(); c.contribs.clear(); c.status.clear(); c.age_hyd.clear(); c.age_div.clear();
  resize_container(c, n, false);
  c.buffer_cap = 0; allocate_buffer(c);
  const Arr arr{c.model_v.data(), c.n_var};
  const uint64_t max_c = uniform_pos ? c.n_comp : 1;
  double m = 0;
  for (uint64_t i = 0; i < n; ++i) {
    Gen g(c.seed, c.rank, (uint32_t)i, 0xFFFFFFFFu);  // step id reserved for init
    switch (c.model) {
      case M_FIXED_LENGTH: FixedLength::init(g, i, arr, linit ? linit[i] : 1.5e-6f); m += FixedLength::mass(i, arr); break;
      case M_MONOD: Monod::init(g, i, arr); m += Monod::mass(i, arr); break;
      case M_SIMPLE_ACETATE: SimpleAcetate::init(g, i, arr); m += SimpleAcetate::mass(i, arr); break;
      case M_WIDE_UDF: WideUdf::init(g, i, arr, linit ? linit[i] : 1.5e-6f); m += WideUdf::mass(i, arr); break;
      case M_UDF_MINIMAL: UdfMinimal::init(g, i, arr, linit ? linit[i] : 1.5e-6f); m += UdfMinimal::mass(i, arr); break;
    }
    c.position[i] = g.urand64(0, max_c);
  }
  if (total_mass) *total_mass = m;
  return 0;
}
// distribution samplers for the moment tests (test_rng_2.cpp)
int orc_sample(int kind, uint64_t seed, uint64_t n, double p0, double p1, double p2, double p3, double* out) {
  for (uint64_t i = 0; i < n; ++i) {
    Gen g(seed, 0, (uint32_t)i, (uint32_t)(i >> 32));
    switch (kind) {
      case 0: out[i] = g.normal(p0, p1); break;
      case 1: out[i] = lognormal(g, p0, p1); break;
      case 2: out[i] = truncated_normal<double>(g, p0, p1, p2, p3); break;
      case 3: out[i] = (double)truncated_normal<float>(g, (float)p0, (float)p1, (float)p2, (float)p3); break;
      case 4: out[i] = (double)(-1.0f * ln_f32(g.frand()) / (float)p0); break;  // Exponential<float>
      case 5: out[i] = g.drand(); break;
      case 6: out[i] = (double)g.frand(); break;
      case 7: out[i] = (double)norminv<double>(g.drand(), p0, p1); break;
      default: return -1;
    }
  }
  return 0;
}
int orc_ode_step(uint64_t ns, uint64_t ncomp, double dt, double* C, double* mass, const double* vol, const double* sink,
                 const double* sources, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals) {
  ode_step(ns, ncomp, dt, C, mass, vol, sink, sources, nnz, rows, cols, vals);
  return 0;
}
int orc_ode_step_gl(uint64_t ns, uint64_t ncomp, double dt, double* Cl, double* ml, const double* vl, const double* sink_l,
                    const double* src_l, uint64_t nnz_l, const uint64_t* rl, const uint64_t* cl, const double* valsl, double* Cg,
                    double* mg, const double* vg, const double* sink_g, const double* src_g, uint64_t nnz_g, const uint64_t* rg,
                    const uint64_t* cg, const double* valsg, const double* kla, const double* henry, double* mtr) {
  ode_step_gl(ns, ncomp, dt, Cl, ml, vl, sink_l, src_l, nnz_l, rl, cl, valsl, Cg, mg, vg, sink_g, src_g, nnz_g, rg, cg, valsg, kla, henry, mtr);
  return 0;
}
int orc_kla_flowmap_turbulence(uint64_t ns, uint64_t ncomp, double db, const double* eps, const double* vl, const double* vg, double* kla) {
  kla_flowmap_turbulence(ns, ncomp, db, eps, vl, vg, kla);
  return 0;
}
const char* orc_last_error(void* h) { return ((Ctx*)h)->err.c_str(); }
int orc_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}  // extern "C"
