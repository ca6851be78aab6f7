// Counter-based Philox4x32-10 for the sm_100a particle kernels.
//
// Replaces MC::pool_type = Kokkos::Random_XorShift1024_Pool
// (apps/libs/mc/public/mc/alias.hpp:98-102): no generator state lives in memory,
// no pool acquire/release (move_kernel.hpp:237-259) — every draw is a pure
// function of (seed, rank | slot, step, draw-block), held in registers.
//   key = { seed_lo, seed_hi }
//   ctr = { index, step, draw_block, rank }
// draw_block 0: index = slot >> 2; word (slot & 3) = u1 of that slot (leave-compartment
//               test).  One Philox block serves the four slots a thread owns.
// draw_block 1: index = slot >> 2; word (slot & 3) = u3 (outlet test), computed only by
//               warps that have a particle sitting in an outlet compartment.
// draw_block 2: index = slot; word 0 = u2 (neighbour pick), computed only for movers.
// draw_block 3.. (index = slot): generator handed to M::update / M::init;
// draw_block 0x40000001.. : generator handed to M::division.
#pragma once
#ifdef __CUDACC_RTC__  // NVRTC (user-defined models are JIT-compiled): no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#else
#include <cstdint>
#include <cuda_runtime.h>
#endif

namespace bmc {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// -----------------------------------------------------------------------------
// The same function with everything that does not depend on the particle hoisted to the host.
// In the step kernel only counter word 0 (the slot index) differs between threads; step, draw
// block and rank are launch constants.  Following the rounds with (T)hread / (U)niform values:
//   round 0:  c0' = hi(M1*blk) ^ step ^ k0[0]   (U)        c1' = lo(M1*blk)            (U)
//             c2' = hi(M0*c0) ^ (rank ^ k1[0])  (T)        c3' = lo(M0*c0)             (T)
//   round 1:  c0" = hi(M1*c2') ^ (c1' ^ k0[1])  (T)        c1" = lo(M1*c2')            (T)
//             c2" = c3' ^ (hi(M0*c0') ^ k1[1])  (T)        c3" = lo(M0*c0')            (U)
//   round 2:  the only uniform input left is c3", folded into the key: hi(M0*c0") ^ (c3" ^ k1[2])
// so rounds 0 and 1 cost one multiplication each instead of two, and the round keys (k + r*W)
// come from the constant bank instead of being re-derived by every thread: 37 instructions per
// block instead of ~66.  philox_pre() is evaluated by the host once per launch and draw block.
// -----------------------------------------------------------------------------
struct PhiloxPre {
  uint32_t n2x;           // rank ^ k1[0]
  uint32_t x2, y2, z3;    // folded uniform terms of rounds 1 and 2
  uint32_t k0[10], k1[10];  // round keys (entries 0..1 of k0 and 0..2 of k1 are folded above)
};
__host__ __device__ inline PhiloxPre philox_pre(uint32_t step, uint32_t blk, uint32_t rank, uint32_t key0, uint32_t key1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  PhiloxPre p;
  for (int r = 0; r < 10; ++r) { p.k0[r] = key0 + (uint32_t)r * W0; p.k1[r] = key1 + (uint32_t)r * W1; }
  const uint64_t pb = (uint64_t)M1 * blk;
  const uint32_t a = (uint32_t)(pb >> 32) ^ step ^ p.k0[0];  // c0 after round 0
  const uint32_t b = (uint32_t)pb;                           // c1 after round 0
  const uint64_t pa = (uint64_t)M0 * a;
  p.n2x = rank ^ p.k1[0];
  p.x2 = b ^ p.k0[1];
  p.y2 = (uint32_t)(pa >> 32) ^ p.k1[1];
  p.z3 = (uint32_t)pa ^ p.k1[2];
  return p;
}
#ifdef __CUDACC__
__device__ __forceinline__ void philox4x32_10_idx(uint32_t c0, const PhiloxPre& P, uint32_t out[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  // round 0
  uint32_t c2 = __umulhi(M0, c0) ^ P.n2x, c3 = M0 * c0;
  // round 1
  uint32_t n0 = __umulhi(M1, c2) ^ P.x2, c1 = M1 * c2;
  c2 = c3 ^ P.y2; c0 = n0;
  // round 2
  {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ P.k0[2]; c1 = lo1; c2 = hi0 ^ P.z3; c3 = lo0;
  }
#pragma unroll
  for (int r = 3; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ P.k0[r]; c1 = lo1; c2 = hi0 ^ c3 ^ P.k1[r]; c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
#endif

// gen.frand(0.,1.): 24-bit uniform in [0,1)
__host__ __device__ __forceinline__ float u01f(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// gen.drand(): 53-bit uniform in [0,1)
__host__ __device__ __forceinline__ double u01d(uint32_t hi, uint32_t lo) {
  const unsigned long long v = (((unsigned long long)hi << 32) | lo) >> 11;
  return (double)v * (1.0 / 9007199254740992.0);
}

// Generator object passed to the model hooks in place of the Kokkos pool
// (traits.hpp:116-118).  Exposes the subset of the Kokkos generator interface
// the reference's models use: frand / drand / urand64 / normal.
struct Gen {
  uint32_t k0, k1, c0, c1, c2, c3;
  uint32_t buf[4];
  int have;
  __device__ __forceinline__ Gen(uint32_t seed_lo, uint32_t seed_hi, uint32_t rank, uint32_t slot, uint32_t step,
                                 uint32_t first_block)
      : k0(seed_lo), k1(seed_hi), c0(slot), c1(step), c2(first_block), c3(rank), have(0) {}
  __device__ __forceinline__ uint32_t next32() {
    if (have == 0) { c2 += 1; philox4x32_10(c0, c1, c2, c3, k0, k1, buf); have = 4; }
    const int i = 4 - have;
    --have;
    return i == 0 ? buf[0] : (i == 1 ? buf[1] : (i == 2 ? buf[2] : buf[3]));
  }
  __device__ __forceinline__ float frand() { return u01f(next32()); }
  __device__ __forceinline__ float frand(float a, float b) { return a + (b - a) * frand(); }
  __device__ __forceinline__ double drand() { const uint32_t hi = next32(); const uint32_t lo = next32(); return u01d(hi, lo); }
  __device__ __forceinline__ double drand(double a, double b) { return a + (b - a) * drand(); }
  __device__ __forceinline__ unsigned long long urand64(unsigned long long lo, unsigned long long hi) {
    const unsigned long long v = ((unsigned long long)next32() << 32) | next32();
    return lo + v % (hi - lo);
  }
  // Kokkos 5.1.1 normal(): Marsaglia polar method on drand()
  __device__ __forceinline__ double normal() {
    double S = 2.0, U = 0.0;
    while (S >= 1.0) {
      U = 2.0 * drand() - 1.0;
      const double V = 2.0 * drand() - 1.0;
      S = U * U + V * V;
    }
    return U * sqrt(-2.0 * log(S) / S);
  }
  __device__ __forceinline__ double normal(double mu, double sigma) { return mu + sigma * normal(); }
};

}  // namespace bmc
