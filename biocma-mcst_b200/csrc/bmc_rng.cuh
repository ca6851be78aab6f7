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
