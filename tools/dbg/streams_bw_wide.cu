// Streaming bound of the WIDE-model access pattern: R read + W write float columns, VEC floats per lane per column
// (VEC 1: 128 B per warp access, VEC 2: 256 B, VEC 4: 512 B), THREADS per block, one block per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC> struct V;
template <> struct V<1> { using T = float; };
template <> struct V<2> { using T = float2; };
template <> struct V<4> { using T = float4; };
template <int R, int W, int VEC, int THREADS> __global__ void __launch_bounds__(THREADS, 1) k(const float* __restrict__ in, float* __restrict__ out, size_t cap, size_t n_groups) {
  using T = typename V<VEC>::T;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t gw = (size_t)blockIdx.x * (THREADS / 32) + warp, nw = (size_t)gridDim.x * (THREADS / 32);
  for (size_t g = gw; g < n_groups; g += nw) {
    const size_t i0 = g * (32 * VEC) + lane * VEC;
    T v[R];
#pragma unroll
    for (int c = 0; c < R; ++c) v[c] = *reinterpret_cast<const T*>(in + (size_t)c * cap + i0);
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < R; ++c) s += reinterpret_cast<const float*>(&v[c])[0];
#pragma unroll
    for (int c = 0; c < W; ++c) { T o = v[c % R]; reinterpret_cast<float*>(&o)[0] = s + c; *reinterpret_cast<T*>(out + (size_t)c * cap + i0) = o; }
  }
}
template <int R, int W, int VEC, int THREADS> void run(size_t n, float* in, float* out, size_t cap) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t ng = n / (32 * VEC);
  for (int i = 0; i < 2; ++i) k<R, W, VEC, THREADS><<<148, THREADS>>>(in, out, cap, ng);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int i = 0; i < reps; ++i) k<R, W, VEC, THREADS><<<148, THREADS>>>(in, out, cap, ng);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)(R + W) * 4.0 * (double)n * reps;
  printf("n=%zu R=%d W=%d VEC=%d threads=%d  %.1f us/launch  %.0f GB/s  (%s)\n", n, R, W, VEC, THREADS, ms * 1e3 / reps, bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const size_t n = (size_t)250000000 / 1024 * 1024;
  const size_t cap = n;
  float *in, *out;
  cudaMalloc(&in, cap * 33 * 4); cudaMalloc(&out, cap * 32 * 4);
  cudaMemset(in, 0, cap * 33 * 4); cudaMemset(out, 0, cap * 32 * 4);
  run<33, 31, 1, 768>(n, in, out, cap);
  run<33, 31, 1, 1024>(n, in, out, cap);
  run<33, 31, 2, 512>(n, in, out, cap);
  run<33, 31, 2, 768>(n, in, out, cap);
  run<33, 31, 4, 512>(n, in, out, cap);
  run<33, 31, 4, 256>(n, in, out, cap);
  return 0;
}
