// 6 read + 4 write columns (the monod pass) at different block sizes: how many bytes in flight does the pattern need?
#include <cstdio>
#include <cuda_runtime.h>
template <int R, int W, int THREADS> __global__ void __launch_bounds__(THREADS, 1) k(const float* __restrict__ in, float* __restrict__ out, size_t cap, size_t n_groups) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t gw = (size_t)blockIdx.x * (THREADS / 32) + warp, nw = (size_t)gridDim.x * (THREADS / 32);
  for (size_t g = gw; g < n_groups; g += nw) {
    const size_t i0 = g * 128 + lane * 4;
    float4 v[R];
#pragma unroll
    for (int c = 0; c < R; ++c) v[c] = *reinterpret_cast<const float4*>(in + (size_t)c * cap + i0);
    float4 s = v[0];
#pragma unroll
    for (int c = 1; c < R; ++c) { s.x += v[c].x; s.y += v[c].y; s.z += v[c].z; s.w += v[c].w; }
#pragma unroll
    for (int c = 0; c < W; ++c) { float4 o = s; o.x += c; *reinterpret_cast<float4*>(out + (size_t)c * cap + i0) = o; }
  }
}
template <int THREADS> void run(size_t n, float* in, float* out, size_t cap) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t ng = n / 128;
  for (int i = 0; i < 2; ++i) k<6, 4, THREADS><<<148, THREADS>>>(in, out, cap, ng);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) k<6, 4, THREADS><<<148, THREADS>>>(in, out, cap, ng);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("6R4W threads=%4d (%2d warps/SM, %3d KB of loads in flight per SM)  %.1f us/launch  %.0f GB/s\n", THREADS, THREADS / 32, THREADS / 32 * 3,
         ms * 1e3 / reps, 40.0 * (double)n * reps / (ms * 1e-3) / 1e9);
}
int main() {
  const size_t n = (size_t)125000000 / 128 * 128, cap = (n * 2 + 1023) / 1024 * 1024;
  float *in, *out;
  cudaMalloc(&in, cap * 6 * 4); cudaMalloc(&out, cap * 4 * 4);
  cudaMemset(in, 0, cap * 6 * 4); cudaMemset(out, 0, cap * 4 * 4);
  run<128>(n, in, out, cap); run<256>(n, in, out, cap); run<384>(n, in, out, cap); run<512>(n, in, out, cap); run<768>(n, in, out, cap); run<1024>(n, in, out, cap);
  return 0;
}
