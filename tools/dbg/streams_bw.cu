// How much HBM bandwidth does the ACCESS PATTERN of the particle pass allow, with no compute at all?
// R read columns + W write columns of float, each warp handles groups of 128 consecutive slots with 128-bit accesses
// (512 B per warp per column), persistent grid of one 1024-thread block per SM, grid-stride over the groups.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o streams_bw streams_bw.cu && ./streams_bw
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
template <int R, int W> __global__ void __launch_bounds__(1024, 1) k(const float* __restrict__ in, float* __restrict__ out, size_t cap, size_t n_groups) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t gw = (size_t)blockIdx.x * 32 + warp, nw = (size_t)gridDim.x * 32;
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
template <int R, int W> void run(size_t n, float* in, float* out, size_t cap) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t ng = n / 128;
  for (int i = 0; i < 3; ++i) k<R, W><<<148, 1024>>>(in, out, cap, ng);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) k<R, W><<<148, 1024>>>(in, out, cap, ng);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)(R + W) * 4.0 * (double)n * reps;
  printf("n=%zu R=%d W=%d  %.1f us/launch  %.0f GB/s\n", n, R, W, ms * 1e3 / reps, bytes / (ms * 1e-3) / 1e9);
}
int main() {
  for (size_t n : {(size_t)10000000 / 128 * 128, (size_t)125000000 / 128 * 128}) {
    const size_t cap = (n * 2 + 1023) / 1024 * 1024;
    float *in, *out;
    cudaMalloc(&in, cap * 8 * 4); cudaMalloc(&out, cap * 8 * 4);
    cudaMemset(in, 0, cap * 8 * 4); cudaMemset(out, 0, cap * 8 * 4);
    run<1, 1>(n, in, out, cap); run<2, 2>(n, in, out, cap); run<4, 4>(n, in, out, cap); run<6, 4>(n, in, out, cap); run<8, 8>(n, in, out, cap);
    run<5, 5>(n, in, out, cap); run<1, 0>(n, in, out, cap); run<5, 0>(n, in, out, cap);
    cudaFree(in); cudaFree(out);
  }
  return 0;
}
