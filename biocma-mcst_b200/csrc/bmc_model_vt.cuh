// Per-model launch table.  Each model's kernels are instantiated in their own translation
// unit (bmc_inst_*.cu, compiled in parallel); bmc_api.cu only sees these function pointers.
#pragma once
#include <string>
#include "bmc_kernels.cuh"

namespace bmc {

struct ModelVT {
  int n_var, n_c, vec;
  const void* cycle_fn;
  void (*launch_cycle)(const CycleParams&, int grid, size_t smem, cudaStream_t);
  void (*launch_init)(float*, size_t, uint32_t*, uint8_t*, float*, float*, unsigned long long, uint32_t, const float*,
                      uint32_t, uint32_t, uint32_t, DevState*, int, cudaStream_t);
  int ct;              // floats per compartment-table row
  size_t stage_bytes;  // dynamic shared memory of the cp.async pipeline (0 = direct loads)
  void (*launch_pre)(const PreParams&, int grid, cudaStream_t);
};

template <class M, int VEC, int MINB, bool PIPE> static void launch_cycle_t(const CycleParams& p, int grid, size_t smem, cudaStream_t s) {
  cycle_kernel<M, VEC, MINB, PIPE><<<grid, kBlock, smem, s>>>(p);
}
template <class M>
static void launch_init_t(float* props, size_t cap, uint32_t* pos, uint8_t* status, float* ah, float* ad, unsigned long long n,
                          uint32_t ncomp_hi, const float* linit, uint32_t slo, uint32_t shi, uint32_t rank, DevState* st, int grid,
                          cudaStream_t s) {
  init_kernel<M><<<grid, 256, 0, s>>>(props, cap, pos, status, ah, ad, n, ncomp_hi, linit, slo, shi, rank, st);
}
template <class M> static void launch_pre_t(const PreParams& p, int grid, cudaStream_t s) {
  pre_step_kernel<M><<<grid, 256, 0, s>>>(p);
}
template <class M, int VEC, int MINB = 1, bool PIPE = false> static ModelVT make_vt() {
  ModelVT v;
  v.n_var = M::n_var; v.n_c = M::n_c; v.vec = VEC;
  v.cycle_fn = (const void*)cycle_kernel<M, VEC, MINB, PIPE>;
  v.launch_cycle = &launch_cycle_t<M, VEC, MINB, PIPE>;
  v.stage_bytes = PIPE ? 2 * StageBytes<M, VEC>::value : 0;
  v.launch_init = &launch_init_t<M>;
  v.ct = 1 + M::n_pre;
  v.launch_pre = &launch_pre_t<M>;
  return v;
}

// Kernel variant = (slots per thread, min resident blocks per SM[, cp.async pipeline]).  Defaults
// were picked from sweeps on a B200 (tools/sweep.py); BMC_VARIANT="v<VEC>b<MINB>"
// overrides MINB for tuning runs ("p<VEC>b3" selects the cp.async pipeline where it is built).
template <class M, int VEC, bool WITH_PIPE = false> static bool pick_variant(const std::string& var, int def_minb, ModelVT& vt) {
  int minb = def_minb; bool pipe = false;
  if (var.size() == 4 && (var[0] == 'v' || var[0] == 'p') && var[2] == 'b' && var[1] - '0' == VEC) {
    minb = var[3] - '0'; pipe = var[0] == 'p';
  }
  if constexpr (WITH_PIPE) {
    if (pipe) { vt = make_vt<M, VEC, 3, true>(); return true; }
  }
  switch (minb) {
    case 1: case 2: vt = make_vt<M, VEC, 2>(); return true;
    case 3: vt = make_vt<M, VEC, 3>(); return true;
    default: vt = make_vt<M, VEC, 4>(); return true;
  }
}

bool pick_fixed_length(const std::string& var, ModelVT& vt);
bool pick_monod(const std::string& var, ModelVT& vt);
bool pick_simple_acetate(const std::string& var, ModelVT& vt);
bool pick_wide_udf_small(const std::string& var, int n_var, ModelVT& vt);   // P = 8, 16
bool pick_wide_udf_large(const std::string& var, int n_var, ModelVT& vt);   // P = 32, 64

}  // namespace bmc
