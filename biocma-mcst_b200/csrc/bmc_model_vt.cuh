// Per-model launch table.  Each model's kernels are instantiated in their own translation
// unit (bmc_inst_*.cu, compiled in parallel); bmc_api.cu only sees these function pointers.
#pragma once
#include <string>
#include <cuda_runtime.h>
#include "bmc_kernels.cuh"

namespace bmc {

struct ModelVT {
  int n_var, n_c, vec, minb;
  int block, block_eager;  // threads per block of cycle_fn / cycle_eager_fn (one block per SM)
  int ct;              // 32-bit words per compartment-table row
  int n_read;          // property columns the pass loads (n_var minus the write-only ones): sizes the prefetch staging
  int dyn_shift;       // default work distribution of the particle pass (CycleParams::dyn_shift), DynTail below
  // kernel handles: addresses of __global__ instantiations, or cudaKernel_t of a JIT-compiled
  // user model; all launched with cudaLaunchKernel(handle, grid, block, {&params}, smem, stream)
  const void* cycle_fn;   // (CycleParams)  block = 256*minb threads, one block per SM; step-stamped ages (bmc_kernels.cuh)
  const void* cycle_eager_fn;  // same with float ages loaded, incremented and stored every step
  const void* pre_fn;     // (PreParams)    block = 256
  const void* init_fn;    // (InitParams)   block = 256
  const void* export_fn;  // (ExportParams) block = 256
  void* jit_library;      // cudaLibrary_t of a user model (unloaded with the context), else nullptr
};

// Work distribution of the particle pass (cycle_body): groups per warp >> shift are drawn dynamically at the end of the
// pass, the rest is grid-stride; 0 = every group dynamic.  Measured at 1.25e8 particles on a B200: the models whose
// pass is bound by instruction issue gain what the tickets cost (fixed_length 0.601 -> 0.553 ms per step,
// simple_acetate, 1.875e8 live particles, 1.99 -> 1.91 with an eighth dynamic), the HBM-bound monod pass LOSES with any static share (0.802 ->
// 0.838 with a quarter dynamic, 0.878 with an eighth): when the memory system is the limit the SMs do not progress at
// the same rate, and only the fully dynamic scheme keeps all of them busy until the end.
template <class M> struct DynTail { static constexpr int shift = 0; };
template <> struct DynTail<FixedLength> { static constexpr int shift = 3; };
template <> struct DynTail<SimpleAcetate> { static constexpr int shift = 3; };

// MINB = 256-thread units per block of the stamped-age kernel, MINB_E of the eager-age kernel (two more
// columns in flight per slot: more registers per thread, fewer threads)
template <class M, int VEC, int MINB, int MINB_E> static ModelVT make_vt() {
  ModelVT v;
  v.n_var = M::n_var; v.n_c = M::n_c; v.vec = VEC; v.minb = MINB;
  v.block = kBlock * MINB; v.block_eager = kBlock * MINB_E;
  v.ct = 1 + M::n_pre; v.n_read = ReadCols<M>::value; v.dyn_shift = DynTail<M>::shift;
  v.cycle_fn = (const void*)cycle_kernel<M, VEC, MINB, true>;
  v.cycle_eager_fn = (const void*)cycle_kernel<M, VEC, MINB_E, false>;
  v.pre_fn = (const void*)pre_step_kernel<M>;
  v.init_fn = (const void*)init_kernel<M>;
  v.export_fn = (const void*)export_kernel<M>;
  v.jit_library = nullptr;
  return v;
}

// Kernel variant = (slots per thread VEC, 256-thread units per block).  Every model ships its default and at most one
// alternative block size; BMC_VARIANT="v<VEC>b<MINB>" selects the alternative (tuning runs, and the parity tests that
// cover every instantiation in the library), anything else is ignored.
//   DEF = 4: 1024 threads x <= 64 registers (eager ages: 768 x <= 80), ALT = 3: 768 x <= 80
template <class M, int VEC, int DEF, int ALT = DEF> static bool pick_variant(const std::string& var, ModelVT& vt) {
  constexpr int DEF_E = DEF > 3 ? 3 : DEF, ALT_E = ALT > 3 ? 3 : ALT;
  if (ALT != DEF && (var == "alt" || (var.size() == 4 && var[0] == 'v' && var[2] == 'b' && var[1] - '0' == VEC && var[3] - '0' == ALT))) {
    vt = make_vt<M, VEC, ALT, ALT_E>();
    return true;
  }
  vt = make_vt<M, VEC, DEF, DEF_E>();
  return true;
}

bool pick_fixed_length(const std::string& var, ModelVT& vt);
bool pick_monod(const std::string& var, ModelVT& vt);
bool pick_simple_acetate(const std::string& var, ModelVT& vt);
bool pick_wide_udf_small(const std::string& var, int n_var, ModelVT& vt);   // P = 8, 16
bool pick_wide_udf_large(const std::string& var, int n_var, ModelVT& vt);   // P = 32, 64
// BMC_MODEL_UDF: NVRTC-compile a user model source against these headers (bmc_udf.cu)
bool load_udf_model(const char* source_path, ModelVT& vt, std::string& err);
void unload_udf_model(ModelVT& vt);

}  // namespace bmc
