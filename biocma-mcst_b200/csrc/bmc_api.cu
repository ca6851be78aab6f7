// C-ABI implementation (include/bmc.h) of the B200-native particle loop.
// Host side: context, device memory, launch sequencing of one cycleProcess.
// No CPU fallback exists: every entry point that computes runs CUDA kernels and
// reports BMC_ERR_CUDA if the device is unavailable.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <sched.h>
#include <new>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/bmc.h"
#include "bmc_kernels_common.cuh"
#include "bmc_model_vt.cuh"

using namespace bmc;

namespace {

static bool pick_model(int model, int n_var_udf, ModelVT& vt) {
  const char* v = getenv("BMC_VARIANT");
  const std::string var = v ? v : "";
  switch (model) {
    case BMC_MODEL_FIXED_LENGTH: return pick_fixed_length(var, vt);
    case BMC_MODEL_MONOD: return pick_monod(var, vt);
    case BMC_MODEL_SIMPLE_ACETATE: return pick_simple_acetate(var, vt);
    case BMC_MODEL_WIDE_UDF:
      return n_var_udf <= 16 ? pick_wide_udf_small(var, n_var_udf, vt) : pick_wide_udf_large(var, n_var_udf, vt);
    default: return false;
  }
}

}  // namespace

static int (*g_nccl_destroy)(void*) = nullptr;

struct bmc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int model = 0, n_var_udf = 0;
  ModelVT vt{};
  uint64_t n_species = 1, n_comp = 1;
  uint64_t seed = 0; uint32_t rank = 0;
  double allocation_factor = 1.5, buffer_ratio = 0.6, dead_ratio = 0.01;
  uint64_t min_removal = 0;
  // container
  size_t cap = 0, buf_cap = 0;
  float* props = nullptr; uint32_t* pos = nullptr; uint8_t* status = nullptr; float* age_hyd = nullptr; float* age_div = nullptr;
  float* buf_props = nullptr; uint32_t* buf_pos = nullptr; uint32_t* buf_mother = nullptr;
  uint32_t *div_mask = nullptr, *tile_off = nullptr, *blk_total = nullptr;
  uint32_t *tile_gap_off = nullptr, *tile_idle_off = nullptr, *blk_gap = nullptr, *blk_idle = nullptr, *src = nullptr;
  // domain
  int m = 0; bool domain_set = false; double table_dt = -1.0;
  double *d_vol = nullptr, *d_diag = nullptr, *d_cdf = nullptr;
  float *d_ctab = nullptr, *d_cdf_f = nullptr; uint32_t* d_neigh = nullptr;
  std::vector<bmc_leaving_flow> flows;
  // liquid
  double *d_conc = nullptr, *d_sources = nullptr, *d_acc = nullptr;
  unsigned long long* d_acc_fix = nullptr;  // fixed-point source accumulator (bmc_kernels.cuh, "Scatter")
  double *d_conc_next = nullptr, *d_mass = nullptr; bool mass_dirty = true;
  uint32_t *d_csc_ptr = nullptr, *d_csc_row = nullptr; double* d_csc_val = nullptr; bool transition_set = false;
  std::vector<bmc_feed> feeds; std::vector<double> h_vol;
  // gas phase (two-phase flow: second scalar field + gas-liquid mass transfer, bmc_gas_*)
  bool two_phase = false, gas_mass_dirty = true, mtr_set = false;
  double *d_gconc = nullptr, *d_gconc_next = nullptr, *d_gmass = nullptr, *d_gvol = nullptr, *d_kla = nullptr, *d_henry = nullptr, *d_mtr = nullptr;
  uint32_t *d_gcsc_ptr = nullptr, *d_gcsc_row = nullptr; double* d_gcsc_val = nullptr; size_t gcsc_cap = 0; bool gas_transition_set = false;
  std::vector<bmc_feed> gas_feeds;
  float weight = 1.0f;
  // state
  DevState* st = nullptr;
  DevState* h_st = nullptr;     // pinned landing buffer of sync_state
  // capacity policy (ensure_room): zero-copy mirror written by the commit thread of every step
  PinState* h_pin = nullptr; PinState* d_pin = nullptr;
  // zero-copy mirror of the published sources ({value, tag} records): bmc_get_sources reads it without a CUDA call
  unsigned long long* h_src_mirror = nullptr; unsigned long long* d_src_mirror = nullptr;
  uint64_t src_mirror_tag = 0;  // tag of the cycle whose sources are (or will be) in the mirror; 0 = mirror not current
  uint64_t host_step = 0;       // cycles enqueued since the particles were (re)loaded
  uint64_t recent_max_add = 0;  // decaying maximum of the newborns per step seen in the mirror
  uint64_t pin_seen_step = ~0ull, pin_history = 0;  // last mirrored step looked at, number of distinct ones since the (re)load
  double shrink_ratio = 0.0; bool exact_capacity = false; uint64_t n_regrow = 0;
  bool maybe_inactive = false;  // false only when the host KNOWS no slot is inactive (compaction kernels skipped)
  // launch config
  int n_sm = 148, grid_cycle = 148, blocks_per_sm = 1; size_t smem_bins = 0; int bins_in_smem = 0;
  uint64_t launches = 0;
  size_t queue_offset = 0, smem_total = 0, stage_offset = 0, stage_warp_bytes = 0; int ctab_in_smem = 0; size_t ctab_offset = 0; int grid_post = 148;
  int grid_cycle_eager = 148; size_t smem_eager = 0;
  bool fuse_post = true;  // whole step in one cooperative launch (BMC_FUSE_POST=0: particle pass + post_only_kernel)
  uint32_t dyn_min = 4, dyn_shift = 0;  // dynamically drawn tail of the particle pass (CycleParams; default from the model's launch table; BMC_DYN_MIN, BMC_DYN_SHIFT: tuning)
  // staging
  void* d_stage = nullptr; size_t stage_bytes = 0;
  // profiling
  int profile = 0;  // 0 = off, N = CUDA events around every N-th step kernel
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events; size_t prof_used = 0;  // pool, reused after bmc_profile_read
  double prof_ms = 0.0; uint64_t prof_n = 0;
  // step-stamped ages (bmc_kernels.cuh): valid while d_t / outlet configuration stay constant and
  // every age started at zero; otherwise the columns hold floats updated every step ("eager")
  bool lazy_ages = true; bool epoch_set = false; double epoch_dt = 0.0;
  uint64_t hyd_clock = 0;  // steps with an outlet so far: the clock of the hydraulic age (the division age ticks with host_step)
  float *d_tab_div = nullptr, *d_tab_hyd = nullptr; size_t tab_cap = 0;
  // pinned staging of the per-step host buffers (concentrations in, sources out)
  static constexpr int kPinRing = 4;
  double* h_pin_in[kPinRing] = {nullptr, nullptr, nullptr, nullptr}; cudaEvent_t ev_pin_in[kPinRing] = {nullptr, nullptr, nullptr, nullptr};
  int pin_next = 0; double* h_pin_out = nullptr;
  // pinned staging of flow-map updates (bmc_domain_update / bmc_liquid_set_transition): two slots, so that a new map
  // can be staged while the copy of the previous one may still be in flight; nothing synchronises the stream
  static constexpr int kMapRing = 2;
  unsigned char* h_map[kMapRing] = {nullptr, nullptr}; size_t h_map_bytes[kMapRing] = {0, 0};
  cudaEvent_t ev_map[kMapRing] = {nullptr, nullptr}; int map_next = 0;
  size_t csc_cap = 0;
  // nccl
  void* nccl_comm = nullptr; int nccl_ranks = 0;
  // peer-memory all-reduce (bmc_p2p_*)
  unsigned char* p2p_region = nullptr; size_t p2p_bytes = 0; int p2p_world = 0, p2p_rank = 0; bool p2p_on = false, p2p_ipc = false;
  unsigned char* p2p_base[kMaxPeers] = {}; unsigned long long p2p_epoch = 0;
  unsigned long long p2p_published = 0;  // epoch of the sources currently in d_sources if a cycle published them, else 0
  unsigned long long p2p_pending = 0;    // epoch whose all-reduce was requested and not finished yet (bmc_allreduce_sources is lazy)
  std::string err;
};

namespace {

#define CK(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                  \
      return (e__ == cudaErrorMemoryAllocation) ? BMC_ERR_NOMEM : BMC_ERR_CUDA;                        \
    }                                                                                                  \
  } while (0)

static size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

template <class T> static int dev_alloc(bmc_ctx* ctx, T** p, size_t n) {
  *p = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc((void**)p, n * sizeof(T)));
  return BMC_OK;
}
template <class T> static void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

static int check_launch(bmc_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return BMC_ERR_CUDA; }
  ctx->launches++;
  return BMC_OK;
}

static void free_container(bmc_ctx* c) {
  dev_free(c->props); dev_free(c->pos); dev_free(c->status); dev_free(c->age_hyd); dev_free(c->age_div);
  dev_free(c->buf_props); dev_free(c->buf_pos); dev_free(c->buf_mother);
  dev_free(c->div_mask); dev_free(c->tile_off);
  dev_free(c->tile_gap_off); dev_free(c->tile_idle_off); dev_free(c->src);
  c->cap = 0; c->buf_cap = 0;
}

// ParticlesContainer::_resize + __allocate_buffer__ (particles_container.hpp:601-643, 669-685):
// (re)allocate every column for `new_cap` slots, keeping the first `keep` slots.
static int resize_container(bmc_ctx* ctx, size_t new_cap, size_t keep) {
  new_cap = round_up(std::max<size_t>(new_cap, kTile), kTile);
  if (new_cap > 0xFFFFFFF0ull) { ctx->err = "capacity exceeds 2^32 slots per context"; return BMC_ERR_RANGE; }
  const int nv = ctx->vt.n_var;
  float* props = nullptr; uint32_t* pos = nullptr; uint8_t* status = nullptr; float *ah = nullptr, *ad = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &props, new_cap * nv))) return rc;
  if ((rc = dev_alloc(ctx, &pos, new_cap))) return rc;
  if ((rc = dev_alloc(ctx, &status, new_cap))) return rc;
  if ((rc = dev_alloc(ctx, &ah, new_cap))) return rc;
  if ((rc = dev_alloc(ctx, &ad, new_cap))) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaMemsetAsync(status, 0, new_cap, s));
  CK(cudaMemsetAsync(pos, 0, new_cap * 4, s));
  CK(cudaMemsetAsync(ah, 0, new_cap * 4, s));
  CK(cudaMemsetAsync(ad, 0, new_cap * 4, s));
  CK(cudaMemsetAsync(props, 0, new_cap * nv * 4, s));
  if (keep && ctx->props) {
    for (int k = 0; k < nv; ++k)
      CK(cudaMemcpyAsync(props + (size_t)k * new_cap, ctx->props + (size_t)k * ctx->cap, keep * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(pos, ctx->pos, keep * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(status, ctx->status, keep, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(ah, ctx->age_hyd, keep * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(ad, ctx->age_div, keep * 4, cudaMemcpyDeviceToDevice, s));
  }
  CK(cudaStreamSynchronize(s));
  free_container(ctx);
  ctx->props = props; ctx->pos = pos; ctx->status = status; ctx->age_hyd = ah; ctx->age_div = ad;
  ctx->cap = new_cap;
  // division buffer: ceil(buffer_ratio * n_allocated)
  ctx->buf_cap = round_up((size_t)std::ceil(ctx->buffer_ratio * (double)new_cap), 4);
  if ((rc = dev_alloc(ctx, &ctx->buf_props, ctx->buf_cap * nv))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->buf_pos, ctx->buf_cap))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->buf_mother, ctx->buf_cap))) return rc;
  const size_t n_tiles = new_cap / kTile;
  if ((rc = dev_alloc(ctx, &ctx->div_mask, new_cap / 32))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->tile_off, n_tiles))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->tile_gap_off, n_tiles))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->tile_idle_off, n_tiles))) return rc;
  if ((rc = dev_alloc(ctx, &ctx->src, new_cap))) return rc;
  CK(cudaMemsetAsync(ctx->div_mask, 0, new_cap / 32 * 4, s));
  CK(cudaMemsetAsync(ctx->tile_off, 0, n_tiles * 4, s));
  prepare_kernel<<<1, 32, 0, s>>>(ctx->st, (unsigned long long)ctx->cap, (unsigned long long)ctx->buf_cap, 0, ctx->allocation_factor,
                                  ctx->buffer_ratio, ctx->d_pin);  // room of the new capacity
  if ((rc = check_launch(ctx, "prepare"))) return rc;
  CK(cudaStreamSynchronize(s));
  return BMC_OK;
}

// next pinned staging slot of at least `bytes` (waits only for the copy that last read the slot, two updates ago)
static int map_stage(bmc_ctx* ctx, size_t bytes, unsigned char** out, int* slot_out) {
  const int slot = ctx->map_next; ctx->map_next = (ctx->map_next + 1) % bmc_ctx::kMapRing;
  if (!ctx->ev_map[slot]) CK(cudaEventCreateWithFlags(&ctx->ev_map[slot], cudaEventDisableTiming));
  else CK(cudaEventSynchronize(ctx->ev_map[slot]));
  if (ctx->h_map_bytes[slot] < bytes) {
    if (ctx->h_map[slot]) cudaFreeHost(ctx->h_map[slot]);
    ctx->h_map[slot] = nullptr; ctx->h_map_bytes[slot] = 0;
    CK(cudaMallocHost((void**)&ctx->h_map[slot], bytes + bytes / 4));
    ctx->h_map_bytes[slot] = bytes + bytes / 4;
  }
  *out = ctx->h_map[slot]; *slot_out = slot;
  return BMC_OK;
}

static int ensure_stage(bmc_ctx* ctx, size_t bytes) {
  if (ctx->stage_bytes >= bytes) return BMC_OK;
  if (ctx->d_stage) cudaFree(ctx->d_stage);
  ctx->d_stage = nullptr; ctx->stage_bytes = 0;
  CK(cudaMalloc(&ctx->d_stage, bytes));
  ctx->stage_bytes = bytes;
  return BMC_OK;
}

// pick the persistent grid and the shared-memory source bins for the cycle kernel
static int configure_launch(bmc_ctx* ctx) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->n_sm = prop.multiProcessorCount;
  // One block per SM: its dynamic shared memory holds the fp64 source bins (shared by all its warps),
  // the compartment table when it is small, and the staging buffers of the bulk-copy pipeline.
  const size_t smem_budget = (size_t)prop.sharedMemPerBlockOptin;
  const size_t static_reserve = 12 * 1024;  // static shared memory of the step kernel (post-cycle scratch, barriers)
  const size_t ctab_bytes = ctx->n_comp * (size_t)ctx->vt.ct * sizeof(float);
  const size_t bins_bytes = ctx->n_species * ctx->n_comp * sizeof(unsigned long long);
  // one deferred queue per warp (movers / outlet candidates waiting to be handled 32 at a time)
  const size_t queue_bytes = (size_t)(std::max(ctx->vt.block, ctx->vt.block_eager) / 32) * queue_entries(ctx->vt.vec) * sizeof(uint32_t);
  size_t room = smem_budget > static_reserve + queue_bytes + 1024 ? smem_budget - static_reserve - queue_bytes - 1024 : 0;
  // first the bins (block-private accumulation instead of L2 atomics per particle), then the table
  // (rebuilt by every block: no pre_step launch, gathers served from shared memory instead of L1/L2)
  ctx->bins_in_smem = (ctx->n_comp > 1 && bins_bytes <= room) ? 1 : 0;
  ctx->smem_bins = ctx->bins_in_smem ? bins_bytes : 0;
  room -= ctx->smem_bins;
  ctx->ctab_in_smem = (ctx->n_comp > 1 && ctab_bytes + 256 <= room) ? 1 : 0;
  if (const char* e = getenv("BMC_CTAB_SMEM")) ctx->ctab_in_smem = (ctx->ctab_in_smem && atoi(e) != 0) ? 1 : 0;  // tuning runs
  ctx->ctab_offset = (ctx->smem_bins + 15) / 16 * 16;
  ctx->queue_offset = (ctx->ctab_offset + (ctx->ctab_in_smem ? ctab_bytes : 0) + 127) / 128 * 128;
  ctx->smem_total = ctx->queue_offset + queue_bytes;
  if (ctx->smem_total > smem_budget) { ctx->err = "shared memory budget exceeded"; return BMC_ERR_UNSUPPORTED; }
  // prefetch staging (VEC == 4): one buffer per warp holding the columns of ONE group — pos, the read properties, both
  // ages (eager kernel) and the status bytes — when it fits beside bins, table and queues (BMC_PREFETCH=0 turns it off)
  ctx->stage_offset = 0; ctx->stage_warp_bytes = 0;
  {
    const char* e = getenv("BMC_PREFETCH");
    const size_t warp_bytes = (size_t)(1 + ctx->vt.n_read + 2) * 128 * (size_t)ctx->vt.vec + 128;
    const size_t stage_off = (ctx->smem_total + 127) / 128 * 128;
    const size_t total = stage_off + warp_bytes * (size_t)(std::max(ctx->vt.block, ctx->vt.block_eager) / 32);
    if (ctx->vt.vec == 4 && !(e && atoi(e) == 0) && total + static_reserve <= smem_budget) {
      ctx->stage_offset = stage_off; ctx->stage_warp_bytes = warp_bytes; ctx->smem_total = total;
    }
  }
  ctx->smem_eager = ctx->smem_total;
  ctx->grid_post = ctx->n_sm;           // cooperative launch: one block per SM is always co-resident
  if (const char* e = getenv("BMC_FUSE_POST")) ctx->fuse_post = atoi(e) != 0;
  ctx->dyn_shift = (uint32_t)ctx->vt.dyn_shift;
  if (const char* e = getenv("BMC_DYN_MIN")) ctx->dyn_min = (uint32_t)std::max(1, atoi(e));
  if (const char* e = getenv("BMC_DYN_SHIFT")) ctx->dyn_shift = (uint32_t)std::min(31, std::max(0, atoi(e)));
  const char* env = getenv("BMC_BLOCKS_PER_SM");
  auto grid_of = [&](const void* fn, int block, size_t smem, int& grid, int* occ_out) -> int {
    // always: static shared memory of the kernel counts against the 48 KB default as well
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, block, smem) != cudaSuccess) {
      (void)cudaGetLastError();
      occ = 1;  // JIT kernel handle not accepted by the occupancy query: trust __launch_bounds__(block, 1)
    } else if (occ < 1) {
      ctx->err = "step kernel does not fit on an SM (registers / shared memory)"; return BMC_ERR_UNSUPPORTED;
    }
    if (env && atoi(env) > 0) occ = std::min(occ, atoi(env));
    grid = std::min(ctx->n_sm * occ, kMaxGrid);
    if (occ_out) *occ_out = occ;
    return BMC_OK;
  };
  int rc;
  if ((rc = grid_of(ctx->vt.cycle_fn, ctx->vt.block, ctx->smem_total, ctx->grid_cycle, &ctx->blocks_per_sm))) return rc;
  if ((rc = grid_of(ctx->vt.cycle_eager_fn, ctx->vt.block_eager, ctx->smem_eager, ctx->grid_cycle_eager, nullptr))) return rc;
  if (getenv("BMC_VERBOSE")) {
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, ctx->vt.cycle_fn);
    fprintf(stderr, "[bmc] step kernel: grid %d (%d blocks/SM x %d SMs) x %d threads, %d regs, smem static %zu + dynamic %zu B (bins %zu, table %s, queues %zu, prefetch staging %zu per warp), eager grid %d\n",
            ctx->grid_cycle, ctx->blocks_per_sm, ctx->n_sm, ctx->vt.block, fa.numRegs, fa.sharedSizeBytes, ctx->smem_total, ctx->smem_bins,
            ctx->ctab_in_smem ? "smem" : "global", queue_bytes, ctx->stage_warp_bytes, ctx->grid_cycle_eager);
  }
  return BMC_OK;
}

static void fill_peer_exchange(const bmc_ctx* ctx, PeerExchange& x, unsigned long long publish, unsigned long long consume) {
  memset(&x, 0, sizeof(x));
  if (!ctx->p2p_on) return;
  x.world = ctx->p2p_world; x.rank = ctx->p2p_rank; x.publish_epoch = publish; x.consume_epoch = consume;
  x.spin_limit = 20000000000ll;  // ~10 s at 2 GHz
  for (int r = 0; r < ctx->p2p_world; ++r) x.base[r] = ctx->p2p_base[r];
}

// bmc_allreduce_sources is lazy on the peer-memory path: the sum is finished by the next step kernel (overlapped with
// its particle pass) or, when anything else touches the sources or synchronises first, by this small kernel
static int finish_pending_sum(bmc_ctx* ctx) {
  if (!ctx->p2p_pending) return BMC_OK;
  PeerExchange x;
  fill_peer_exchange(ctx, x, 0, ctx->p2p_pending);
  ctx->p2p_pending = 0; ctx->p2p_published = 0;  // d_sources now holds the global sum
  // ... which the host mirror of the last cycle does not: the kernel rewrites it with the summed values under a new tag,
  // so that bmc_get_sources can still read the result without a copy or a stream synchronisation
  ctx->src_mirror_tag = ctx->d_src_mirror ? ctx->launches + 1 : 0;
  p2p_exchange_kernel<<<1, 1024, 0, ctx->stream>>>(x, ctx->d_sources, (uint32_t)(ctx->n_species * ctx->n_comp), ctx->st, ctx->d_src_mirror,
                                                   ctx->src_mirror_tag);
  return check_launch(ctx, "p2p_exchange");
}

static int sync_state(bmc_ctx* ctx, DevState* out) {
  { int rc0 = finish_pending_sum(ctx); if (rc0) return rc0; }
  CK(cudaMemcpyAsync(ctx->h_st, ctx->st, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *out = *ctx->h_st;
  if (out->inactive == 0 && ctx->flows.empty()) ctx->maybe_inactive = false;
  if (out->error & 2u) { ctx->err = "compaction found fewer idle tail particles than gaps (inactive counter inconsistent)"; return BMC_ERR_INVALID; }
  if (out->error & 4u) { ctx->err = "peer-memory all-reduce: a peer did not publish its sources in time"; return BMC_ERR_NCCL; }
  if (out->error & kErrCapacity) {
    ctx->err = "capacity exhausted: a division was refused for lack of physical room, results differ from the reference from that step on "
               "(reserve more with bmc_reserve, or run with BMC_CAPACITY_MODE=exact)";
    return BMC_ERR_RANGE;
  }
  return BMC_OK;
}

// -----------------------------------------------------------------------------
// Capacity policy.  The reference resizes the container inside merge_buffer (particles_container.hpp:575-643):
// births are only ever limited by the division buffer, never by the capacity.  Here a step is enqueued without
// waiting for the previous one, so the arrays must already be large enough when the step runs.
//   * The device follows the reference's extents (n_allocated_elements, buffer extent) exactly in DevState
//     (logical_alloc / logical_buf) and limits the births of a step by the LOGICAL buffer, like the reference.
//   * The commit thread of every step writes {step, n_used, n_add, logical extents, error} into pinned host memory;
//     the host reads that mirror before every launch without any CUDA call.
//   * With no cycle in flight the counters are exact and one step can never add more than min(buffer, n_used)
//     particles: the arrays are always kept large enough for that, so a launch onto an idle stream is PROVABLY safe.
//   * Running ahead (at most kMaxAhead cycles) is allowed while the free room covers every cycle in flight at twice
//     the largest number of newborns seen recently; otherwise the host waits on the mirror (no CUDA call) until
//     enough cycles have committed — in the limit until the stream is idle, where it reallocates if it must.
//   * Should the physical room bind nevertheless (births more than doubled from one step to the next while the host
//     was ahead), the device sets kErrCapacity and every following call FAILS instead of silently counting an
//     Overflow the reference would not have had.  BMC_CAPACITY_MODE=exact never runs ahead: it cannot fail, at the
//     price of one launch latency per step.
// -----------------------------------------------------------------------------
constexpr uint64_t kMaxAhead = 4;

struct PinSnap { uint64_t step, n_used, n_add, la, lb; uint32_t error; };
// The mirror carries the low 32 bits of the step counter; the host knows the full value to within kMaxAhead + 1.
static uint64_t widen_step(const bmc_ctx* ctx, uint32_t low) {
  const uint64_t hs = ctx->host_step;
  uint64_t v = (hs & ~0xffffffffull) | low;
  if (v > hs) v -= 0x100000000ull;  // the device is never ahead of the host
  return v;
}
static PinSnap read_pin(const bmc_ctx* ctx) {
  const volatile unsigned* a = reinterpret_cast<const volatile unsigned*>(&ctx->h_pin->a);
  const volatile unsigned* b = reinterpret_cast<const volatile unsigned*>(&ctx->h_pin->b);
  for (;;) {  // each record arrives whole (one 16-byte write); a snapshot is consistent when both carry the same step
    unsigned ra[4], rb[4];
    for (int i = 0; i < 4; ++i) ra[i] = a[i];
    std::atomic_thread_fence(std::memory_order_acquire);
    for (int i = 0; i < 4; ++i) rb[i] = b[i];
    std::atomic_thread_fence(std::memory_order_acquire);
    if (ra[0] != rb[0] || a[0] != ra[0]) continue;
    PinSnap sn;
    sn.step = widen_step(ctx, ra[0]); sn.n_used = ra[1]; sn.n_add = ra[2]; sn.error = ra[3];
    sn.la = ((uint64_t)(rb[1] & 0xffffu) << 32) | rb[2]; sn.lb = ((uint64_t)(rb[1] >> 16) << 32) | rb[3];
    return sn;
  }
}
static uint64_t pin_step(const bmc_ctx* ctx) {
  return widen_step(ctx, reinterpret_cast<const volatile unsigned*>(&ctx->h_pin->a)[0]);
}

// wait on the pinned mirror (no CUDA synchronisation) until the device has committed a cycle beyond `seen`
static int wait_commit_beyond(bmc_ctx* ctx, uint64_t seen) {
  unsigned spins = 0;
  while (pin_step(ctx) <= seen) {
    if ((++spins & 0x3ffu) == 0u) {  // a fault on the device must not hang the host
      const cudaError_t q = cudaStreamQuery(ctx->stream);
      if (q == cudaSuccess) { if (pin_step(ctx) <= seen) { ctx->err = "device finished without committing the enqueued cycles"; return BMC_ERR_CUDA; } break; }
      if (q != cudaErrorNotReady) { ctx->err = std::string("cudaStreamQuery: ") + cudaGetErrorString(q); return BMC_ERR_CUDA; }
      sched_yield();
    }
  }
  return BMC_OK;
}

static uint64_t logical_alloc_after(const bmc_ctx* ctx, uint64_t la, uint64_t n) {  // _resize(n) on the logical extent
  return n > la ? (uint64_t)std::ceil((double)n * ctx->allocation_factor) : la;
}
static uint64_t worst_case_births(uint64_t n, uint64_t lb) { return std::min<uint64_t>(n, lb); }  // one division per particle, up to the buffer
static uint64_t predicted_births(uint64_t n, uint64_t recent) { return 2 * recent + n / 1024 + 64; }

// slots to allocate for a container (re)constructed with n particles: the reference's ceil(n * factor), and room for
// the worst case of one step
static size_t initial_capacity(const bmc_ctx* ctx, uint64_t n) {
  unsigned long long la = 0, lb = 0;
  if (n) logical_grow(n, ctx->allocation_factor, ctx->buffer_ratio, la, lb);
  return (size_t)std::max<uint64_t>(std::max<uint64_t>(la, n + worst_case_births(n, lb) + n / 8 + 1024), 1);
}

static int ensure_room(bmc_ctx* ctx) {
  int rc;
  for (;;) {
    const PinSnap sn = read_pin(ctx);
    if (sn.error & kErrCapacity) { DevState s; return sync_state(ctx, &s); }  // reports it
    if (sn.step != ctx->pin_seen_step) {  // history of the newborns per step (decaying maximum)
      ctx->recent_max_add = std::max<uint64_t>(sn.n_add, ctx->recent_max_add - ctx->recent_max_add / 8);
      ctx->pin_seen_step = sn.step; ctx->pin_history++;
    }
    const uint64_t d = ctx->host_step - sn.step;  // cycles in flight
    const uint64_t room = ctx->cap > sn.n_used ? ctx->cap - sn.n_used : 0;
    if (d == 0) {
      // idle stream: exact counters, provable bound
      if (room >= worst_case_births(sn.n_used, sn.lb) && sn.lb <= ctx->buf_cap && sn.la <= ctx->cap) return BMC_OK;
      DevState s;
      if ((rc = sync_state(ctx, &s))) return rc;
      const uint64_t n = s.n_used, pred = predicted_births(n, std::max<uint64_t>(s.n_add, ctx->recent_max_add));
      uint64_t want = n + worst_case_births(n, s.logical_buf) + n / 8 + 1024;
      want = std::max<uint64_t>(want, (uint64_t)std::ceil((double)(n + (kMaxAhead + 1) * pred) * ctx->allocation_factor));
      want = std::max<uint64_t>(want, s.logical_alloc);
      want = std::max<uint64_t>(want, (uint64_t)std::ceil((double)s.logical_buf / ctx->buffer_ratio));
      if (want <= ctx->cap) return BMC_OK;  // (only the pinned mirror was behind)
      want = std::max<uint64_t>(want, ctx->cap + ctx->cap / 4);  // geometric: a reallocation copies every particle
      ctx->n_regrow++;
      return resize_container(ctx, (size_t)want, (size_t)n);
    }
    if (!ctx->exact_capacity && d <= kMaxAhead && ctx->pin_history >= 2) {
      const uint64_t pred = predicted_births(sn.n_used, std::max<uint64_t>(sn.n_add, ctx->recent_max_add));
      const uint64_t la_next = logical_alloc_after(ctx, sn.la, sn.n_used + d * pred);  // a logical resize enlarges the logical buffer
      const uint64_t lb_next = std::max<uint64_t>(sn.lb, (uint64_t)std::ceil(ctx->buffer_ratio * (double)la_next));
      if (room >= (d + 1) * pred && lb_next <= ctx->buf_cap && la_next <= ctx->cap) return BMC_OK;
    }
    if ((rc = wait_commit_beyond(ctx, sn.step))) return rc;  // not covered: let the device catch up, then look again
  }
}

static void fill_post_params(bmc_ctx* ctx, PostParams& ip) {
  memset(&ip, 0, sizeof(ip));
  ip.props = ctx->props; ip.cap = ctx->cap; ip.n_var = ctx->vt.n_var; ip.pos = ctx->pos; ip.status = ctx->status;
  ip.age_hyd = ctx->age_hyd; ip.age_div = ctx->age_div; ip.st = ctx->st;
  ip.tile_gap_off = ctx->tile_gap_off; ip.tile_idle_off = ctx->tile_idle_off; ip.blk_gap = ctx->blk_gap; ip.blk_idle = ctx->blk_idle;
  ip.src = ctx->src;
  ip.buf_props = ctx->buf_props; ip.buf_stride = ctx->buf_cap; ip.buf_pos = ctx->buf_pos; ip.buf_mother = ctx->buf_mother;
  ip.div_mask = ctx->div_mask; ip.tile_off = ctx->tile_off; ip.blk_total = ctx->blk_total;
  ip.buf_cap = ctx->buf_cap;
  ip.tab_div = ctx->d_tab_div; ip.tab_hyd = ctx->d_tab_hyd;
  ip.acc = ctx->d_acc; ip.sources = ctx->d_sources; ip.n_bins = (uint32_t)(ctx->n_species * ctx->n_comp);
  ip.acc_fix = ctx->d_acc_fix; ip.weight = (double)ctx->weight; ip.n_species = (uint32_t)ctx->n_species; ip.n_c = ctx->vt.n_c;
  ip.vec = ctx->vt.vec;
  ip.min_removal = ctx->min_removal; ip.dead_ratio = ctx->dead_ratio;
  ip.allocation_factor = ctx->allocation_factor; ip.buffer_ratio = ctx->buffer_ratio; ip.shrink_ratio = ctx->shrink_ratio; ip.pin = ctx->d_pin;
}

// ---- step-stamped ages: host side ---------------------------------------------------------
static bool force_eager_ages() { const char* e = getenv("BMC_EAGER_AGES"); return e && atoi(e) != 0; }

// table entries [0, need) must exist; tab[0] = 0
static int ensure_age_tables(bmc_ctx* ctx, size_t need) {
  if (need <= ctx->tab_cap) return BMC_OK;
  const size_t new_cap = std::max<size_t>(std::max<size_t>(2 * ctx->tab_cap, need), 1u << 16);
  float *nd = nullptr, *nh = nullptr;
  int rc;
  if ((rc = dev_alloc(ctx, &nd, new_cap)) || (rc = dev_alloc(ctx, &nh, new_cap))) return rc;
  // everything on the context's stream: it is a non-blocking stream, so work issued on the legacy default stream
  // (cudaMemset is asynchronous for device memory) would not be ordered with what the caller enqueues next
  cudaStream_t s = ctx->stream;
  CK(cudaMemsetAsync(nd, 0, new_cap * 4, s)); CK(cudaMemsetAsync(nh, 0, new_cap * 4, s));
  if (ctx->tab_cap) {
    CK(cudaMemcpyAsync(nd, ctx->d_tab_div, ctx->tab_cap * 4, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(nh, ctx->d_tab_hyd, ctx->tab_cap * 4, cudaMemcpyDeviceToDevice, s));
  }
  CK(cudaStreamSynchronize(s));  // the old tables are freed next
  dev_free(ctx->d_tab_div); dev_free(ctx->d_tab_hyd);
  ctx->d_tab_div = nd; ctx->d_tab_hyd = nh; ctx->tab_cap = new_cap;
  return BMC_OK;
}

// start of an epoch (particles (re)loaded): stamps if every age is zero, floats otherwise
static int begin_age_epoch(bmc_ctx* ctx, bool lazy) {
  ctx->lazy_ages = lazy && !force_eager_ages();
  ctx->epoch_set = false; ctx->hyd_clock = 0;
  if (!ctx->lazy_ages) return BMC_OK;
  int rc;
  if ((rc = ensure_age_tables(ctx, 2))) return rc;
  CK(cudaMemsetAsync(ctx->d_tab_div, 0, 8, ctx->stream));
  CK(cudaMemsetAsync(ctx->d_tab_hyd, 0, 8, ctx->stream));
  return BMC_OK;
}

// d_t / outlets changed (or the caller asked for it): turn the stamps into the floats they stand
// for, in place, and continue with the eager kernel
static int make_ages_eager(bmc_ctx* ctx) {
  if (!ctx->lazy_ages) return BMC_OK;
  if (ctx->cap) {
    const unsigned grid = (unsigned)((ctx->cap + 255) / 256);
    ages_to_eager_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->age_div, ctx->d_tab_div, (uint32_t)ctx->host_step, ctx->cap);
    int rc;
    if ((rc = check_launch(ctx, "ages_to_eager"))) return rc;
    ages_to_eager_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->age_hyd, ctx->d_tab_hyd, (uint32_t)ctx->hyd_clock, ctx->cap);
    if ((rc = check_launch(ctx, "ages_to_eager"))) return rc;
  }
  ctx->lazy_ages = false;
  return BMC_OK;
}

}  // namespace

// =============================================================================
extern "C" {

int bmc_create(bmc_ctx** out, const bmc_config* cfg) {
  if (!out || !cfg) return BMC_ERR_INVALID;
  *out = nullptr;
  if (cfg->n_species == 0 || cfg->n_compartments == 0 || cfg->n_compartments > 0xFFFFFFF0ull) return BMC_ERR_INVALID;
  bmc_ctx* ctx = new (std::nothrow) bmc_ctx();
  if (!ctx) return BMC_ERR_NOMEM;
  auto fail = [&](int rc) { fprintf(stderr, "bmc_create: %s\n", ctx->err.c_str()); bmc_ctx* t = ctx; bmc_destroy(&t); return rc; };
  {
    cudaError_t e0 = cudaSetDevice(cfg->device);
    if (e0 != cudaSuccess) { ctx->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e0); return fail(BMC_ERR_CUDA); }
  }
  if (cfg->model == BMC_MODEL_UDF) {
    // `-mn udf_model` + BIOMC_LIB_UDF (apps/api/src/udf_handle.cpp:22-39): the path names the model SOURCE
    const char* path = cfg->udf_source_path ? cfg->udf_source_path : getenv("BIOMC_LIB_UDF");
    if (!path) { ctx->err = "BMC_MODEL_UDF needs udf_source_path or BIOMC_LIB_UDF"; return fail(BMC_ERR_INVALID); }
    if (!load_udf_model(path, ctx->vt, ctx->err)) return fail(BMC_ERR_UNSUPPORTED);
  } else if (!pick_model(cfg->model, cfg->n_var_udf, ctx->vt)) { ctx->err = "unknown model / unsupported n_var_udf"; return fail(BMC_ERR_INVALID); }
  if ((uint64_t)ctx->vt.n_c > cfg->n_species) { ctx->err = "model n_c exceeds n_species"; return fail(BMC_ERR_INVALID); }
  ctx->device = cfg->device; ctx->model = cfg->model; ctx->n_var_udf = cfg->n_var_udf;
  ctx->n_species = cfg->n_species; ctx->n_comp = cfg->n_compartments;
  ctx->seed = cfg->seed; ctx->rank = cfg->rank;
  if (cfg->allocation_factor > 0) ctx->allocation_factor = std::max(1.0, cfg->allocation_factor);
  if (cfg->buffer_ratio > 0) ctx->buffer_ratio = std::min(1.0, cfg->buffer_ratio);
  if (cfg->dead_particle_ratio_threshold > 0) ctx->dead_ratio = cfg->dead_particle_ratio_threshold;
  ctx->min_removal = cfg->minimum_dead_particle_removal;
  if (cfg->shrink_ratio > 0) ctx->shrink_ratio = cfg->shrink_ratio;
  cudaError_t e = cudaSetDevice(ctx->device);
  if (e != cudaSuccess) { ctx->err = std::string("cudaSetDevice: ") + cudaGetErrorString(e); return fail(BMC_ERR_CUDA); }
  auto ck = [&](cudaError_t e2, const char* w) { if (e2 != cudaSuccess) { ctx->err = std::string(w) + ": " + cudaGetErrorString(e2); return false; } return true; };
  if (!ck(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return fail(BMC_ERR_CUDA);
  if (!ck(cudaMalloc((void**)&ctx->st, sizeof(DevState)), "cudaMalloc state")) return fail(BMC_ERR_NOMEM);
  if (!ck(cudaMemset(ctx->st, 0, sizeof(DevState)), "memset state")) return fail(BMC_ERR_CUDA);
  if (!ck(cudaMallocHost((void**)&ctx->h_st, sizeof(DevState)), "cudaMallocHost")) return fail(BMC_ERR_NOMEM);
  if (!ck(cudaHostAlloc((void**)&ctx->h_pin, sizeof(PinState), cudaHostAllocMapped), "cudaHostAlloc")) return fail(BMC_ERR_NOMEM);
  memset(ctx->h_pin, 0, sizeof(PinState));
  if (!ck(cudaHostGetDevicePointer((void**)&ctx->d_pin, ctx->h_pin, 0), "cudaHostGetDevicePointer")) return fail(BMC_ERR_CUDA);
  {
    const size_t nbm = (size_t)cfg->n_species * (size_t)cfg->n_compartments;
    if (!ck(cudaHostAlloc((void**)&ctx->h_src_mirror, nbm * 16, cudaHostAllocMapped), "cudaHostAlloc")) return fail(BMC_ERR_NOMEM);
    memset(ctx->h_src_mirror, 0, nbm * 16);
    if (!ck(cudaHostGetDevicePointer((void**)&ctx->d_src_mirror, ctx->h_src_mirror, 0), "cudaHostGetDevicePointer")) return fail(BMC_ERR_CUDA);
  }
  if (const char* e = getenv("BMC_CAPACITY_MODE")) ctx->exact_capacity = std::string(e) == "exact";
  const size_t nb = ctx->n_species * ctx->n_comp;
  int rc;
  if ((rc = dev_alloc(ctx, &ctx->d_conc, nb)) || (rc = dev_alloc(ctx, &ctx->d_sources, nb)) || (rc = dev_alloc(ctx, &ctx->d_acc, nb)) ||
      (rc = dev_alloc(ctx, &ctx->d_acc_fix, nb)) ||
      (rc = dev_alloc(ctx, &ctx->d_conc_next, nb)) || (rc = dev_alloc(ctx, &ctx->d_mass, nb)) || (rc = dev_alloc(ctx, &ctx->d_csc_ptr, ctx->n_comp + 1)) ||
      (rc = dev_alloc(ctx, &ctx->d_vol, ctx->n_comp)) || (rc = dev_alloc(ctx, &ctx->d_diag, ctx->n_comp)) ||
      (rc = dev_alloc(ctx, &ctx->d_ctab, ctx->n_comp * (size_t)ctx->vt.ct)) || (rc = dev_alloc(ctx, &ctx->blk_total, kMaxGrid + 1)) ||
      (rc = dev_alloc(ctx, &ctx->blk_gap, kMaxGrid + 1)) || (rc = dev_alloc(ctx, &ctx->blk_idle, kMaxGrid + 1)))
    return fail(rc);
  cudaMemset(ctx->d_conc, 0, nb * 8); cudaMemset(ctx->d_sources, 0, nb * 8); cudaMemset(ctx->d_acc, 0, nb * 8);
  cudaMemset(ctx->d_acc_fix, 0, nb * 8);
  for (int i = 0; i < bmc_ctx::kPinRing; ++i) {
    if (!ck(cudaMallocHost((void**)&ctx->h_pin_in[i], nb * 8), "cudaMallocHost")) return fail(BMC_ERR_NOMEM);
    if (!ck(cudaEventCreateWithFlags(&ctx->ev_pin_in[i], cudaEventDisableTiming), "cudaEventCreate")) return fail(BMC_ERR_CUDA);
  }
  if (!ck(cudaMallocHost((void**)&ctx->h_pin_out, nb * 8), "cudaMallocHost")) return fail(BMC_ERR_NOMEM);
  cudaMemset(ctx->d_mass, 0, nb * 8); cudaMemset(ctx->d_csc_ptr, 0, (ctx->n_comp + 1) * 4);
  ctx->h_vol.assign(ctx->n_comp, 1.0);
  cudaMemset(ctx->d_ctab, 0, ctx->n_comp * (size_t)ctx->vt.ct * 4);
  cudaMemset(ctx->d_vol, 0, ctx->n_comp * 8); cudaMemset(ctx->d_diag, 0, ctx->n_comp * 8);
  cudaMemset(ctx->blk_total, 0, (kMaxGrid + 1) * 4);
  if ((rc = configure_launch(ctx))) return fail(rc);
  if (cfg->capacity) { if ((rc = resize_container(ctx, cfg->capacity, 0))) return fail(rc); }
  // the clears above went to the legacy default stream and are asynchronous for device memory; the context's stream
  // is non-blocking, so they must have completed before the caller's first enqueue (e.g. bmc_set_concentrations)
  if (!ck(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return fail(BMC_ERR_CUDA);
  *out = ctx;
  return BMC_OK;
}

int bmc_destroy(bmc_ctx** h) {
  if (!h || !*h) return BMC_ERR_INVALID;
  bmc_ctx* c = *h;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->nccl_comm && g_nccl_destroy) g_nccl_destroy(c->nccl_comm);
  if (c->p2p_ipc) for (int r = 0; r < c->p2p_world; ++r) if (r != c->p2p_rank && c->p2p_base[r]) cudaIpcCloseMemHandle(c->p2p_base[r]);
  if (c->p2p_region) cudaFree(c->p2p_region);
  unload_udf_model(c->vt);
  free_container(c);
  dev_free(c->d_conc); dev_free(c->d_sources); dev_free(c->d_acc); dev_free(c->d_acc_fix); dev_free(c->d_conc_next); dev_free(c->d_mass);
  dev_free(c->d_csc_ptr); dev_free(c->d_csc_row); dev_free(c->d_csc_val);
  dev_free(c->d_gconc); dev_free(c->d_gconc_next); dev_free(c->d_gmass); dev_free(c->d_gvol); dev_free(c->d_kla); dev_free(c->d_henry); dev_free(c->d_mtr);
  dev_free(c->d_gcsc_ptr); dev_free(c->d_gcsc_row); dev_free(c->d_gcsc_val);
  dev_free(c->d_vol); dev_free(c->d_diag); dev_free(c->d_cdf);
  dev_free(c->d_ctab); dev_free(c->d_cdf_f); dev_free(c->d_neigh);
  dev_free(c->blk_total); dev_free(c->blk_gap); dev_free(c->blk_idle);
  dev_free(c->d_tab_div); dev_free(c->d_tab_hyd);
  dev_free(c->st);
  if (c->d_stage) cudaFree(c->d_stage);
  if (c->h_st) cudaFreeHost(c->h_st);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_src_mirror) cudaFreeHost(c->h_src_mirror);
  for (int i = 0; i < bmc_ctx::kPinRing; ++i) { if (c->h_pin_in[i]) cudaFreeHost(c->h_pin_in[i]); if (c->ev_pin_in[i]) cudaEventDestroy(c->ev_pin_in[i]); }
  if (c->h_pin_out) cudaFreeHost(c->h_pin_out);
  for (int i = 0; i < bmc_ctx::kMapRing; ++i) { if (c->h_map[i]) cudaFreeHost(c->h_map[i]); if (c->ev_map[i]) cudaEventDestroy(c->ev_map[i]); }
  for (auto& pr : c->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  *h = nullptr;
  return BMC_OK;
}

const char* bmc_last_error(const bmc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int bmc_model_dims(const bmc_ctx* ctx, int32_t* n_var, int32_t* n_c) {
  if (!ctx) return BMC_ERR_INVALID;
  if (n_var) *n_var = ctx->vt.n_var;
  if (n_c) *n_c = ctx->vt.n_c;
  return BMC_OK;
}

int bmc_reserve(bmc_ctx* ctx, uint64_t capacity) {
  if (!ctx) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  DevState s; int rc;
  if ((rc = sync_state(ctx, &s))) return rc;
  if (capacity <= ctx->cap) return BMC_OK;
  return resize_container(ctx, capacity, (size_t)s.n_used);
}

int bmc_set_particles(bmc_ctx* ctx, uint64_t n, const float* props, const uint64_t* position, const uint8_t* status,
                      const float* age_h, const float* age_d) {
  if (!ctx || (n && !props)) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  int rc;
  const size_t want = initial_capacity(ctx, n);
  if (want > ctx->cap || ctx->cap == 0) { if ((rc = resize_container(ctx, want, 0))) return rc; }
  cudaStream_t s = ctx->stream;
  const int nv = ctx->vt.n_var;
  for (int k = 0; k < nv; ++k)
    CK(cudaMemcpyAsync(ctx->props + (size_t)k * ctx->cap, props + (size_t)k * n, n * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(ctx->st, 0, sizeof(DevState), s));
  if (position) {
    const size_t chunk = 1u << 22;
    if ((rc = ensure_stage(ctx, chunk * 8))) return rc;
    for (size_t o = 0; o < n; o += chunk) {
      const size_t c = std::min<size_t>(chunk, n - o);
      CK(cudaMemcpyAsync(ctx->d_stage, position + o, c * 8, cudaMemcpyHostToDevice, s));
      pos_narrow_kernel<<<(unsigned)((c + 255) / 256), 256, 0, s>>>((const unsigned long long*)ctx->d_stage, ctx->pos + o, c,
                                                                    (uint32_t)ctx->n_comp, &ctx->st->error);
      if ((rc = check_launch(ctx, "pos_narrow"))) return rc;
      CK(cudaStreamSynchronize(s));
    }
  } else {
    CK(cudaMemsetAsync(ctx->pos, 0, n * 4, s));
  }
  if (status) CK(cudaMemcpyAsync(ctx->status, status, n, cudaMemcpyHostToDevice, s)); else CK(cudaMemsetAsync(ctx->status, 0, n, s));
  // slots beyond n must be Idle (appended newborns rely on it)
  if (ctx->cap > n) CK(cudaMemsetAsync(ctx->status + n, 0, ctx->cap - n, s));
  {
    auto all_zero = [n](const float* a) { if (a) for (uint64_t i = 0; i < n; ++i) if (a[i] != 0.0f) return false; return true; };
    if ((rc = begin_age_epoch(ctx, all_zero(age_h) && all_zero(age_d)))) return rc;
    if (ctx->lazy_ages) {  // every age is zero: step stamps (0 for idle particles, frozen for the others)
      if (n) {
        ages_init_stamps_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ctx->status, ctx->age_hyd, ctx->age_div, n);
        if ((rc = check_launch(ctx, "ages_init_stamps"))) return rc;
      }
    } else {
      if (age_h) CK(cudaMemcpyAsync(ctx->age_hyd, age_h, n * 4, cudaMemcpyHostToDevice, s)); else CK(cudaMemsetAsync(ctx->age_hyd, 0, n * 4, s));
      if (age_d) CK(cudaMemcpyAsync(ctx->age_div, age_d, n * 4, cudaMemcpyHostToDevice, s)); else CK(cudaMemsetAsync(ctx->age_div, 0, n * 4, s));
    }
  }
  if (status && n) {
    count_inactive_kernel<<<std::min<unsigned>(1024, (unsigned)((n + 255) / 256)), 256, 0, s>>>(ctx->status, n, ctx->st);
    if ((rc = check_launch(ctx, "count_inactive"))) return rc;
  }
  CK(cudaMemcpyAsync(&ctx->st->n_used, &n, 8, cudaMemcpyHostToDevice, s));
  prepare_kernel<<<1, 32, 0, s>>>(ctx->st, (unsigned long long)ctx->cap, (unsigned long long)ctx->buf_cap, 1, ctx->allocation_factor,
                                  ctx->buffer_ratio, ctx->d_pin);
  if ((rc = check_launch(ctx, "prepare"))) return rc;
  DevState hs;
  if ((rc = sync_state(ctx, &hs))) return rc;
  if (hs.error & 1u) { ctx->err = "particle position out of range"; return BMC_ERR_RANGE; }
  ctx->maybe_inactive = hs.inactive != 0;
  ctx->host_step = 0; ctx->recent_max_add = 0; ctx->pin_seen_step = ~0ull; ctx->pin_history = 0;
  return BMC_OK;
}

int bmc_get_particles(bmc_ctx* ctx, uint64_t n, float* props, uint64_t* position, uint8_t* status, float* age_h, float* age_d) {
  if (!ctx) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  DevState hs; int rc;
  if ((rc = sync_state(ctx, &hs))) return rc;
  if (n > hs.n_used) { ctx->err = "bmc_get_particles: n exceeds n_used"; return BMC_ERR_RANGE; }
  cudaStream_t s = ctx->stream;
  const int nv = ctx->vt.n_var;
  if (props) for (int k = 0; k < nv; ++k)
    CK(cudaMemcpyAsync(props + (size_t)k * n, ctx->props + (size_t)k * ctx->cap, n * 4, cudaMemcpyDeviceToHost, s));
  if (position) {
    const size_t chunk = 1u << 22;
    if ((rc = ensure_stage(ctx, chunk * 8))) return rc;
    for (size_t o = 0; o < n; o += chunk) {
      const size_t c = std::min<size_t>(chunk, n - o);
      pos_widen_kernel<<<(unsigned)((c + 255) / 256), 256, 0, s>>>(ctx->pos + o, (unsigned long long*)ctx->d_stage, c);
      if ((rc = check_launch(ctx, "pos_widen"))) return rc;
      CK(cudaMemcpyAsync(position + o, ctx->d_stage, c * 8, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    }
  }
  if (status) CK(cudaMemcpyAsync(status, ctx->status, n, cudaMemcpyDeviceToHost, s));
  for (int a = 0; a < 2; ++a) {
    float* out = a ? age_d : age_h;
    const float* col = a ? ctx->age_div : ctx->age_hyd;
    if (!out) continue;
    if (!ctx->lazy_ages) { CK(cudaMemcpyAsync(out, col, n * 4, cudaMemcpyDeviceToHost, s)); continue; }
    const size_t chunk = 1u << 23;  // stamps -> floats through the staging buffer
    if ((rc = ensure_stage(ctx, std::min<size_t>(chunk, std::max<uint64_t>(n, 1)) * 4))) return rc;
    for (size_t o = 0; o < n; o += chunk) {
      const size_t c = std::min<size_t>(chunk, n - o);
      ages_read_kernel<<<(unsigned)((c + 255) / 256), 256, 0, s>>>(col + o, a ? ctx->d_tab_div : ctx->d_tab_hyd, (uint32_t)(a ? ctx->host_step : ctx->hyd_clock),
                                                                  (float*)ctx->d_stage, c);
      if ((rc = check_launch(ctx, "ages_read"))) return rc;
      CK(cudaMemcpyAsync(out + o, ctx->d_stage, c * 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
    }
  }
  CK(cudaStreamSynchronize(s));
  return BMC_OK;
}

int bmc_init_particles(bmc_ctx* ctx, uint64_t n, int uniform_position, const float* linit, double* total_mass) {
  if (!ctx || n == 0) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  int rc;
  const size_t want = initial_capacity(ctx, n);
  if (want > ctx->cap || ctx->cap == 0) { if ((rc = resize_container(ctx, want, 0))) return rc; }
  cudaStream_t s = ctx->stream;
  float* d_linit = nullptr;
  if (linit) {
    if ((rc = ensure_stage(ctx, n * 4))) return rc;
    d_linit = (float*)ctx->d_stage;
    CK(cudaMemcpyAsync(d_linit, linit, n * 4, cudaMemcpyHostToDevice, s));
  }
  CK(cudaMemsetAsync(ctx->st, 0, sizeof(DevState), s));
  CK(cudaMemsetAsync(ctx->status, 0, ctx->cap, s));
  if ((rc = begin_age_epoch(ctx, true))) return rc;  // init writes zero ages = stamp 0
  const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->n_sm * 8);
  InitParams ipar{ctx->props, ctx->cap, ctx->pos, ctx->status, ctx->age_hyd, ctx->age_div, n,
                  uniform_position ? (uint32_t)ctx->n_comp : 1u, d_linit, (uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32), ctx->rank, ctx->st};
  void* iargs[] = {&ipar};
  CK(cudaLaunchKernel(ctx->vt.init_fn, dim3(grid), dim3(256), iargs, 0, s));
  if ((rc = check_launch(ctx, "init_kernel"))) return rc;
  CK(cudaMemcpyAsync(&ctx->st->n_used, &n, 8, cudaMemcpyHostToDevice, s));
  prepare_kernel<<<1, 32, 0, s>>>(ctx->st, (unsigned long long)ctx->cap, (unsigned long long)ctx->buf_cap, 1, ctx->allocation_factor,
                                  ctx->buffer_ratio, ctx->d_pin);
  if ((rc = check_launch(ctx, "prepare"))) return rc;
  DevState hs;
  if ((rc = sync_state(ctx, &hs))) return rc;
  if (total_mass) *total_mass = hs.init_mass;
  ctx->maybe_inactive = false;
  ctx->host_step = 0; ctx->recent_max_add = 0; ctx->pin_seen_step = ~0ull; ctx->pin_history = 0;
  return BMC_OK;
}

int bmc_set_weight(bmc_ctx* ctx, double w) {
  if (!ctx || !(w > 0)) return BMC_ERR_INVALID;
  ctx->weight = (float)w;  // deep_copy(container.weights, new_weight): float view
  return BMC_OK;
}

int bmc_domain_update(bmc_ctx* ctx, const double* volumes, const uint64_t* neighbors_flat, const double* out_flows,
                      const double* proba_flat, uint64_t n_cols) {
  if (!ctx || !volumes || !out_flows) return BMC_ERR_INVALID;
  if (ctx->n_comp > 1 && (!neighbors_flat || !proba_flat || n_cols == 0)) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  // Stream-ordered, no host synchronisation: the arrays are staged in pinned memory and copied on the context's
  // stream, i.e. after the cycles already enqueued (which read the old tables) and before the ones that follow.
  // A flow-map switch of a transitioner (host_specific.cpp:81-87, 263-266) therefore costs a few small copies.
  const size_t nc = ctx->n_comp, nn = nc * n_cols;
  for (size_t i = 0; i < nn; ++i)
    if (neighbors_flat[i] >= nc) { ctx->err = "neighbor index out of range"; return BMC_ERR_RANGE; }
  int rc;
  if (nn && (int)n_cols != ctx->m) {  // first map, or a map with another neighbour count: the tables change size
    CK(cudaStreamSynchronize(s));
    dev_free(ctx->d_cdf); dev_free(ctx->d_cdf_f); dev_free(ctx->d_neigh);
    if ((rc = dev_alloc(ctx, &ctx->d_cdf, nn)) || (rc = dev_alloc(ctx, &ctx->d_cdf_f, nn)) || (rc = dev_alloc(ctx, &ctx->d_neigh, nn))) return rc;
    ctx->m = (int)n_cols;
  }
  const size_t o_vol = 0, o_diag = nc * 8, o_cdf = 2 * nc * 8, o_nb = o_cdf + nn * 8, bytes = o_nb + nn * 4;
  unsigned char* h = nullptr; int slot = 0;
  if ((rc = map_stage(ctx, bytes, &h, &slot))) return rc;
  memcpy(h + o_vol, volumes, nc * 8);
  memcpy(h + o_diag, out_flows, nc * 8);
  if (nn) {
    memcpy(h + o_cdf, proba_flat, nn * 8);
    uint32_t* nb = reinterpret_cast<uint32_t*>(h + o_nb);
    for (size_t i = 0; i < nn; ++i) nb[i] = (uint32_t)neighbors_flat[i];
  }
  ctx->h_vol.assign(volumes, volumes + nc);
  CK(cudaMemcpyAsync(ctx->d_vol, h + o_vol, nc * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ctx->d_diag, h + o_diag, nc * 8, cudaMemcpyHostToDevice, s));
  if (nn) {
    CK(cudaMemcpyAsync(ctx->d_cdf, h + o_cdf, nn * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_neigh, h + o_nb, nn * 4, cudaMemcpyHostToDevice, s));
    derive_cdf_table_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(ctx->d_cdf, ctx->d_cdf_f, nn);
    if ((rc = check_launch(ctx, "derive_cdf"))) return rc;
  }
  CK(cudaEventRecord(ctx->ev_map[slot], s));
  ctx->domain_set = true;
  ctx->table_dt = -1.0;    // leave table depends on dt: rebuilt by the next cycle
  return BMC_OK;
}

int bmc_set_leaving_flows(bmc_ctx* ctx, uint64_t n, const bmc_leaving_flow* f) {
  if (!ctx || (n && !f)) return BMC_ERR_INVALID;
  if (n > (uint64_t)kMaxFlows) { ctx->err = "too many leaving flows (max 16)"; return BMC_ERR_UNSUPPORTED; }
  for (uint64_t i = 0; i < n; ++i) if (f[i].index >= ctx->n_comp) { ctx->err = "leaving flow index out of range"; return BMC_ERR_RANGE; }
  ctx->flows.assign(f, f + n);
  return BMC_OK;
}

int bmc_set_concentrations(bmc_ctx* ctx, const double* c) {
  if (!ctx || !c) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  // The caller's buffer is copied into a pinned ring slot (free on return, api_raw.cpp:328-379 convention);
  // the H2D copy is then truly asynchronous and ordered with the enqueued cycles by the stream.
  const size_t bytes = ctx->n_species * ctx->n_comp * 8;
  const int slot = ctx->pin_next; ctx->pin_next = (ctx->pin_next + 1) % bmc_ctx::kPinRing;
  CK(cudaEventSynchronize(ctx->ev_pin_in[slot]));  // the copy that last used this slot has executed
  memcpy(ctx->h_pin_in[slot], c, bytes);
  CK(cudaMemcpyAsync(ctx->d_conc, ctx->h_pin_in[slot], bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaEventRecord(ctx->ev_pin_in[slot], ctx->stream));
  ctx->mass_dirty = true;  // total_mass = C * V is rebuilt by the next bmc_liquid_step
  return BMC_OK;
}

int bmc_get_concentrations(bmc_ctx* ctx, double* out) {
  if (!ctx || !out) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_conc, ctx->n_species * ctx->n_comp * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BMC_OK;
}

// COO -> CSC of a transition matrix into (ptr, row, val) device arrays, staged in pinned memory and copied on the
// context's stream (stable: COO order kept per column, so every element accumulates its inflow terms in the order a
// sequential COO sweep does).  No host synchronisation unless the arrays have to grow.
static int upload_transition(bmc_ctx* ctx, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals,
                             uint32_t* d_ptr, uint32_t** d_row, double** d_val, size_t* cap) {
  const size_t nc = ctx->n_comp;
  for (uint64_t e = 0; e < nnz; ++e)
    if (rows[e] >= nc || cols[e] >= nc) { ctx->err = "transition index out of range"; return BMC_ERR_RANGE; }
  cudaStream_t s = ctx->stream;
  int rc;
  if (nnz > *cap || !*d_row) {  // more non-zeros than any map before: larger arrays (the only case that waits for the stream)
    CK(cudaStreamSynchronize(s));
    dev_free(*d_row); dev_free(*d_val);
    const size_t c = (size_t)nnz + (size_t)nnz / 4;
    if ((rc = dev_alloc(ctx, d_row, c)) || (rc = dev_alloc(ctx, d_val, c))) return rc;
    *cap = c;
  }
  const size_t o_ptr = 0, o_val = ((nc + 1) * 4 + 7) / 8 * 8, o_row = o_val + nnz * 8, bytes = o_row + nnz * 4;
  unsigned char* h = nullptr; int slot = 0;
  if ((rc = map_stage(ctx, bytes, &h, &slot))) return rc;
  uint32_t* ptr = reinterpret_cast<uint32_t*>(h + o_ptr);
  double* val = reinterpret_cast<double*>(h + o_val);
  uint32_t* row = reinterpret_cast<uint32_t*>(h + o_row);
  for (size_t j = 0; j <= nc; ++j) ptr[j] = 0;
  for (uint64_t e = 0; e < nnz; ++e) ptr[cols[e] + 1]++;
  for (size_t j = 0; j < nc; ++j) ptr[j + 1] += ptr[j];
  std::vector<uint32_t> fill(ptr, ptr + nc);
  for (uint64_t e = 0; e < nnz; ++e) { const uint32_t d = fill[cols[e]]++; row[d] = (uint32_t)rows[e]; val[d] = vals[e]; }
  // The reference builds an Eigen::SparseMatrix from the triplets (sparse_from_coo, implScalar.cpp:79-129): duplicate
  // entries are summed and the rows of a column end up in ascending order, which is the order `c * m_transition` adds
  // the inflow terms of an element in.  Same here: every column sorted by row (stable), duplicates merged in input order.
  {
    std::vector<std::pair<uint32_t, double>> col;
    uint32_t out = 0, begin = 0;
    for (size_t j = 0; j < nc; ++j) {
      const uint32_t end = ptr[j + 1];
      col.clear();
      for (uint32_t e = begin; e < end; ++e) col.emplace_back(row[e], val[e]);
      std::stable_sort(col.begin(), col.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      ptr[j] = out;
      for (size_t e = 0; e < col.size(); ++e) {
        if (e > 0 && col[e].first == col[e - 1].first) val[out - 1] += col[e].second;
        else { row[out] = col[e].first; val[out] = col[e].second; ++out; }
      }
      begin = end;
    }
    ptr[nc] = out;
    nnz = out;
  }
  CK(cudaMemcpyAsync(d_ptr, ptr, (nc + 1) * 4, cudaMemcpyHostToDevice, s));
  if (nnz) {
    CK(cudaMemcpyAsync(*d_row, row, nnz * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(*d_val, val, nnz * 8, cudaMemcpyHostToDevice, s));
  }
  CK(cudaEventRecord(ctx->ev_map[slot], s));
  return BMC_OK;
}

int bmc_liquid_set_transition(bmc_ctx* ctx, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals) {
  if (!ctx || (nnz && (!rows || !cols || !vals))) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const int rc = upload_transition(ctx, nnz, rows, cols, vals, ctx->d_csc_ptr, &ctx->d_csc_row, &ctx->d_csc_val, &ctx->csc_cap);
  if (rc) return rc;
  ctx->transition_set = true;
  return BMC_OK;
}

// ---- gas phase: second scalar field + gas-liquid mass transfer -----------------------------
static int check_feeds(bmc_ctx* ctx, uint64_t n, const bmc_feed* f) {
  if (n > (uint64_t)kMaxFlows) { ctx->err = "too many feed entries (max 16)"; return BMC_ERR_UNSUPPORTED; }
  for (uint64_t i = 0; i < n; ++i) {
    if (f[i].species >= ctx->n_species || f[i].input_position >= ctx->n_comp || (f[i].has_output && f[i].output_position >= ctx->n_comp)) {
      ctx->err = "feed index out of range"; return BMC_ERR_RANGE;
    }
    if (f[i].flow < 0) { ctx->err = "negative feed flow"; return BMC_ERR_INVALID; }
  }
  return BMC_OK;
}

int bmc_gas_enable(bmc_ctx* ctx, const double* gas_volumes, const double* gas_concentrations) {
  if (!ctx || !gas_volumes) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const size_t nb = ctx->n_species * ctx->n_comp, nc = ctx->n_comp;
  int rc;
  if (!ctx->d_gconc) {
    if ((rc = dev_alloc(ctx, &ctx->d_gconc, nb)) || (rc = dev_alloc(ctx, &ctx->d_gconc_next, nb)) || (rc = dev_alloc(ctx, &ctx->d_gmass, nb)) ||
        (rc = dev_alloc(ctx, &ctx->d_gvol, nc)) || (rc = dev_alloc(ctx, &ctx->d_kla, nb)) || (rc = dev_alloc(ctx, &ctx->d_henry, ctx->n_species)) ||
        (rc = dev_alloc(ctx, &ctx->d_mtr, nb)) || (rc = dev_alloc(ctx, &ctx->d_gcsc_ptr, nc + 1)))
      return rc;
    CK(cudaMemsetAsync(ctx->d_gcsc_ptr, 0, (nc + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_kla, 0, nb * 8, ctx->stream)); CK(cudaMemsetAsync(ctx->d_mtr, 0, nb * 8, ctx->stream));
    // Henry = 0 except species 1 (oxygen): MassTransferModel's constructor (hydro/mass_transfer.cpp:113-116)
    std::vector<double> hen(ctx->n_species, 0.0);
    if (ctx->n_species > 1) hen[1] = 3.181e-2;
    CK(cudaMemcpy(ctx->d_henry, hen.data(), ctx->n_species * 8, cudaMemcpyHostToDevice));
  }
  for (size_t j = 0; j < nc; ++j) if (!(gas_volumes[j] > 0.0)) { ctx->err = "gas volumes must be positive"; return BMC_ERR_INVALID; }
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemcpy(ctx->d_gvol, gas_volumes, nc * 8, cudaMemcpyHostToDevice));
  if (gas_concentrations) {
    for (size_t k = 0; k < nb; ++k) if (gas_concentrations[k] < 0) { ctx->err = "gas concentrations must be >= 0"; return BMC_ERR_INVALID; }  // simulation.cpp:189-199
    CK(cudaMemcpy(ctx->d_gconc, gas_concentrations, nb * 8, cudaMemcpyHostToDevice));
  } else {
    CK(cudaMemset(ctx->d_gconc, 0, nb * 8));
  }
  ctx->two_phase = true; ctx->gas_mass_dirty = true;
  return BMC_OK;
}

int bmc_gas_update_hydro(bmc_ctx* ctx, const double* gas_volumes, uint64_t nnz, const uint64_t* rows, const uint64_t* cols, const double* vals) {
  if (!ctx || !gas_volumes || (nnz && (!rows || !cols || !vals))) return BMC_ERR_INVALID;
  if (!ctx->two_phase) { ctx->err = "bmc_gas_enable not called"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  int rc;
  unsigned char* h = nullptr; int slot = 0;
  if ((rc = map_stage(ctx, ctx->n_comp * 8, &h, &slot))) return rc;
  memcpy(h, gas_volumes, ctx->n_comp * 8);
  CK(cudaMemcpyAsync(ctx->d_gvol, h, ctx->n_comp * 8, cudaMemcpyHostToDevice, ctx->stream));  // setVolumes (simulation.cpp:133-136)
  CK(cudaEventRecord(ctx->ev_map[slot], ctx->stream));
  if ((rc = upload_transition(ctx, nnz, rows, cols, vals, ctx->d_gcsc_ptr, &ctx->d_gcsc_row, &ctx->d_gcsc_val, &ctx->gcsc_cap))) return rc;  // set_transition (:137)
  ctx->gas_transition_set = true;
  return BMC_OK;
}

int bmc_gas_set_feeds(bmc_ctx* ctx, uint64_t n, const bmc_feed* f) {
  if (!ctx || (n && !f)) return BMC_ERR_INVALID;
  const int rc = check_feeds(ctx, n, f);
  if (rc) return rc;
  ctx->gas_feeds.assign(f, f + n);
  return BMC_OK;
}

int bmc_mass_transfer_set(bmc_ctx* ctx, const double* kla, const double* henry) {
  if (!ctx || !kla) return BMC_ERR_INVALID;
  if (!ctx->two_phase) { ctx->err = "bmc_gas_enable not called"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  const size_t nb = ctx->n_species * ctx->n_comp;
  int rc;
  unsigned char* h = nullptr; int slot = 0;
  if ((rc = map_stage(ctx, (nb + ctx->n_species) * 8, &h, &slot))) return rc;
  memcpy(h, kla, nb * 8);
  CK(cudaMemcpyAsync(ctx->d_kla, h, nb * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (henry) {
    memcpy(h + nb * 8, henry, ctx->n_species * 8);
    CK(cudaMemcpyAsync(ctx->d_henry, h + nb * 8, ctx->n_species * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaEventRecord(ctx->ev_map[slot], ctx->stream));
  ctx->mtr_set = true;
  return BMC_OK;
}

int bmc_get_gas_concentrations(bmc_ctx* ctx, double* out) {
  if (!ctx || !out || !ctx->two_phase) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_gconc, ctx->n_species * ctx->n_comp * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BMC_OK;
}

int bmc_get_mass_transfer(bmc_ctx* ctx, double* out) {
  if (!ctx || !out || !ctx->two_phase) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, ctx->d_mtr, ctx->n_species * ctx->n_comp * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return BMC_OK;
}

int bmc_liquid_set_feeds(bmc_ctx* ctx, uint64_t n, const bmc_feed* f) {
  if (!ctx || (n && !f)) return BMC_ERR_INVALID;
  { const int rc = check_feeds(ctx, n, f); if (rc) return rc; }
  std::vector<bmc_leaving_flow> out;
  for (uint64_t i = 0; i < n; ++i) {
    if (f[i].has_output && f[i].first_of_feed)  // set_leaving_flow(mc_flow_counter, output_position, flow, volume)
      out.push_back(bmc_leaving_flow{f[i].output_position, f[i].flow, ctx->h_vol[f[i].output_position]});
  }
  ctx->feeds.assign(f, f + n);
  ctx->flows = out;
  return BMC_OK;
}

int bmc_liquid_step(bmc_ctx* ctx, double d_t) {
  if (!ctx || !(d_t >= 0)) return BMC_ERR_INVALID;
  if (ctx->n_comp > 1 && !ctx->transition_set) { ctx->err = "bmc_liquid_set_transition not called"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const uint32_t nb = (uint32_t)(ctx->n_species * ctx->n_comp);
  int rc;
  // a pending all-reduce of the sources is finished INSIDE the liquid kernel (every element sums what the ranks
  // published): cycle -> allreduce -> liquid step costs no extra launch.  (Contexts sharing one device use the small kernel.)
  const unsigned long long consume = (ctx->p2p_pending && ctx->p2p_ipc) ? ctx->p2p_pending : 0ull;
  if (consume) ctx->p2p_pending = 0;
  else if ((rc = finish_pending_sum(ctx))) return rc;
  ctx->p2p_published = 0;  // the sources are consumed (cleared) by this step
  ctx->src_mirror_tag = 0;
  if (ctx->mass_dirty) {
    liquid_mass_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ctx->d_conc, ctx->d_vol, ctx->d_mass, (uint32_t)ctx->n_species, nb);
    if ((rc = check_launch(ctx, "liquid_mass"))) return rc;
    ctx->mass_dirty = false;
  }
  LiquidParams lp;
  memset(&lp, 0, sizeof(lp));
  lp.c_old = ctx->d_conc; lp.c_new = ctx->d_conc_next; lp.mass = ctx->d_mass; lp.vol = ctx->d_vol; lp.sources = ctx->d_sources;
  lp.csc_ptr = ctx->d_csc_ptr; lp.csc_row = ctx->d_csc_row; lp.csc_val = ctx->d_csc_val;
  lp.n_species = (uint32_t)ctx->n_species; lp.n_comp = (uint32_t)ctx->n_comp; lp.dt = d_t; lp.n_feeds = (int)ctx->feeds.size();
  fill_peer_exchange(ctx, lp.px, 0, consume);
  for (int i = 0; i < lp.n_feeds; ++i) {
    const bmc_feed& f = ctx->feeds[i];
    lp.feeds[i] = FeedDev{(uint32_t)f.species, (uint32_t)f.input_position, (uint32_t)f.output_position, f.has_output, f.first_of_feed, f.flow, f.concentration};
  }
  if (ctx->two_phase) {  // ode_step with a gas phase (simulation.model.cpp:131-154)
    if (ctx->n_comp > 1 && !ctx->gas_transition_set) { ctx->err = "bmc_gas_update_hydro not called"; return BMC_ERR_INVALID; }
    if (ctx->gas_mass_dirty) {  // set_mass (simulation.cpp:189-195)
      liquid_mass_kernel<<<(nb + 255) / 256, 256, 0, s>>>(ctx->d_gconc, ctx->d_gvol, ctx->d_gmass, (uint32_t)ctx->n_species, nb);
      if ((rc = check_launch(ctx, "gas_mass"))) return rc;
      ctx->gas_mass_dirty = false;
    }
    GasLiquidParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.liq = lp;
    gp.g_old = ctx->d_gconc; gp.g_new = ctx->d_gconc_next; gp.g_mass = ctx->d_gmass; gp.g_vol = ctx->d_gvol;
    gp.g_csc_ptr = ctx->d_gcsc_ptr; gp.g_csc_row = ctx->d_gcsc_row; gp.g_csc_val = ctx->d_gcsc_val;
    gp.n_gas_feeds = (int)ctx->gas_feeds.size();
    for (int i = 0; i < gp.n_gas_feeds; ++i) {
      const bmc_feed& f = ctx->gas_feeds[i];
      gp.gas_feeds[i] = FeedDev{(uint32_t)f.species, (uint32_t)f.input_position, (uint32_t)f.output_position, f.has_output, f.first_of_feed, f.flow, f.concentration};
    }
    gp.kla = ctx->d_kla; gp.henry = ctx->d_henry; gp.mtr = ctx->d_mtr;
    gas_liquid_step_kernel<<<(nb + 255) / 256, 256, 0, s>>>(gp, &ctx->st->error);
    if ((rc = check_launch(ctx, "gas_liquid_step"))) return rc;
    std::swap(ctx->d_gconc, ctx->d_gconc_next);
  } else {
    liquid_step_kernel<<<(nb + 255) / 256, 256, 0, s>>>(lp, &ctx->st->error);
    if ((rc = check_launch(ctx, "liquid_step"))) return rc;
  }
  std::swap(ctx->d_conc, ctx->d_conc_next);
  return BMC_OK;
}

int bmc_get_sources(bmc_ctx* ctx, double* out) {
  if (!ctx || !out) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  const size_t bytes = ctx->n_species * ctx->n_comp * 8;
  // a pending all-reduce is finished first (one small kernel, which also refreshes the tagged host mirror)
  { int rc = finish_pending_sum(ctx); if (rc) return rc; }
  if (ctx->src_mirror_tag) {
    // The last thing that wrote the sources was a cycle, whose publish phase also stored every value, tagged, into
    // pinned host memory: wait for the tags (no CUDA call, no copy, no stream synchronisation) and read them.
    const size_t nb = ctx->n_species * ctx->n_comp;
    const volatile unsigned long long* m = ctx->h_src_mirror;
    const unsigned long long tag = ctx->src_mirror_tag;
    unsigned spins = 0;
    bool ok = true;
    for (size_t k = 0; k < nb && ok; ++k) {
      while (m[2 * k + 1] != tag) {
        if ((++spins & 0xfffu) == 0u) {  // a fault on the device must not hang the host
          const cudaError_t q = cudaStreamQuery(ctx->stream);
          if (q == cudaSuccess) { if (m[2 * k + 1] != tag) ok = false; break; }
          if (q != cudaErrorNotReady) { ctx->err = std::string("cudaStreamQuery: ") + cudaGetErrorString(q); return BMC_ERR_CUDA; }
        }
      }
    }
    if (ok) {
      std::atomic_thread_fence(std::memory_order_acquire);
      for (size_t k = 0; k < nb; ++k) { const unsigned long long v = m[2 * k]; memcpy(out + k, &v, 8); }
      return BMC_OK;
    }
  }
  CK(cudaMemcpyAsync(ctx->h_pin_out, ctx->d_sources, bytes, cudaMemcpyDeviceToHost, ctx->stream));  // pinned: one DMA, no staging
  CK(cudaStreamSynchronize(ctx->stream));
  memcpy(out, ctx->h_pin_out, bytes);
  return BMC_OK;
}

int bmc_cycle(bmc_ctx* ctx, double d_t) {
  if (!ctx || !(d_t >= 0)) return BMC_ERR_INVALID;
  if (ctx->cap == 0) { ctx->err = "no particles: call bmc_set_particles / bmc_init_particles first"; return BMC_ERR_INVALID; }
  if (ctx->n_comp > 1 && !ctx->domain_set) { ctx->err = "multi-compartment case without bmc_domain_update"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_room(ctx))) return rc;
  cudaStream_t s = ctx->stream;
  const bool enable_move = ctx->n_comp > 1;        // kernels.hpp:53-55
  const bool enable_leave = !ctx->flows.empty();   // kernels.hpp:56
  if (ctx->lazy_ages) {
    if (!ctx->epoch_set) { ctx->epoch_set = true; ctx->epoch_dt = d_t; }
    if (ctx->epoch_dt != d_t || ctx->host_step >= 0x7ffffff0ull) {
      if ((rc = make_ages_eager(ctx))) return rc;  // the stamps assume a constant increment per step
    } else if ((rc = ensure_age_tables(ctx, ctx->host_step + 2))) return rc;
  }
  const int grid_cycle = ctx->lazy_ages ? ctx->grid_cycle : ctx->grid_cycle_eager;
  if (!ctx->ctab_in_smem) {  // large compartment table: built once per step in global memory
    PreParams pp;
    pp.diag = ctx->d_diag; pp.vol = ctx->d_vol; pp.dt = d_t; pp.conc = ctx->d_conc; pp.n_species = (uint32_t)ctx->n_species;
    pp.ctab = ctx->d_ctab; pp.n_comp = (uint32_t)ctx->n_comp; pp.enable_move = enable_move ? 1 : 0;
    pp.n_flows = (int)ctx->flows.size();
    for (int i = 0; i < pp.n_flows; ++i) { memset(&pp.outlets[i], 0, sizeof(Outlet)); pp.outlets[i].index = (uint32_t)ctx->flows[i].index; pp.outlets[i].flow = ctx->flows[i].flow; }
    const int grid = (int)std::min<uint32_t>(((uint32_t)ctx->n_comp + 255) / 256, (uint32_t)ctx->n_sm * 2);
    void* pargs[] = {&pp};
    CK(cudaLaunchKernel(ctx->vt.pre_fn, dim3(grid), dim3(256), pargs, 0, s));
    if ((rc = check_launch(ctx, "pre_step"))) return rc;
  }

  CycleParams p;
  memset(&p, 0, sizeof(p));
  p.props = ctx->props; p.cap = ctx->cap; p.pos = ctx->pos; p.status = ctx->status; p.age_hyd = ctx->age_hyd; p.age_div = ctx->age_div;
  p.st = ctx->st;
  p.buf_props = ctx->buf_props; p.buf_stride = ctx->buf_cap; p.buf_pos = ctx->buf_pos; p.buf_mother = ctx->buf_mother;
  p.div_mask = ctx->div_mask; p.tile_off = ctx->tile_off; p.blk_total = ctx->blk_total;
  p.ctab = ctx->d_ctab; p.cdf = ctx->d_cdf_f; p.neigh = ctx->d_neigh; p.m = ctx->m; p.n_comp = (uint32_t)ctx->n_comp;
  p.n_flows = (int)ctx->flows.size();
  for (int i = 0; i < p.n_flows; ++i) {
    p.outlets[i].index = (uint32_t)ctx->flows[i].index;
    p.outlets[i].flow = ctx->flows[i].flow;
    p.outlets[i].dt_flow = d_t * ctx->flows[i].flow;  // (dt * flow), probability_leaving.hpp:28
    p.outlets[i].volume = ctx->flows[i].volume;
  }
  p.conc = ctx->d_conc; p.n_species = (uint32_t)ctx->n_species; p.sources = ctx->d_sources; p.acc = ctx->d_acc; p.acc_fix = ctx->d_acc_fix;
  p.diag = ctx->d_diag; p.vol = ctx->d_vol; p.ctab_in_smem = ctx->ctab_in_smem; p.ctab_offset = (uint32_t)ctx->ctab_offset;
  p.weight = ctx->weight; p.dt = d_t; p.dt_f = (float)d_t;
  p.step = (uint32_t)ctx->host_step; p.rank = ctx->rank; p.seed_lo = (uint32_t)ctx->seed; p.seed_hi = (uint32_t)(ctx->seed >> 32);
  p.enable_move = enable_move; p.enable_leave = enable_leave; p.bins_in_smem = ctx->bins_in_smem;
  p.queue_offset = (uint32_t)ctx->queue_offset;
  p.stage_offset = (uint32_t)ctx->stage_offset; p.stage_warp_bytes = (uint32_t)ctx->stage_warp_bytes;
  p.dyn_min = ctx->dyn_min; p.dyn_shift = ctx->dyn_shift;
  // Philox constants of this step's three draw blocks, folded on the host (bmc_rng.cuh)
  p.ph0 = philox_pre(p.step, 0u, p.rank, p.seed_lo, p.seed_hi);  // u1: leaves its compartment
  p.ph1 = philox_pre(p.step, 1u, p.rank, p.seed_lo, p.seed_hi);  // u3: outlet test
  p.ph2 = philox_pre(p.step, 2u, p.rank, p.seed_lo, p.seed_hi);  // u2: neighbour pick

  // second phase of the step kernel: compaction (when triggered), newborn insertion, commit
  fill_post_params(ctx, p.post);
  p.post.count_step = 1;
  // newborns age from the next step on: their stamps are the clock values AFTER this step
  p.post.newborn_stamp_div = ctx->lazy_ages ? (uint32_t)ctx->host_step + 1u : 0u;
  p.post.newborn_stamp_hyd = ctx->lazy_ages ? (uint32_t)(ctx->hyd_clock + (enable_leave ? 1u : 0u)) : 0u;
  p.post.tab_idx_div = (uint32_t)ctx->host_step; p.post.tab_idx_hyd = (uint32_t)ctx->hyd_clock;
  p.post.tab_extend = ctx->lazy_ages ? 1 : 0; p.post.enable_leave = enable_leave ? 1 : 0; p.post.dt_f = (float)d_t; p.post.dt = d_t;
  // tags are unique per context life (launch counter), never 0
  ctx->src_mirror_tag = ctx->launches + 1;
  p.post.src_mirror = ctx->d_src_mirror; p.post.src_tag = ctx->src_mirror_tag;
  if (ctx->p2p_on) {  // this step publishes its sources to the peers and finishes the previous all-reduce, if one is pending
    // (peers that are contexts on the SAME device — the test harness — cannot be waited for from inside a kernel that
    // fills the device: their all-reduce is finished by the small kernel instead)
    if (ctx->p2p_pending && !ctx->p2p_ipc && (rc = finish_pending_sum(ctx))) return rc;
    fill_peer_exchange(ctx, p.post.px, ++ctx->p2p_epoch, ctx->p2p_pending);
    ctx->p2p_pending = 0; ctx->p2p_published = ctx->p2p_epoch;
  }

  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool prof_this = ctx->profile > 0 && ctx->host_step % (uint64_t)ctx->profile == 0;
  if (prof_this) {  // event pairs come from a pool that bmc_profile_read recycles: nothing is created in steady state
    if (ctx->prof_used == ctx->prof_events.size()) {
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      ctx->prof_events.emplace_back(e0, e1);
    }
    e0 = ctx->prof_events[ctx->prof_used].first; e1 = ctx->prof_events[ctx->prof_used].second;
    CK(cudaEventRecord(e0, s));
  }
  // ONE cooperative launch per time step (grid = resident blocks: the grid barrier between the
  // particle pass and the post-cycle phase needs every block on the device)
  void* cargs[] = {&p};
  p.fuse_post = ctx->fuse_post ? 1 : 0;
  const void* fn = ctx->lazy_ages ? ctx->vt.cycle_fn : ctx->vt.cycle_eager_fn;
  const int block = ctx->lazy_ages ? ctx->vt.block : ctx->vt.block_eager;
  const size_t smem = ctx->lazy_ages ? ctx->smem_total : ctx->smem_eager;
  if (ctx->fuse_post) {
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid_cycle), dim3(block), cargs, smem, s));
    if ((rc = check_launch(ctx, "cycle_kernel"))) return rc;
  } else {
    CK(cudaLaunchKernel(fn, dim3(grid_cycle), dim3(block), cargs, smem, s));
    if ((rc = check_launch(ctx, "cycle_kernel"))) return rc;
    void* pargs[] = {&p.post};
    CK(cudaLaunchCooperativeKernel((const void*)post_only_kernel, dim3(ctx->grid_post), dim3(kBlock), pargs, 0, s));
    if ((rc = check_launch(ctx, "post_only"))) return rc;
  }
  if (prof_this) { CK(cudaEventRecord(e1, s)); ctx->prof_used++; }
  if (enable_leave) ctx->maybe_inactive = true;  // exits may happen from now on

  ctx->host_step++;
  if (enable_leave) ctx->hyd_clock++;
  return BMC_OK;
}

int bmc_sync(bmc_ctx* ctx) {
  if (!ctx) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  { int rc = finish_pending_sum(ctx); if (rc) return rc; }
  CK(cudaStreamSynchronize(ctx->stream));
  return BMC_OK;
}

int bmc_get_counters(bmc_ctx* ctx, bmc_counters* out) {
  if (!ctx || !out) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  DevState s; int rc;
  if ((rc = sync_state(ctx, &s))) return rc;
  for (int i = 0; i < BMC_N_EVENTS; ++i) out->events[i] = s.events[i];
  out->n_used = s.n_used; out->n_inactive = s.inactive; out->last_out = s.last_out; out->last_dead = s.last_dead;
  out->last_waiting_allocation = s.last_waiting; out->buffer_index = 0; out->capacity = s.logical_alloc;
  out->total_out = s.total_out; out->total_new = s.total_new; out->n_compactions = s.n_compactions; out->step = s.step;
  out->buffer_capacity = s.logical_buf;
  out->physical_capacity = ctx->cap; out->physical_buffer_capacity = ctx->buf_cap; out->n_reallocations = ctx->n_regrow;
  return BMC_OK;
}

int bmc_repartition(bmc_ctx* ctx, uint64_t* out) {
  if (!ctx || !out) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_stage(ctx, ctx->n_comp * 8))) return rc;
  cudaStream_t s = ctx->stream;
  CK(cudaMemsetAsync(ctx->d_stage, 0, ctx->n_comp * 8, s));
  if (ctx->cap) {
    repartition_kernel<<<ctx->n_sm * 4, 256, 0, s>>>(ctx->pos, ctx->status, ctx->st, (unsigned long long*)ctx->d_stage);
    if ((rc = check_launch(ctx, "repartition"))) return rc;
  }
  CK(cudaMemcpyAsync(out, ctx->d_stage, ctx->n_comp * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return BMC_OK;
}

int bmc_get_properties(bmc_ctx* ctx, const uint64_t* indices, uint64_t n_indices, double* particle_values, double* spatial_values,
                       double* ages, uint64_t* n_particles) {
  if (!ctx || ctx->cap == 0) return BMC_ERR_INVALID;
  const int nv = ctx->vt.n_var;
  const uint32_t n_exp = indices ? (uint32_t)n_indices : (uint32_t)nv;
  if (indices) for (uint64_t e = 0; e < n_indices; ++e) if (indices[e] >= (uint64_t)nv) { ctx->err = "exported property index out of range"; return BMC_ERR_RANGE; }
  int rc;
  if ((rc = bmc_compact(ctx))) return rc;  // container.force_remove_dead() (post_process.hpp:181)
  DevState hs;
  if ((rc = sync_state(ctx, &hs))) return rc;
  const uint64_t n_p = hs.n_used;
  if (n_particles) *n_particles = n_p;
  if (!particle_values && !spatial_values && !ages) return BMC_OK;  // size query
  cudaStream_t s = ctx->stream;
  const size_t rows = (size_t)n_exp + 1, chunk = 1u << 21;
  double* d_spatial = nullptr; uint32_t* d_idx = nullptr;
  if ((rc = dev_alloc(ctx, &d_spatial, rows * ctx->n_comp))) return rc;
  CK(cudaMemsetAsync(d_spatial, 0, rows * ctx->n_comp * 8, s));
  if (indices) {
    std::vector<uint32_t> idx(indices, indices + n_indices);
    if ((rc = dev_alloc(ctx, &d_idx, n_exp))) { dev_free(d_spatial); return rc; }
    CK(cudaMemcpyAsync(d_idx, idx.data(), n_exp * 4, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
  }
  if ((rc = ensure_stage(ctx, rows * std::min<size_t>(chunk, std::max<uint64_t>(n_p, 1)) * 8))) { dev_free(d_spatial); dev_free(d_idx); return rc; }
  for (uint64_t o = 0; o < n_p; o += chunk) {
    const uint64_t c = std::min<uint64_t>(chunk, n_p - o);
    ExportParams ep{ctx->props, ctx->cap, ctx->pos, ctx->status, o, c, d_idx, n_exp, (double*)ctx->d_stage, d_spatial, (uint32_t)ctx->n_comp};
    void* args[] = {&ep};
    CK(cudaLaunchKernel(ctx->vt.export_fn, dim3((unsigned)std::min<uint64_t>((c + 255) / 256, (uint64_t)ctx->n_sm * 8)), dim3(256), args, 0, s));
    if ((rc = check_launch(ctx, "export_kernel"))) { dev_free(d_spatial); dev_free(d_idx); return rc; }
    if (particle_values)  // (n_exp + 1, n_p) LayoutRight like ParticlePropertyViewType
      for (size_t r = 0; r < rows; ++r)
        CK(cudaMemcpyAsync(particle_values + r * n_p + o, (double*)ctx->d_stage + r * c, c * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  if (spatial_values) CK(cudaMemcpyAsync(spatial_values, d_spatial, rows * ctx->n_comp * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  dev_free(d_spatial); dev_free(d_idx);
  if (ages && n_p) {  // ages_value(0,i) = hydraulic, (1,i) = since division; doubles; zero for non-idle (none after the compaction)
    std::vector<float> ah(n_p), ad(n_p);
    if ((rc = bmc_get_particles(ctx, n_p, nullptr, nullptr, nullptr, ah.data(), ad.data()))) return rc;
    for (uint64_t i = 0; i < n_p; ++i) { ages[i] = (double)ah[i]; ages[n_p + i] = (double)ad[i]; }
  }
  return BMC_OK;
}

// ---------------------------------------------------------------------------------------------
// Checkpoint / resume of the Monte-Carlo unit.  Reference: SerDe::save_simulation /
// load_simulation (apps/core/src/serde.cpp:64-219) archive {version, number_particle, dimensions, liquid
// concentrations, time, MonteCarloUnit{init_weight, events, domain ids, container{n_allocated, n_used,
// rt_params, weights, position, status, model, ages}}} with cereal.  cereal is not available and its byte
// layout is not part of the reference tree, so the same content is written in a self-described
// little-endian layout (include/bmc.h).  The domain (flow map) is NOT part of it, as in the reference: the
// caller re-applies bmc_domain_update / bmc_set_leaving_flows from its case before resuming.
// Counter-based RNG: restoring `step` resumes every random stream exactly, so a resumed run is bit-identical
// to an uninterrupted one.
// ---------------------------------------------------------------------------------------------
namespace {
struct CkptHeader {
  char magic[8];            // "BMCCKPT2"
  uint32_t header_bytes, model, n_var, lazy_ages;
  uint64_t n_species, n_comp, seed, rank, step, n_used, capacity_hint;
  uint64_t inactive, total_out, total_new, n_compactions, last_out, last_dead, last_waiting;
  uint64_t events[6];
  double weight, allocation_factor, buffer_ratio, dead_ratio, epoch_dt;
  uint64_t min_removal;
  uint32_t epoch_set, reserved_epoch;
  uint64_t hyd_clock;
  uint64_t tab_entries;     // lazy ages: entries of each age table that follow
  float src_bound[8], src_scale[8];  // fixed-point scatter state (DevState): a resumed run adds up the very same integers
  uint64_t logical_alloc, logical_buf;  // n_allocated_elements / buffer extent of the reference's container (serde.cpp archives n_allocated)
};
}  // namespace

static uint64_t ckpt_bytes(const bmc_ctx* ctx, uint64_t n, uint64_t tab_entries) {
  const uint64_t nb = ctx->n_species * ctx->n_comp;
  return sizeof(CkptHeader) + 2 * nb * 8 + (uint64_t)ctx->vt.n_var * n * 4 + n * 4 + n + 2 * n * 4 + 2 * tab_entries * 4;
}

int bmc_checkpoint_size(bmc_ctx* ctx, uint64_t* bytes) {
  if (!ctx || !bytes) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  DevState hs; int rc;
  if ((rc = sync_state(ctx, &hs))) return rc;
  *bytes = ckpt_bytes(ctx, hs.n_used, ctx->lazy_ages ? ctx->host_step + 1 : 0);
  return BMC_OK;
}

int bmc_checkpoint_save(bmc_ctx* ctx, void* buffer, uint64_t bytes) {
  if (!ctx || !buffer) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  DevState hs; int rc;
  if ((rc = sync_state(ctx, &hs))) return rc;
  // entries [0, host_step] of the age tables are written (bmc_cycle guarantees host_step + 1 allocated entries)
  const uint64_t n = hs.n_used, tab = ctx->lazy_ages ? ctx->host_step + 1 : 0;
  if (bytes < ckpt_bytes(ctx, n, tab)) { ctx->err = "bmc_checkpoint_save: buffer too small (see bmc_checkpoint_size)"; return BMC_ERR_RANGE; }
  CkptHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "BMCCKPT2", 8);
  h.header_bytes = (uint32_t)sizeof(h); h.model = (uint32_t)ctx->model; h.n_var = (uint32_t)ctx->vt.n_var; h.lazy_ages = ctx->lazy_ages ? 1u : 0u;
  h.n_species = ctx->n_species; h.n_comp = ctx->n_comp; h.seed = ctx->seed; h.rank = ctx->rank; h.step = ctx->host_step; h.n_used = n;
  h.capacity_hint = ctx->cap;
  h.inactive = hs.inactive; h.total_out = hs.total_out; h.total_new = hs.total_new; h.n_compactions = hs.n_compactions;
  h.last_out = hs.last_out; h.last_dead = hs.last_dead; h.last_waiting = hs.last_waiting;
  for (int i = 0; i < 6; ++i) h.events[i] = hs.events[i];
  h.weight = (double)ctx->weight; h.allocation_factor = ctx->allocation_factor; h.buffer_ratio = ctx->buffer_ratio; h.dead_ratio = ctx->dead_ratio;
  h.epoch_dt = ctx->epoch_dt; h.min_removal = ctx->min_removal; h.epoch_set = ctx->epoch_set ? 1u : 0u; h.hyd_clock = ctx->hyd_clock;
  h.tab_entries = tab;
  for (int i = 0; i < 8; ++i) { h.src_bound[i] = hs.src_bound[i]; h.src_scale[i] = hs.src_scale[i]; }
  h.logical_alloc = hs.logical_alloc; h.logical_buf = hs.logical_buf;
  unsigned char* o = (unsigned char*)buffer;
  memcpy(o, &h, sizeof(h)); o += sizeof(h);
  cudaStream_t s = ctx->stream;
  const uint64_t nb = ctx->n_species * ctx->n_comp;
  CK(cudaMemcpyAsync(o, ctx->d_conc, nb * 8, cudaMemcpyDeviceToHost, s)); o += nb * 8;
  CK(cudaMemcpyAsync(o, ctx->d_sources, nb * 8, cudaMemcpyDeviceToHost, s)); o += nb * 8;
  for (int k = 0; k < ctx->vt.n_var; ++k) { CK(cudaMemcpyAsync(o, ctx->props + (size_t)k * ctx->cap, n * 4, cudaMemcpyDeviceToHost, s)); o += n * 4; }
  CK(cudaMemcpyAsync(o, ctx->pos, n * 4, cudaMemcpyDeviceToHost, s)); o += n * 4;
  CK(cudaMemcpyAsync(o, ctx->status, n, cudaMemcpyDeviceToHost, s)); o += n;
  CK(cudaMemcpyAsync(o, ctx->age_hyd, n * 4, cudaMemcpyDeviceToHost, s)); o += n * 4;  // raw columns: step stamps or floats
  CK(cudaMemcpyAsync(o, ctx->age_div, n * 4, cudaMemcpyDeviceToHost, s)); o += n * 4;
  if (tab) {
    CK(cudaMemcpyAsync(o, ctx->d_tab_hyd, tab * 4, cudaMemcpyDeviceToHost, s)); o += tab * 4;
    CK(cudaMemcpyAsync(o, ctx->d_tab_div, tab * 4, cudaMemcpyDeviceToHost, s)); o += tab * 4;
  }
  CK(cudaStreamSynchronize(s));
  return BMC_OK;
}

int bmc_checkpoint_load(bmc_ctx* ctx, const void* buffer, uint64_t bytes) {
  if (!ctx || !buffer || bytes < sizeof(CkptHeader)) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CkptHeader h;
  memcpy(&h, buffer, sizeof(h));
  if (memcmp(h.magic, "BMCCKPT2", 8) != 0 || h.header_bytes != sizeof(h)) { ctx->err = "bmc_checkpoint_load: not a checkpoint of this version"; return BMC_ERR_INVALID; }
  if (h.model != (uint32_t)ctx->model || h.n_var != (uint32_t)ctx->vt.n_var || h.n_species != ctx->n_species || h.n_comp != ctx->n_comp) {
    ctx->err = "bmc_checkpoint_load: model / dimensions differ from this context (serde.cpp: \"model number of property mismatch\")";
    return BMC_ERR_INVALID;
  }
  const uint64_t n = h.n_used, tab = h.tab_entries;
  if (h.lazy_ages ? (tab != h.step + 1) : (tab != 0)) { ctx->err = "bmc_checkpoint_load: inconsistent age tables"; return BMC_ERR_INVALID; }
  // (n, tab <= bytes first: a corrupt header must not wrap the size computation)
  if (n > bytes || tab > bytes || bytes < ckpt_bytes(ctx, n, tab)) { ctx->err = "bmc_checkpoint_load: truncated buffer"; return BMC_ERR_RANGE; }
  CK(cudaStreamSynchronize(ctx->stream));
  int rc;
  ctx->seed = h.seed; ctx->rank = (uint32_t)h.rank; ctx->weight = (float)h.weight;
  ctx->allocation_factor = h.allocation_factor; ctx->buffer_ratio = h.buffer_ratio; ctx->dead_ratio = h.dead_ratio; ctx->min_removal = h.min_removal;
  const size_t want = std::max<size_t>(initial_capacity(ctx, n), (size_t)h.logical_alloc);
  if (want > ctx->cap || ctx->cap == 0) { if ((rc = resize_container(ctx, want, 0))) return rc; }
  cudaStream_t s = ctx->stream;
  const unsigned char* in = (const unsigned char*)buffer + sizeof(h);
  const uint64_t nb = ctx->n_species * ctx->n_comp;
  CK(cudaMemcpyAsync(ctx->d_conc, in, nb * 8, cudaMemcpyHostToDevice, s)); in += nb * 8;
  CK(cudaMemcpyAsync(ctx->d_sources, in, nb * 8, cudaMemcpyHostToDevice, s)); in += nb * 8;
  ctx->mass_dirty = true; ctx->src_mirror_tag = 0; ctx->p2p_pending = 0; ctx->p2p_published = 0;
  for (int k = 0; k < ctx->vt.n_var; ++k) { CK(cudaMemcpyAsync(ctx->props + (size_t)k * ctx->cap, in, n * 4, cudaMemcpyHostToDevice, s)); in += n * 4; }
  CK(cudaMemcpyAsync(ctx->pos, in, n * 4, cudaMemcpyHostToDevice, s)); in += n * 4;
  CK(cudaMemcpyAsync(ctx->status, in, n, cudaMemcpyHostToDevice, s)); in += n;
  if (ctx->cap > n) CK(cudaMemsetAsync(ctx->status + n, 0, ctx->cap - n, s));  // slots beyond n_used are Idle
  CK(cudaMemcpyAsync(ctx->age_hyd, in, n * 4, cudaMemcpyHostToDevice, s)); in += n * 4;
  CK(cudaMemcpyAsync(ctx->age_div, in, n * 4, cudaMemcpyHostToDevice, s)); in += n * 4;
  ctx->lazy_ages = h.lazy_ages != 0; ctx->epoch_set = h.epoch_set != 0; ctx->epoch_dt = h.epoch_dt;
  if (ctx->lazy_ages) {
    if ((rc = ensure_age_tables(ctx, tab + 1))) return rc;  // zero-extended: the next cycle writes entry `tab`
    CK(cudaMemcpyAsync(ctx->d_tab_hyd, in, tab * 4, cudaMemcpyHostToDevice, s)); in += tab * 4;
    CK(cudaMemcpyAsync(ctx->d_tab_div, in, tab * 4, cudaMemcpyHostToDevice, s)); in += tab * 4;
    ctx->hyd_clock = h.hyd_clock;
    if (force_eager_ages()) { ctx->host_step = h.step; if ((rc = make_ages_eager(ctx))) return rc; }
  }
  DevState ds;
  memset(&ds, 0, sizeof(ds));
  ds.n_used = n; ds.inactive = h.inactive; ds.total_out = h.total_out; ds.total_new = h.total_new; ds.n_compactions = h.n_compactions;
  ds.last_out = h.last_out; ds.last_dead = h.last_dead; ds.last_waiting = h.last_waiting; ds.step = h.step;
  for (int i = 0; i < 6; ++i) ds.events[i] = h.events[i];
  for (int i = 0; i < 8; ++i) { ds.src_bound[i] = h.src_bound[i]; ds.src_scale[i] = h.src_scale[i]; }
  ds.logical_alloc = h.logical_alloc; ds.logical_buf = h.logical_buf;
  CK(cudaMemcpyAsync(ctx->st, &ds, sizeof(ds), cudaMemcpyHostToDevice, s));
  prepare_kernel<<<1, 32, 0, s>>>(ctx->st, (unsigned long long)ctx->cap, (unsigned long long)ctx->buf_cap, 0, ctx->allocation_factor,
                                  ctx->buffer_ratio, ctx->d_pin);
  if ((rc = check_launch(ctx, "prepare"))) return rc;
  DevState hs;
  if ((rc = sync_state(ctx, &hs))) return rc;  // also waits for the copies out of the caller's buffer
  ctx->maybe_inactive = h.inactive != 0 || !ctx->flows.empty();
  ctx->host_step = h.step; ctx->recent_max_add = 0; ctx->pin_seen_step = ~0ull; ctx->pin_history = 0;
  return BMC_OK;
}

int bmc_compact(bmc_ctx* ctx) {
  if (!ctx || ctx->cap == 0) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  int rc;
  const unsigned int one = 1;
  CK(cudaMemcpyAsync(&ctx->st->force_compact, &one, 4, cudaMemcpyHostToDevice, s));
  PostParams ip;
  fill_post_params(ctx, ip);
  ip.count_step = 0;  // no newborn is waiting (the buffer index is reset by every commit): compaction + commit only
  ip.newborn_stamp_div = 0; ip.newborn_stamp_hyd = 0; ip.tab_idx_div = 0; ip.tab_idx_hyd = 0; ip.tab_extend = 0; ip.enable_leave = 0; ip.dt_f = 0.f; ip.dt = 0.0;
  void* pargs[] = {&ip};
  CK(cudaLaunchCooperativeKernel((const void*)post_only_kernel, dim3(ctx->grid_post), dim3(kBlock), pargs, 0, s));
  if ((rc = check_launch(ctx, "post_only"))) return rc;
  DevState hs;
  return sync_state(ctx, &hs);
}

int bmc_sources_device(bmc_ctx* ctx, double** ptr, uint64_t* n) {
  if (!ctx || !ptr) return BMC_ERR_INVALID;
  { int rc = finish_pending_sum(ctx); if (rc) return rc; }
  *ptr = ctx->d_sources; if (n) *n = ctx->n_species * ctx->n_comp;
  return BMC_OK;
}
int bmc_concentrations_device(bmc_ctx* ctx, double** ptr, uint64_t* n) {
  if (!ctx || !ptr) return BMC_ERR_INVALID;
  *ptr = ctx->d_conc; if (n) *n = ctx->n_species * ctx->n_comp;
  return BMC_OK;
}
int bmc_stream(bmc_ctx* ctx, void** s) {
  if (!ctx || !s) return BMC_ERR_INVALID;
  *s = (void*)ctx->stream;
  return BMC_OK;
}
int bmc_launch_count(const bmc_ctx* ctx, uint64_t* n) {
  if (!ctx || !n) return BMC_ERR_INVALID;
  *n = ctx->launches;
  return BMC_OK;
}
int bmc_kernel_config(const bmc_ctx* ctx, int32_t* out6) {
  if (!ctx || !out6) return BMC_ERR_INVALID;
  out6[0] = ctx->vt.vec; out6[1] = ctx->vt.block; out6[2] = ctx->vt.block_eager;
  out6[3] = ctx->lazy_ages ? ctx->grid_cycle : ctx->grid_cycle_eager; out6[4] = ctx->lazy_ages ? 1 : 0; out6[5] = (int32_t)ctx->smem_total;
  return BMC_OK;
}
int bmc_profile_enable(bmc_ctx* ctx, int on) {
  if (!ctx) return BMC_ERR_INVALID;
  ctx->profile = on < 0 ? 0 : on;  // 1 = every step kernel, N = every N-th (two event records cost a few microseconds each)
  return BMC_OK;
}
// tuning aid (not declared in bmc.h): %globaltimer stamps of block 0 of the last step kernel; only
// BMC_TIMELINE builds write them
int bmc_debug_timeline(bmc_ctx* ctx, unsigned long long* out16) {
  if (!ctx || !out16) return BMC_ERR_INVALID;
  DevState s; int rc;
  if ((rc = sync_state(ctx, &s))) return rc;
  memcpy(out16, s.dbg, sizeof(s.dbg));
  return BMC_OK;
}
int bmc_debug_blocks(bmc_ctx* ctx, uint32_t* out, uint64_t n_words) {  // per-block stamps of BMC_TIMELINE builds
  if (!ctx || !out) return BMC_ERR_INVALID;
  CK(cudaMemcpy(out, ctx->src, n_words * 4, cudaMemcpyDeviceToHost));
  return BMC_OK;
}
int bmc_profile_read(bmc_ctx* ctx, double* ms_total, uint64_t* n) {
  if (!ctx) return BMC_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < ctx->prof_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_events[i].first, ctx->prof_events[i].second) == cudaSuccess) { ctx->prof_ms += ms; ctx->prof_n++; }
  }
  ctx->prof_used = 0;
  if (ms_total) *ms_total = ctx->prof_ms;
  if (n) *n = ctx->prof_n;
  ctx->prof_ms = 0.0; ctx->prof_n = 0;
  return BMC_OK;
}

// ---- NCCL (resolved lazily with dlopen so the library has no link-time
// dependency and shares whatever libnccl the process already loaded) ----------
namespace {
struct NcclId { char b[128]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
}  // namespace
static NcclApi g_nccl;
static bool load_nccl(std::string& err) {
  if (g_nccl.lib) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nme : names) { g_nccl.lib = dlopen(nme, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
  if (!g_nccl.lib) { err = "libnccl.so.2 not found"; return false; }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl_destroy = g_nccl.CommDestroy;
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) { err = "libnccl symbols missing"; return false; }
  return true;
}
int bmc_nccl_unique_id(uint8_t* id128) {
  std::string err;
  if (!id128 || !load_nccl(err)) return BMC_ERR_NCCL;
  return g_nccl.GetUniqueId(id128) == 0 ? BMC_OK : BMC_ERR_NCCL;
}
int bmc_comm_init(bmc_ctx* ctx, int n_ranks, int rank, const uint8_t* id128) {
  if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return BMC_ERR_INVALID;
  if (!load_nccl(ctx->err)) return BMC_ERR_NCCL;
  CK(cudaSetDevice(ctx->device));
  NcclId id; memcpy(id.b, id128, 128);
  const int r = g_nccl.CommInitRank(&ctx->nccl_comm, n_ranks, id, rank);
  if (r != 0) { ctx->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return BMC_ERR_NCCL; }
  ctx->nccl_ranks = n_ranks;
  return BMC_OK;
}
// ---- peer-memory all-reduce ------------------------------------------------------------------
static int p2p_alloc_region(bmc_ctx* ctx) {
  if (ctx->p2p_region) return BMC_OK;
  CK(cudaSetDevice(ctx->device));
  ctx->p2p_bytes = 128 + 2 * ctx->n_species * ctx->n_comp * sizeof(double);
  CK(cudaMalloc((void**)&ctx->p2p_region, ctx->p2p_bytes));  // plain cudaMalloc: exportable with cudaIpcGetMemHandle
  CK(cudaMemset(ctx->p2p_region, 0, ctx->p2p_bytes));
  CK(cudaDeviceSynchronize());
  return BMC_OK;
}
int bmc_p2p_export(bmc_ctx* ctx, uint8_t* handle64) {
  if (!ctx || !handle64) return BMC_ERR_INVALID;
  int rc = p2p_alloc_region(ctx);
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, ctx->p2p_region));
  memcpy(handle64, &h, 64);
  return BMC_OK;
}
int bmc_p2p_region(bmc_ctx* ctx, void** base) {
  if (!ctx || !base) return BMC_ERR_INVALID;
  int rc = p2p_alloc_region(ctx);
  if (rc) return rc;
  *base = ctx->p2p_region;
  return BMC_OK;
}
static int p2p_finish_attach(bmc_ctx* ctx, int n_ranks, int rank, bool ipc) {
  ctx->p2p_world = n_ranks; ctx->p2p_rank = rank; ctx->p2p_ipc = ipc; ctx->p2p_epoch = 0; ctx->p2p_on = true;
  return BMC_OK;
}
int bmc_p2p_attach(bmc_ctx* ctx, int n_ranks, int rank, const uint8_t* handles64) {
  if (!ctx || !handles64 || n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks) return BMC_ERR_INVALID;
  if (!ctx->p2p_region) { ctx->err = "bmc_p2p_export not called"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  for (int r = 0; r < n_ranks; ++r) {
    if (r == rank) { ctx->p2p_base[r] = ctx->p2p_region; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles64 + 64 * (size_t)r, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      ctx->err = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
      for (int q = 0; q < r; ++q) if (q != rank && ctx->p2p_base[q]) { cudaIpcCloseMemHandle(ctx->p2p_base[q]); ctx->p2p_base[q] = nullptr; }
      return BMC_ERR_CUDA;
    }
    ctx->p2p_base[r] = (unsigned char*)ptr;
  }
  return p2p_finish_attach(ctx, n_ranks, rank, true);
}
int bmc_p2p_attach_local(bmc_ctx* ctx, int n_ranks, int rank, void* const* bases) {
  if (!ctx || !bases || n_ranks < 1 || n_ranks > kMaxPeers || rank < 0 || rank >= n_ranks) return BMC_ERR_INVALID;
  if (!ctx->p2p_region || bases[rank] != ctx->p2p_region) { ctx->err = "bmc_p2p_attach_local: bases[rank] must be this context's region"; return BMC_ERR_INVALID; }
  for (int r = 0; r < n_ranks; ++r) ctx->p2p_base[r] = (unsigned char*)bases[r];
  return p2p_finish_attach(ctx, n_ranks, rank, false);
}

int bmc_p2p_disable(bmc_ctx* ctx) {
  if (!ctx) return BMC_ERR_INVALID;
  { int rc = finish_pending_sum(ctx); if (rc) return rc; }
  ctx->p2p_on = false;  // bmc_allreduce_sources goes back to the NCCL communicator
  return BMC_OK;
}

int bmc_allreduce_sources(bmc_ctx* ctx) {
  if (!ctx) return BMC_ERR_INVALID;
  if (ctx->p2p_on) {
    // Peer-memory path (bmc_kernels.cuh: PeerExchange).  The last cycle's commit block has already written this rank's
    // sources into its exchange buffer; the sum over the ranks is finished by whoever touches the sources next — the
    // next step kernel does it in its first block while the particle pass runs, so the collective costs no launch and
    // no time on the critical path.  Nothing is enqueued here.
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = finish_pending_sum(ctx))) return rc;  // (two requests in a row)
    if (!ctx->p2p_published) {  // the sources were not published by a cycle (first call after a load, liquid step in between, ...)
      PeerExchange x;
      fill_peer_exchange(ctx, x, ++ctx->p2p_epoch, 0);
      p2p_exchange_kernel<<<1, 1024, 0, ctx->stream>>>(x, ctx->d_sources, (uint32_t)(ctx->n_species * ctx->n_comp), ctx->st, nullptr, 0ull);
      if ((rc = check_launch(ctx, "p2p_exchange"))) return rc;
      ctx->p2p_published = ctx->p2p_epoch;
    }
    ctx->p2p_pending = ctx->p2p_published;
    return BMC_OK;
  }
  if (!ctx->nccl_comm) { ctx->err = "bmc_comm_init / bmc_p2p_attach not called"; return BMC_ERR_INVALID; }
  CK(cudaSetDevice(ctx->device));
  ctx->src_mirror_tag = 0;
  // ncclFloat64 = 8, ncclSum = 0
  const int r = g_nccl.AllReduce(ctx->d_sources, ctx->d_sources, ctx->n_species * ctx->n_comp, 8, 0, ctx->nccl_comm, ctx->stream);
  if (r != 0) { ctx->err = std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return BMC_ERR_NCCL; }
  ctx->launches++;
  return BMC_OK;
}

}  // extern "C"
