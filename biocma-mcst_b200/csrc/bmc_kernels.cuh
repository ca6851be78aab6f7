// Hand-written sm_100a kernels of the Monte-Carlo particle step.
//
// One time step of SimulationUnit::cycleProcess
// (apps/libs/simulation/public/simulation/simulation.hpp:183-239) is, on the
// reference, four full passes over the particle arrays (cycle_model,
// cycle_model_contribs, cycle_move, cycle_move_leave: kernels.hpp:123-224) plus a
// host synchronisation.  Here it is ONE cooperative launch (`cycle_kernel`) with no host
// synchronisation:
//
//   particle pass   one streaming pass over structure-of-arrays state: fused model update +
//                   division + contribution scatter + move + outlet exit   [HBM-bound, dominant]
//   grid barrier
//   post-cycle      publish the source terms, update_and_remove_inactive (deterministic stream
//                   compaction, only when triggered), merge_buffer (newborns appended in
//                   ascending-mother order), commit of the device-resident counters
//
// Only compartment tables too large for shared memory add a second launch (`pre_step`).
#pragma once
#include "bmc_models.cuh"
#include "bmc_rng.cuh"

namespace bmc {

// Cache-streaming ld/st hints (ld.global.cs / st.global.cs) on the particle columns were measured to
// HURT: with 32 resident warps per SM and more than ~3e7 particles the step becomes 60-90 % slower
// (1e8 particles, monod: 1620 us with the hints, 855 us without), so plain accesses are the default.
#ifndef BMC_STREAM_HINTS
#define BMC_STREAM_HINTS 0   // tuning builds: 1 = loads, 2 = stores, 3 = both with the cache-streaming policy
#endif
#if BMC_STREAM_HINTS & 1
#define BMC_LD(p) __ldcs(p)
#else
#define BMC_LD(p) (*(p))
#endif
#if BMC_STREAM_HINTS & 2
#define BMC_ST(p, v) __stcs(p, v)
#else
#define BMC_ST(p, v) (*(p) = (v))
#endif

struct FullTile { static constexpr bool value = true; };
struct RaggedTile { static constexpr bool value = false; };

constexpr int kTile = 1024;         // particles per rank-tile (bitmask / prefix granularity)
constexpr int kBlock = 256;         // thread-count unit: the step kernel runs ONE block of 256*WB threads per SM; post_only uses 256
constexpr int kMaxFlows = 16;       // outlets (reference: n_flows <= ~10)
constexpr int kMaxGrid = 2048;      // upper bound of the persistent grid

// Device-resident bookkeeping (replaces the host-side counters of
// ParticlesContainer and the SharedSpace EventContainer).
struct DevState {
  unsigned long long n_used;        // ParticlesContainer::n_used_elements
  unsigned long long inactive;      // inactive_counter
  unsigned long long buf_index;     // buffer_index (atomic slot allocator)
  unsigned long long buf_cap_eff;   // buffer capacity usable this step
  unsigned long long events[6];     // EventContainer::_events
  unsigned long long step_exit;     // move_reducer of this step
  unsigned long long step_waiting;  // cycle_reducer.waiting_allocation_particle
  unsigned long long last_out, last_dead, last_waiting;
  unsigned long long total_out, total_new, n_compactions;
  unsigned long long step;
  // plan of the current post-cycle
  unsigned long long n_add;         // newborns to merge
  unsigned long long cyc_n_used;    // n_used seen by the cycle kernel
  unsigned int cyc_tiles;           // tiles seen by the cycle kernel
  unsigned int cyc_grid;            // grid of the cycle kernel
  unsigned int do_compact;          // 1 = compaction triggered
  unsigned int force_compact;       // host request (force_remove_dead)
  unsigned long long cmp_old_n;     // n_used before compaction
  unsigned long long cmp_new_n;     // n_used after compaction
  unsigned int cmp_tiles;
  unsigned int cmp_total_idle;      // idle particles found in the compaction tail
  unsigned int error;               // sticky device-side error flags (1 = bad position, 2 = compaction mismatch)
  double init_mass;                 // total mass reduce of mc_init_first
  unsigned int done_blocks;         // ticket counter: the last cycle block to finish writes the plan
  unsigned int pad0;
  unsigned int bar_count, bar_gen;  // grid barrier of the cooperatively launched step kernel
  unsigned int next_group, pad2;    // work counter of the particle pass (groups drawn dynamically)
  // fixed-point source accumulation of the step kernel (see "Scatter" in cycle_body): per species the
  // largest |contribution| admitted to the integer bins this step, the power-of-two scale applied to it, and the
  // running maximum observed this step (float bits; the commit thread derives the next step's bound/scale from it)
  float src_bound[8]; float src_scale[8]; unsigned int src_max[8];
  // the reference's container bookkeeping, followed exactly (particles_container.hpp:575-643, 669-685, 784-797):
  // n_allocated_elements and the extent of the division buffer.  The buffer room of a step is the LOGICAL extent;
  // the physical arrays (PostParams::cap / buf_cap) are kept ahead of it by the host (bmc_api.cu: ensure_room).
  unsigned long long logical_alloc, logical_buf;
  unsigned long long dbg[16];       // BMC_TIMELINE builds: %globaltimer stamps of block 0 (tuning only)
};

// Host-visible mirror of the counters the capacity policy needs, written by the commit thread of every step into
// pinned, device-mapped host memory: the host reads it without any CUDA call (bmc_api.cu: ensure_room).
// Two 16-byte records, each written with ONE vector store (a single PCIe write: never torn), both tagged with the
// step they belong to: the host takes a snapshot when the tags agree.  No fence, no dependent store — the commit
// thread does not wait for host memory.
struct PinState {
  uint4 a;   // { step (low 32 bits), n_used, n_add (saturated), error }
  uint4 b;   // { step (low 32 bits), logical_alloc >> 32 | (logical_buf >> 32) << 16, logical_alloc low, logical_buf low }
};
__device__ __forceinline__ void pin_write(PinState* pin, unsigned long long step, unsigned long long n_used, unsigned long long n_add,
                                          unsigned long long la, unsigned long long lb, unsigned error) {
  const unsigned add32 = n_add > 0xffffffffull ? 0xffffffffu : (unsigned)n_add;
  const unsigned b1 = (unsigned)((la >> 32) & 0xffffu) | ((unsigned)((lb >> 32) & 0xffffu) << 16);
  asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(&pin->b), "r"((unsigned)step), "r"(b1), "r"((unsigned)la), "r"((unsigned)lb) : "memory");
  asm volatile("st.volatile.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(&pin->a), "r"((unsigned)step), "r"((unsigned)n_used), "r"(add32), "r"(error) : "memory");
}
constexpr unsigned kErrCapacity = 8u;  // DevState::error: a division was refused for lack of PHYSICAL room (host policy failed)

// merge_buffer's _resize + __allocate_buffer__ on the logical extents (particles_container.hpp:575-643, 669-685)
__host__ __device__ inline void logical_grow(unsigned long long new_size, double allocation_factor, double buffer_ratio,
                                             unsigned long long& la, unsigned long long& lb) {
  if (new_size > 0ull && new_size > la) la = (unsigned long long)ceil((double)new_size * allocation_factor);
  const unsigned long long req = (unsigned long long)ceil(buffer_ratio * (double)la);
  if (lb < req) lb = req;
}

struct Outlet { uint32_t index; uint32_t pad; double flow; double dt_flow; double volume; };

// -----------------------------------------------------------------------------
// Step-stamped ages.  The reference adds d_t to both ages of every idle particle on every
// step (ages(i,1) += float(d_t) model_kernel.hpp:191; ages(i,0) += d_t move_kernel.hpp:596, the
// latter only in steps that run the leave kernel, i.e. while the domain has an outlet) —
// 16 bytes of HBM traffic per particle-step for values no kernel ever reads.  While d_t is
// constant and every age started at zero, an age is a pure function of the number of increments
// since it was last reset, BITWISE: the k-fold floating-point accumulation A[k] = fl(A[k-1] + d_t)
// is the same for every particle.  Each age has its own clock — the division age ticks every
// step, the hydraulic age only in steps with an outlet (adding nothing is exact, so switching
// outlets on and off, as a fed-batch run does, costs nothing) — and the age columns hold 32-bit
// clock stamps
//     idle particle :  s            age = A[clock - s]   (s = clock value at the reset)
//     frozen        :  kFrozen | k  age = A[k]           (exited particle: no longer updated)
// which are written only when an age is reset (division, birth, exit).  A_div / A_hyd are
// extended by one entry per tick on the device (commit thread) and applied when ages are read
// (bmc_get_particles) — bit-identical to the eager accumulation.  If d_t changes, or the caller
// supplies non-zero ages, the columns are converted to floats in place and the eager kernel
// variant (LAZY = false) takes over.
// -----------------------------------------------------------------------------
constexpr uint32_t kFrozen = 0x80000000u;

#if defined(BMC_TIMELINE)
#define BMC_STAMP(st, i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (st)->dbg[i] = t_; } } while (0)
#else
#define BMC_STAMP(st, i) do { } while (0)
#endif

// -----------------------------------------------------------------------------
// Peer-memory exchange of the source vector (the one collective of the path: MPI_Reduce of the sources in
// apps/core/src/sync.cpp:57-78), fused into the step kernel.  Every rank owns one small region that all ranks map
// (NVLink / NVSwitch peer mappings): [flag, padded to 128 B][buffer 0][buffer 1].
//   publish (commit block of step e): sources -> my buffer[e & 1], system-scope fence, flag = e
//   consume (block 0 of the NEXT step kernel, concurrently with its particle pass; or consume_kernel when something
//            else reads the sources first): wait for every peer's flag >= e, sources = sum over ranks IN RANK ORDER of
//            buffer[e & 1] (bitwise the same vector on every rank)
// Two buffers suffice: a rank publishes e + 2 at the end of its step e + 2, whose first block has consumed e + 1 from
// every peer, i.e. every peer has finished step e + 1, whose first block consumed e.
// A peer that never shows up trips a clock limit (error bit 4) instead of hanging the device.
// -----------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerExchange {
  int world, rank;                  // world == 0: no peer exchange
  unsigned long long publish_epoch; // epoch this step publishes (0 = none)
  unsigned long long consume_epoch; // epoch to sum before this step's particle pass (0 = none)
  long long spin_limit;             // clock64 ticks before a missing peer is reported
  unsigned char* base[kMaxPeers];   // exchange region of every rank (own entry = local pointer)
};
__device__ __forceinline__ volatile unsigned long long* p2p_flag(unsigned char* base) { return reinterpret_cast<volatile unsigned long long*>(base); }
__device__ __forceinline__ double* p2p_buf(unsigned char* base, uint32_t n, unsigned par) { return reinterpret_cast<double*>(base + 128) + (size_t)par * n; }
// executed by one whole block, after every block's publish of `sources` is visible
__device__ __forceinline__ void p2p_publish(const PeerExchange& x, const double* sources, uint32_t n) {
  double* mine = p2p_buf(x.base[x.rank], n, (unsigned)(x.publish_epoch & 1ull));
  for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) mine[k] = __ldcg(sources + k);
  __syncthreads();
  // release: the block barrier orders every thread's stores before thread 0's system-scope fence (cumulative)
  if (threadIdx.x == 0) { __threadfence_system(); *p2p_flag(x.base[x.rank]) = x.publish_epoch; }
}
// The two-buffer argument holds when every rank finishes the all-reduce of every epoch it publishes (the SPMD loop
// cycle -> bmc_allreduce_sources).  A rank that skips some lets a peer run two epochs ahead and overwrite the buffer
// being read: detected here (the peer's flag has reached consume_epoch + 2) and reported like a missing peer.
__device__ __forceinline__ void p2p_check_not_overrun(const PeerExchange& x, unsigned int* error) {
  __syncthreads();
  if ((int)threadIdx.x < x.world && (int)threadIdx.x != x.rank && *p2p_flag(x.base[threadIdx.x]) >= x.consume_epoch + 2ull) atomicOr(error, 4u);
}
// executed by one whole block.  `mirror` (optional): pinned host records {value bits, tag} like the publish phase of the
// step kernel writes them (PostParams::src_mirror), so that a host reader needs neither a copy nor a stream synchronisation
__device__ __forceinline__ void p2p_consume(const PeerExchange& x, double* sources, uint32_t n, unsigned int* error,
                                            unsigned long long* mirror = nullptr, unsigned long long tag = 0ull) {
  if ((int)threadIdx.x < x.world && (int)threadIdx.x != x.rank) {
    volatile unsigned long long* f = p2p_flag(x.base[threadIdx.x]);
    const long long t0 = clock64();
    while (*f < x.consume_epoch) {
      if (clock64() - t0 > x.spin_limit) { atomicOr(error, 4u); break; }
    }
    __threadfence_system();  // acquire, by the threads that observed the flags; the barrier passes it on
  }
  __syncthreads();
  const unsigned par = (unsigned)(x.consume_epoch & 1ull);
  for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < x.world; ++r) a += *reinterpret_cast<volatile double*>(p2p_buf(x.base[r], n, par) + k);
    sources[k] = a;
    if (mirror)
      asm volatile("st.volatile.v2.u64 [%0], {%1, %2};" ::"l"(mirror + 2 * k), "l"((unsigned long long)__double_as_longlong(a)), "l"(tag) : "memory");
  }
  p2p_check_not_overrun(x, error);
}

struct PostParams {
  float* props; size_t cap; int n_var;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  // compaction scratch
  uint32_t* tile_gap_off; uint32_t* tile_idle_off; uint32_t* blk_gap; uint32_t* blk_idle; uint32_t* src;
  // division buffer + ranking data written by the cycle kernel
  const float* buf_props; size_t buf_stride; const uint32_t* buf_pos; const uint32_t* buf_mother;
  const uint32_t* div_mask; uint32_t* tile_off; uint32_t* blk_total;
  unsigned long long buf_cap;
  double* acc; double* sources; uint32_t n_bins;
  unsigned long long* acc_fix; double weight; uint32_t n_species; int n_c;  // fixed-point bins of the step kernel (publish: sources = w * fix / scale + acc)
  int vec;  // slots per thread of the step kernel: layout of div_mask (VEC ballot words per group of 32*VEC slots)
  unsigned long long min_removal; double dead_ratio;  // RuntimeParameters of update_and_remove_inactive
  double allocation_factor, buffer_ratio, shrink_ratio;  // RuntimeParameters of _resize / __allocate_buffer__ / remove_inactive_particles
  PinState* pin;                                         // device pointer of the pinned host mirror
  PeerExchange px;                                       // multi-GPU: publish this step's sources to the peers (commit block)
  // host mirror of the published sources (pinned, device-mapped): n_bins records {value bits, tag}, each written with ONE
  // 16-byte store, so the host needs neither a copy nor a fence to read them — it waits until every tag is this step's
  unsigned long long* src_mirror; unsigned long long src_tag;
  int count_step;  // 1 when called from a cycle, 0 from force_remove_dead
  // step-stamped ages (bmc_kernels.cuh): stamp given to newborns (0 = eager float ages, bits of 0.f)
  // and the per-step extension of the age tables A_div / A_hyd
  uint32_t newborn_stamp_div, newborn_stamp_hyd;
  float* tab_div; float* tab_hyd; uint32_t tab_idx_div, tab_idx_hyd;  // clock values before this step
  int tab_extend; int enable_leave; float dt_f; double dt;
};

struct CycleParams {
  // particle SoA columns (ParticlesContainer views, particles_container.hpp:82-88)
  float* props; size_t cap;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  // division buffer (particles_container.hpp:222-227)
  float* buf_props; size_t buf_stride; uint32_t* buf_pos; uint32_t* buf_mother;
  uint32_t* div_mask;   // 1 bit / slot: mother divided this step (allocated a buffer row); rewritten every step
  uint32_t* tile_off;   // per tile: exclusive prefix of the division counts inside the owning block
  uint32_t* blk_total;  // per block: divisions in its range
  // domain (DomainState, domain.hpp:28-35) in derived single-precision form
  const float* ctab;       // compartment table rows {leave threshold, model terms...}
  const float* cdf;        // floor_f32(cumulative_probability), row-major n_comp x m
  const uint32_t* neigh;   // neighbors, row-major n_comp x m
  int m; uint32_t n_comp;
  int n_flows; Outlet outlets[kMaxFlows];
  // liquid coupling
  const double* conc; uint32_t n_species;
  double* acc;       // accumulator of the source terms (zero between steps)
  unsigned long long* acc_fix;  // 64-bit fixed-point accumulator (two's complement, zero between steps)
  double* sources;   // published by the last block: sources = acc, acc = 0
  // compartment table built by every block in shared memory (small n_comp) instead of pre_step
  const double* diag; const double* vol; int ctab_in_smem; uint32_t ctab_offset;
  float weight;
  double dt; float dt_f;
  uint32_t step, rank, seed_lo, seed_hi;
  PhiloxPre ph0, ph1, ph2;  // host-folded Philox constants of draw blocks 0 (u1), 1 (u3), 2 (u2) for this step (bmc_rng.cuh)
  int enable_move, enable_leave, bins_in_smem;
  uint32_t queue_offset;  // byte offset of the per-warp deferred queues inside dynamic shared memory
  // prefetch staging (VEC == 4 only): byte offset of the per-warp staging buffers and their size (0 = no prefetch)
  uint32_t stage_offset, stage_warp_bytes;
  // work distribution of the particle pass: groups per warp drawn dynamically at the end of the pass = max(dyn_min,
  // groups per warp >> dyn_shift); the rest is grid-stride (cycle_body).  dyn_shift = 0: everything dynamic.
  uint32_t dyn_min, dyn_shift;
  PostParams post;   // second phase of the step (post_cycle_body)
  int fuse_post;     // 1 = run it in this launch behind a grid barrier (cooperative launch), 0 = post_only_kernel follows
};

__device__ __forceinline__ unsigned warp_excl_scan(unsigned v, unsigned& total) {
  const unsigned lane = threadIdx.x & 31;
  unsigned incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  return incl - v;
}

__device__ __forceinline__ uint32_t pick4(const uint32_t (&w)[4], unsigned k) {
  return k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
}

// Particle columns are streamed exactly once per step: load/store them with the
// cache-streaming policy (ld/st.global.cs) so the gathered tables (concentrations,
// leave thresholds, CDF rows, neighbours) stay resident in L1/L2.
template <int VEC> struct VecIO;
template <> struct VecIO<4> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[4]) {
    const float4 t = BMC_LD(reinterpret_cast<const float4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[4]) {
    BMC_ST(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
  static __device__ __forceinline__ void ldu(const uint32_t* p, uint32_t (&v)[4]) {
    const uint4 t = BMC_LD(reinterpret_cast<const uint4*>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ uint32_t ldb(const uint8_t* p) { return BMC_LD(reinterpret_cast<const unsigned int*>(p)); }
  static __device__ __forceinline__ uint32_t ldb_plain(const uint8_t* p) { return *reinterpret_cast<const unsigned int*>(p); }
  static __device__ __forceinline__ void ldf_plain(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void ldu_plain(const uint32_t* p, uint32_t (&v)[4]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct VecIO<2> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[2]) {
    const float2 t = BMC_LD(reinterpret_cast<const float2*>(p)); v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[2]) { BMC_ST(reinterpret_cast<float2*>(p), make_float2(v[0], v[1])); }
  static __device__ __forceinline__ void ldu(const uint32_t* p, uint32_t (&v)[2]) {
    const uint2 t = BMC_LD(reinterpret_cast<const uint2*>(p)); v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ uint32_t ldb(const uint8_t* p) { return BMC_LD(reinterpret_cast<const unsigned short*>(p)); }
  static __device__ __forceinline__ uint32_t ldb_plain(const uint8_t* p) { return *reinterpret_cast<const unsigned short*>(p); }
  static __device__ __forceinline__ void ldf_plain(const float* p, float (&v)[2]) {
    const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void ldu_plain(const uint32_t* p, uint32_t (&v)[2]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p); v[0] = t.x; v[1] = t.y;
  }
};
template <> struct VecIO<1> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[1]) { v[0] = BMC_LD(p); }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[1]) { BMC_ST(p, v[0]); }
  static __device__ __forceinline__ void ldu(const uint32_t* p, uint32_t (&v)[1]) { v[0] = BMC_LD(p); }
  static __device__ __forceinline__ uint32_t ldb(const uint8_t* p) { return BMC_LD(p); }
  static __device__ __forceinline__ uint32_t ldb_plain(const uint8_t* p) { return *p; }
  static __device__ __forceinline__ void ldf_plain(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void ldu_plain(const uint32_t* p, uint32_t (&v)[1]) { v[0] = *p; }
};


// -----------------------------------------------------------------------------
// Compartment table, one row per compartment (n_comp rows — 500 .. 10k — not N), 32-bit words:
//     col 0        leave threshold as an INTEGER (below)
//     col 1..n_pre M::compartment_terms(c, compartment)             optional model hook (floats)
// A model whose update starts with a function of the local concentration only
// (Monod: mu = mu_max*s/(k_s+s)) hoists it here: the IEEE division then runs once
// per compartment instead of once per particle, with bit-identical results.
// Small tables are built by every cycle block in shared memory (no extra launch);
// pre_step builds large ones in global memory.
//
// Leave threshold.  The reference tests (dt*flow/volume) > (double)u1 with a float uniform u1
// (move_kernel.hpp:407-411).  u1 = n * 2^-24 with n = the top 24 bits of a random word, so
//     (dt*F/V) > (double)u1  <=>  u1 < ceil_f32(dt*F/V) =: t  <=>  n < t * 2^24  <=>  n < ceil(t * 2^24)
// (t * 2^24 is exact in single precision).  The table holds that integer, clamped to [0, 2^24], in
// bits 0..24: the per-particle test is a shift and an integer compare, no conversion.  Bit 31
// (kOutletBit) marks a compartment whose first matching outlet (find_flow, move_kernel.hpp:113-127)
// has a non-zero flow, so the particle pass needs no loop over the outlet list either.
// -----------------------------------------------------------------------------
constexpr uint32_t kThrMask = 0x01ffffffu;
constexpr uint32_t kOutletBit = 0x80000000u;
struct PreParams {
  const double* diag; const double* vol; double dt; const double* conc; uint32_t n_species; float* ctab; uint32_t n_comp;
  int enable_move;
  int n_flows; Outlet outlets[kMaxFlows];
};

__device__ __forceinline__ uint32_t leave_threshold(double dt, double diag, double vol) {
  const float t = __double2float_ru(dt * diag / vol);
  if (!(t > 0.0f)) return 0u;          // also NaN (0/0): the comparison with a NaN is false for every particle
  if (t >= 1.0f) return 1u << 24;      // every n < 2^24 passes
  return (uint32_t)ceilf(t * 16777216.0f);
}

template <class M> __device__ __forceinline__ void compartment_row(const double* diag, const double* vol, double dt, const double* conc,
                                                                   uint32_t n_species, int enable_move, const Outlet* outlets, int n_flows,
                                                                   uint32_t c, float* row) {
  uint32_t w0 = enable_move ? leave_threshold(dt, diag[c], vol[c]) : 0u;
  for (int f = 0; f < n_flows; ++f)
    if (outlets[f].index == c) { if (outlets[f].flow != 0.) w0 |= kOutletBit; break; }
  row[0] = __uint_as_float(w0);
  if constexpr (M::n_pre > 0) M::compartment_terms(ConcView{conc, n_species, nullptr}, (size_t)c, row + 1);
}

template <class M> __device__ __forceinline__ void pre_step_body(const PreParams& p) {
  constexpr int CT = 1 + M::n_pre;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  for (uint32_t c = i; c < p.n_comp; c += nthreads) {
    float row[CT];
    compartment_row<M>(p.diag, p.vol, p.dt, p.conc, p.n_species, p.enable_move, p.outlets, p.n_flows, c, row);
#pragma unroll
    for (int k = 0; k < CT; ++k) p.ctab[(size_t)c * CT + k] = row[k];
  }
}

// -----------------------------------------------------------------------------
// post_cycle: everything of SimulationUnit::post_cycle (simulation.hpp:213-239) after the
// particle pass.  It is the SECOND PHASE OF THE STEP KERNEL (cooperative launch: every block is
// resident, a grid barrier separates it from the particle pass), so a whole time step is one
// launch:
//
//   publish: scatter_contribute + synchro_sources (simulation.cpp:143-151, implScalar.cpp:194-205):
//     sources = accumulator, accumulator = 0
//   plan: update_and_remove_inactive (particles_container.hpp:539-557) and the merge_buffer size
//     (:575-581) — a pure function of the device counters, evaluated redundantly by every block
//   [only when the plan says so]
//   compaction: remove_inactive_particles + CompactParticlesFunctor
//     (particles_container.hpp:735-796, 292-385), made exact and deterministic: the
//     k-th non-idle slot below new_n (ascending) receives the k-th idle particle of the tail
//     [new_n, old_n) counted from the end — the pairing a serial execution of the reference
//     functor produces.  Three phases separated by grid barriers:
//       count : per-tile counts (gaps below new_n, idle in the tail) + block-local prefixes
//       src   : tail tiles -> src[k] = slot of the k-th idle from the end
//       move  : low tiles  -> the gap with rank k pulls src[k]
//   insert: merge_buffer + InsertFunctor (particles_container.hpp:575-599, 403-443).  The newborn of
//     mother i goes to new_n + (number of dividing mothers with a smaller slot index) — the order
//     the reference's buffer has under serial execution.
//   commit: container counters, the next step's buffer room, one more entry of the age tables;
//     done by the last block to finish (ticket), when no block reads the counters any more.
//
// Blocks of 256 threads own contiguous ranges of 1024-slot tiles; thread t handles slots
// t, t+256, t+512, t+768 of a tile ("virtual warp" vw = 8*r + warp covers 32 consecutive slots).
// -----------------------------------------------------------------------------

// exclusive prefix of per-block totals in shared memory (n <= kMaxGrid): warp 0 scans 32
// entries per step with shuffles; executed by the whole block
__device__ __forceinline__ unsigned block_prefix_of(const uint32_t* blk_tot, unsigned nblk, unsigned b, unsigned* s_tmp,
                                                    unsigned& grand_total) {
  __syncthreads();  // s_tmp may still be read from a previous use
  for (unsigned k = threadIdx.x; k < nblk; k += blockDim.x) s_tmp[k] = __ldcg(blk_tot + k);  // one parallel pass
  __syncthreads();
  if (threadIdx.x < 32) {
    const unsigned lane = threadIdx.x;
    unsigned run = 0;
    for (unsigned base = 0; base < nblk; base += 32) {
      const unsigned k = base + lane;
      const unsigned v = k < nblk ? s_tmp[k] : 0u;
      unsigned tot;
      const unsigned ex = warp_excl_scan(v, tot);
      if (k < nblk) s_tmp[k] = run + ex;
      run += tot;
    }
    if (lane == 0) s_tmp[nblk] = run;
  }
  __syncthreads();
  grand_total = s_tmp[nblk];
  return s_tmp[b];
}

// Grid-wide barrier for cooperatively launched kernels (all blocks resident): arrive counter +
// generation word in DevState.
__device__ __forceinline__ void grid_barrier(DevState* st) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = *reinterpret_cast<volatile unsigned*>(&st->bar_gen);
    if (atomicAdd(&st->bar_count, 1u) == gridDim.x - 1) {
      st->bar_count = 0;
      __threadfence();
      atomicAdd(&st->bar_gen, 1u);
    } else {
      while (*reinterpret_cast<volatile unsigned*>(&st->bar_gen) == gen) __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

// InsertFunctor (particles_container.hpp:403-443): buffer row j -> container slot dst
__device__ __forceinline__ void insert_newborn(const PostParams& p, unsigned long long j, unsigned long long dst, uint32_t npos) {
  for (int c0 = 0; c0 < p.n_var; c0 += 8) {  // loads first, then stores: one round trip per 8 columns
    float t[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c0 + c < p.n_var) t[c] = __ldcg(p.buf_props + (size_t)(c0 + c) * p.buf_stride + j);
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c0 + c < p.n_var) p.props[(size_t)(c0 + c) * p.cap + dst] = t[c];
  }
  p.pos[dst] = npos;
  // both ages reset (eager: 0.f; stamped: the newborn ages from the next step on)
  reinterpret_cast<uint32_t*>(p.age_hyd)[dst] = p.newborn_stamp_hyd;
  reinterpret_cast<uint32_t*>(p.age_div)[dst] = p.newborn_stamp_div;
  p.status[dst] = (uint8_t)Idle;
}

// must be entered by every block of a cooperative launch, after a grid barrier that follows the
// last write to the particle state, the division buffer and the device counters
static __device__ __forceinline__ void post_cycle_body(const PostParams& p) {
  __shared__ unsigned s_pref[kMaxGrid + 1];
  __shared__ unsigned s_w[32];
  DevState* const st = p.st;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  // plan inputs (read-only on the counters: every block derives the same values; six threads per block read them,
  // so that the few counter lines are not hammered by every thread of the grid).  Issued before the publish loop:
  // the round trip overlaps it.
  __shared__ unsigned long long s_plan[6];
  unsigned long long plan_v = 0ull;
  if (threadIdx.x < 6) {
    const unsigned long long* src = threadIdx.x == 0 ? &st->step_exit : threadIdx.x == 1 ? &st->n_used : threadIdx.x == 2 ? &st->inactive
                                  : threadIdx.x == 3 ? &st->buf_index : threadIdx.x == 4 ? &st->buf_cap_eff : nullptr;
    plan_v = src ? __ldcg(src) : (unsigned long long)__ldcg(&st->force_compact);
  }
  // ---- publish the source terms of this step; the accumulator is left zeroed for the next one
  // (not from force_remove_dead: the sources of the last cycle stay what they are)
  if (p.count_step) {
    for (unsigned long long k = gtid; k < p.n_bins; k += gstride) {
      // fast kernel: integer bins hold sum(c_i * scale) exactly; the fp64 accumulator holds what did not go through them
      const long long fix = (long long)__ldcg(p.acc_fix + k);
      double s = __ldcg(p.acc + k);
      if (fix != 0) {
        const uint32_t j = (uint32_t)(k % p.n_species);
        s += (double)fix * (p.weight / (double)__ldcg(&st->src_scale[j < 8u ? j : 7u]));  // scale is a power of two: exact
        p.acc_fix[k] = 0ull;
      }
      p.sources[k] = s;
      p.acc[k] = 0.0;
      if (p.src_mirror)
        asm volatile("st.volatile.v2.u64 [%0], {%1, %2};" ::"l"(p.src_mirror + 2 * k), "l"((unsigned long long)__double_as_longlong(s)), "l"(p.src_tag) : "memory");
    }
  }
  BMC_STAMP(st, 5);
  // ---- plan
  if (threadIdx.x < 6) s_plan[threadIdx.x] = plan_v;
  __syncthreads();
  const unsigned long long out = s_plan[0];
  const unsigned long long n_before = s_plan[1];
  const unsigned long long inactive = s_plan[2] + out;  // inactive_counter += out; += dead (always 0, Q3)
  unsigned long long thr = (unsigned long long)((double)n_before * p.dead_ratio);
  if (p.min_removal > thr) thr = p.min_removal;
  const bool do_compact = (inactive > thr) || (s_plan[5] && inactive > 0);
  const unsigned long long old_n = n_before, new_n = do_compact ? n_before - inactive : n_before;
  const unsigned long long n_add = s_plan[3] < s_plan[4] ? s_plan[3] : s_plan[4];

  if (do_compact) {  // uniform across the grid
    // One WARP per 1024-slot tile: lane l owns slots [32 l, 32 l + 32) of the tile and reads their status bytes with two
    // 128-bit loads (1 KB per warp, fully coalesced); flags live in one 32-bit word per lane, counts are popcounts, ranks
    // inside a tile come from a warp scan.  No block-wide barrier per tile (the first version had two per tile and took
    // 1.6 ms to remove 5e4 particles from 1e8: 660 serial tiles per block, and the ~1000 tail tiles all owned by the
    // last two blocks).
    const uint32_t n_tiles = (uint32_t)((old_n + kTile - 1) / kTile);
    const unsigned G = gridDim.x, nwarp = blockDim.x >> 5;
    const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / G);
    const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / G);
    const unsigned lt = (1u << lane) - 1u;
    // bit b of the result: slot (tile, 32*lane + b) is a GAP (not idle, below new_n) / an idle particle of the TAIL
    auto lane_bits = [&](uint32_t tile, unsigned& gap, unsigned& tail) {
      const unsigned long long base = (unsigned long long)tile * kTile + 32ull * lane;
      const uint4* ptr = reinterpret_cast<const uint4*>(p.status + base);  // the capacity is a multiple of 1024: in range
      const uint4 a = ptr[0], c = ptr[1];
      const unsigned w[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      unsigned nonidle = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k)  // one bit per status byte: 0xff where the byte differs from Idle (0)
        nonidle |= ((((__vcmpne4(w[k], 0u)) & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
      auto below = [&](unsigned long long lim) -> unsigned {  // bits of the slots < lim
        return base + 32ull <= lim ? 0xffffffffu : (base >= lim ? 0u : ((1u << (unsigned)(lim - base)) - 1u));
      };
      const unsigned valid = below(old_n), low = below(new_n);
      gap = nonidle & valid & low;
      tail = ~nonidle & valid & ~low;
    };
    // ---- count: every block counts its own contiguous tile range (equal work per tile), then scans it ----
    for (uint32_t tile = t0 + warp; tile < t1; tile += nwarp) {
      unsigned gb, tb;
      lane_bits(tile, gb, tb);
      const unsigned g = __reduce_add_sync(0xffffffffu, (unsigned)__popc(gb)), i = __reduce_add_sync(0xffffffffu, (unsigned)__popc(tb));
      if (lane == 0) { p.tile_gap_off[tile] = g; p.tile_idle_off[tile] = i; }
    }
    __syncthreads();
    if (warp == 0) {  // counts -> block-local exclusive prefixes, block totals
      unsigned run_g = 0, run_i = 0;
      for (uint32_t bt = t0; bt < t1; bt += 32) {
        const uint32_t t = bt + lane;
        const unsigned g = t < t1 ? p.tile_gap_off[t] : 0u, i = t < t1 ? p.tile_idle_off[t] : 0u;
        unsigned tg, ti;
        const unsigned eg = warp_excl_scan(g, tg), ei = warp_excl_scan(i, ti);
        if (t < t1) { p.tile_gap_off[t] = run_g + eg; p.tile_idle_off[t] = run_i + ei; }
        run_g += tg; run_i += ti;
      }
      if (lane == 0) { p.blk_gap[blockIdx.x] = run_g; p.blk_idle[blockIdx.x] = run_i; }
    }
    grid_barrier(st);
    auto owner_of = [&](uint32_t tile) -> unsigned { return (unsigned)((((unsigned long long)tile + 1ull) * G - 1ull) / n_tiles); };
    const unsigned gwarp = blockIdx.x * nwarp + warp, gwarps = G * nwarp;
    // ---- src: k-th idle tail particle counted from the end; tail tiles spread over ALL warps of the grid ----
    unsigned total_idle;
    {
      block_prefix_of(p.blk_idle, G, blockIdx.x, s_pref, total_idle);
      const uint32_t first_tail_tile = (uint32_t)(new_n / kTile);
      for (uint32_t tile = first_tail_tile + gwarp; tile < n_tiles; tile += gwarps) {
        unsigned gb, tb;
        lane_bits(tile, gb, tb);
        unsigned tot;
        const unsigned ex = warp_excl_scan((unsigned)__popc(tb), tot);
        unsigned asc = s_pref[owner_of(tile)] + p.tile_idle_off[tile] + ex;
        while (tb) {
          const unsigned bit = __ffs(tb) - 1u;
          tb &= tb - 1u;
          p.src[total_idle - 1u - asc] = (uint32_t)((unsigned long long)tile * kTile + 32u * lane + bit);
          ++asc;
        }
      }
    }
    grid_barrier(st);
    // ---- move: gaps below new_n pull their replacement; low tiles spread over all warps of the grid ----
    {
      unsigned total_gap;
      block_prefix_of(p.blk_gap, G, blockIdx.x, s_pref, total_gap);
      const uint32_t last_low_tile = (uint32_t)((new_n + kTile - 1) / kTile);  // exclusive
      for (uint32_t tile = gwarp; tile < last_low_tile; tile += gwarps) {
        unsigned gb, tb;
        lane_bits(tile, gb, tb);
        if (!__any_sync(0xffffffffu, gb != 0u)) continue;  // most tiles have no gap
        unsigned tot;
        const unsigned ex = warp_excl_scan((unsigned)__popc(gb), tot);
        unsigned k = s_pref[owner_of(tile)] + p.tile_gap_off[tile] + ex;
        while (gb) {
          const unsigned bit = __ffs(gb) - 1u;
          gb &= gb - 1u;
          if (k >= total_idle) {
            atomicOr(&st->error, 2u);  // inactive counter inconsistent with the status column
          } else {
            const size_t i = (size_t)tile * kTile + 32u * lane + bit;
            const size_t s2 = p.src[k];
            p.status[i] = (uint8_t)Idle;
            p.pos[i] = p.pos[s2];
            for (int c = 0; c < p.n_var; ++c) p.props[(size_t)c * p.cap + i] = p.props[(size_t)c * p.cap + s2];
            p.age_hyd[i] = p.age_hyd[s2];
            p.age_div[i] = p.age_div[s2];
          }
          ++k;
        }
      }
    }
    grid_barrier(st);
    // slots [new_n, old_n) left the container: mark them Idle so that appended newborns never inherit
    // a stale status (the reference relies on zero-initialised storage, particles_container.hpp:403-443).
    // Newborn slots below are written Idle as well, so the two writers agree where they overlap.
    for (unsigned long long i = new_n + gtid; i < old_n; i += gstride) p.status[i] = (uint8_t)Idle;
  }

  BMC_STAMP(st, 6);
  // Newborn of mother i goes to new_n + (number of dividing mothers with a smaller slot index).
  constexpr unsigned kSmallAdd = kMaxGrid;  // records that fit the shared scratch: ranked by direct comparison
  if (n_add && n_add <= kSmallAdd) {  // uniform across the grid; the usual case (a few hundred divisions per step)
    // newborns are placed by the LAST blocks of the grid: the first ones are busy publishing the source terms
    const unsigned rblock = gridDim.x - 1u - blockIdx.x;
    if ((unsigned long long)rblock * blockDim.x < n_add) {  // blocks that have a newborn to place
      for (unsigned j = threadIdx.x; j < (unsigned)n_add; j += blockDim.x) s_pref[j] = __ldcg(p.buf_mother + j);
      __syncthreads();
      const unsigned long long j = (unsigned long long)rblock * blockDim.x + threadIdx.x;
      if (j < n_add) {
        const uint32_t mother = s_pref[j];
        const uint32_t npos = __ldcg(p.buf_pos + j);
        unsigned rank = 0;
        for (unsigned q = 0; q < (unsigned)n_add; ++q) rank += (s_pref[q] < mother) ? 1u : 0u;
        insert_newborn(p, j, new_n + rank, npos);
      }
    }
  } else if (n_add) {
    // Many divisions: counts per tile = popcount of the tile's 32 mask words (rewritten by the particle
    // pass for every group below n_used); block-local exclusive prefix over this block's contiguous tile
    // range -> tile_off, range total -> blk_total.  Then, behind a barrier, the global offsets.
    const unsigned G = gridDim.x;
    const uint32_t T = (uint32_t)((old_n + kTile - 1) / kTile);
    const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * T) / G);
    const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * T) / G);
    const unsigned long long gslots = 32ull * (unsigned long long)p.vec;
    const unsigned long long words_valid = ((old_n + gslots - 1ull) / gslots) * (unsigned long long)p.vec;  // words of groups beyond the last slot are stale
#pragma unroll 4
    for (uint32_t t = t0 + warp; t < t1; t += blockDim.x / 32) {
      const unsigned long long wi = (unsigned long long)t * (kTile / 32) + lane;
      const unsigned wv = wi < words_valid ? __ldcg(p.div_mask + wi) : 0u;
      const unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)__popc(wv));
      if (lane == 0) p.tile_off[t] = c;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned run = 0;
      for (uint32_t base = t0; base < t1; base += 32) {
        const uint32_t t = base + lane;
        const unsigned cnt = (t < t1) ? __ldcg(p.tile_off + t) : 0u;
        unsigned tot;
        const unsigned ex = warp_excl_scan(cnt, tot);
        if (t < t1) p.tile_off[t] = run + ex;
        run += tot;
      }
      if (lane == 0) p.blk_total[blockIdx.x] = run;
    }
    grid_barrier(st);
    if ((unsigned long long)blockIdx.x * blockDim.x < n_add) {  // blocks that have a newborn to place
      unsigned total;
      block_prefix_of(p.blk_total, G, 0, s_pref, total);
    }
    for (unsigned long long j = gtid; j < n_add; j += gstride) {
      const uint32_t mother = __ldcg(p.buf_mother + j);
      const uint32_t tile = mother >> 10;
      const unsigned b = (unsigned)((((unsigned long long)tile + 1ull) * G - 1ull) / T);  // owner block of the tile
      // rank of the mother among the dividing mothers of its tile: all 32 mask words in one round trip.
      // Layout (written by the particle pass): per group of 32*VEC slots VEC ballot words, word q bit l <-> slot
      // group*32*VEC + l*VEC + q.
      const uint4* w4 = reinterpret_cast<const uint4*>(p.div_mask + (size_t)tile * (kTile / 32));
      const unsigned in_tile = mother & (kTile - 1);
      const unsigned vec = (unsigned)p.vec, gsz = 32u * vec;
      const unsigned gi = in_tile / gsz, l = (in_tile % gsz) / vec, qm = in_tile % vec;
      uint4 w[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) w[q] = __ldcg(w4 + q);
      const unsigned toff = __ldcg(p.tile_off + tile);
      const uint32_t npos = __ldcg(p.buf_pos + j);
      unsigned rank = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned ww[4] = {w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const unsigned idx = (unsigned)(4 * q + c);       // word index inside the tile
          const unsigned wg = idx / vec, wq = idx % vec;    // its group and slot-in-thread
          // slots of word (wg, wq) below the mother: whole word for earlier groups; in the mother's group the lanes
          // l' < l, plus lane l itself when wq < qm
          const unsigned m = wg < gi ? 0xffffffffu : (wg == gi ? (((1u << l) - 1u) | (wq < qm ? (1u << l) : 0u)) : 0u);
          rank += __popc(ww[c] & m);
        }
      }
      insert_newborn(p, j, new_n + s_pref[b] + toff + rank, npos);
    }
  }
  BMC_STAMP(st, 7);
  // commit: the last block to get here (every other block is done reading the counters and has published its bins).
  // Everything the commit needs is stable since the grid barrier, so EVERY block fetches it and does the arithmetic
  // before it takes its ticket (redundantly, but off the critical path): the commit itself is then a handful of
  // stores, not a chain of dependent L2 round trips at the very end of the step.
  __shared__ unsigned s_last;
  __shared__ float s_fb[8], s_fs[8];
  const unsigned long long n = new_n + n_add;
  unsigned long long pre_waiting = 0, pre_total_out = 0, pre_ncompact = 0, pre_total_new = 0, pre_step = 0, la = 0, lb = 0;
  unsigned pre_error = 0;
  float pre_tab_div = 0.f, pre_tab_hyd = 0.f;
  if (threadIdx.x == 0) {
    pre_waiting = __ldcg(&st->step_waiting); pre_total_out = __ldcg(&st->total_out); pre_ncompact = __ldcg(&st->n_compactions);
    pre_total_new = __ldcg(&st->total_new); pre_step = __ldcg(&st->step); la = __ldcg(&st->logical_alloc); lb = __ldcg(&st->logical_buf);
    pre_error = __ldcg(&st->error);
    if (p.tab_extend) { pre_tab_div = __ldcg(p.tab_div + p.tab_idx_div); if (p.enable_leave) pre_tab_hyd = __ldcg(p.tab_hyd + p.tab_idx_hyd); }
  }
  if (threadIdx.x < 8 && p.count_step) {
    // Fixed-point scatter of the particle pass: next step's admission bound and scale per species from this step's
    // largest |contribution| m, with m < 2^e_m and n <= 2^e_n particles:
    //   bound = 2^(e_m + 2)   a contribution may grow 4x from one step to the next before it takes the fp64 path
    //   scale = 2^(62 - e_n - e_m - 2)   so that n contributions at the bound stay below 2^62
    const int j = (int)threadIdx.x;
    const int e_n = n > 1ull ? 64 - __clzll((long long)(n - 1ull)) : 0;
    const float m = __uint_as_float(__ldcg(&st->src_max[j]));
    float bound = 0.0f, scale = 1.0f;  // nothing seen: only exact zeros are admitted (they add nothing)
    if (j < p.n_c && m > 0.0f && m < 3.0e38f) {
      int e_m;
      (void)frexpf(m, &e_m);  // m = f * 2^e_m, 0.5 <= f < 1
      const int kexp = 62 - e_n - e_m - 2;
      if (kexp >= -120 && kexp <= 120 && e_m + 2 <= 120 && e_m >= -120) { bound = ldexpf(1.0f, e_m + 2); scale = ldexpf(1.0f, kexp); }
    }
    s_fb[j] = bound; s_fs[j] = scale;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&st->done_blocks, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (p.count_step && p.px.world > 1 && p.px.publish_epoch) p2p_publish(p.px, p.sources, p.n_bins);  // whole block
  if (threadIdx.x < 8 && p.count_step) { st->src_max[threadIdx.x] = 0u; st->src_bound[threadIdx.x] = s_fb[threadIdx.x]; st->src_scale[threadIdx.x] = s_fs[threadIdx.x]; }
  if (threadIdx.x == 0) {
    st->done_blocks = 0; st->next_group = 0;
    if (p.count_step) { st->last_out = out; st->last_dead = 0; st->last_waiting = pre_waiting; }  // a forced compaction is not a step
    st->total_out = pre_total_out + out;
    st->step_exit = 0; st->step_waiting = 0; st->buf_index = 0; st->force_compact = 0;
    st->inactive = do_compact ? 0ull : inactive;
    if (do_compact) st->n_compactions = pre_ncompact + 1ull;
    st->do_compact = do_compact ? 1u : 0u; st->cmp_old_n = old_n; st->cmp_new_n = new_n; st->n_add = n_add;  // for inspection
    st->n_used = n;
    st->total_new = pre_total_new + n_add;
    const unsigned long long step_now = pre_step + (unsigned long long)p.count_step;
    st->step = step_now;
    // the reference's extents after this post_cycle: shrink after a compaction (remove_inactive_particles,
    // particles_container.hpp:784-797), _resize + __allocate_buffer__ in merge_buffer (:575-599, returns early
    // without newborns)
    if (s_plan[3] > s_plan[4] && s_plan[4] < lb) { atomicOr(&st->error, kErrCapacity); pre_error |= kErrCapacity; }  // divisions refused by the PHYSICAL room
    if (do_compact && new_n != 0ull && new_n <= (unsigned long long)(p.shrink_ratio * (double)la)) {
      const unsigned long long ns = (unsigned long long)((double)new_n * p.allocation_factor);
      if (ns > 0ull) la = (unsigned long long)ceil((double)ns * p.allocation_factor);  // _resize(n_used * factor, force)
    }
    if (n_add) logical_grow(n, p.allocation_factor, p.buffer_ratio, la, lb);
    st->logical_alloc = la; st->logical_buf = lb;
    // room of the next step's division buffer: the logical extent, bounded by the physical arrays (the device can
    // never write past them; the host keeps them ahead, and kErrCapacity reports it if it did not)
    const unsigned long long room = p.cap > n ? p.cap - n : 0ull;
    unsigned long long eff = lb < p.buf_cap ? lb : p.buf_cap;
    st->buf_cap_eff = eff < room ? eff : room;
    if (p.pin) pin_write(p.pin, step_now, n, n_add, la, lb, pre_error);  // host mirror (zero-copy)
    if (p.tab_extend) {  // A[k+1] = fl(A[k] + d_t): exactly the accumulation an eagerly updated age goes through
      p.tab_div[p.tab_idx_div + 1] = pre_tab_div + p.dt_f;                                  // model_kernel.hpp:191 (float d_t)
      if (p.enable_leave) p.tab_hyd[p.tab_idx_hyd + 1] = (float)((double)pre_tab_hyd + p.dt);  // move_kernel.hpp:596 (double d_t)
    }
  }
}

// -----------------------------------------------------------------------------
// cycle: the fused hot kernel.
//   model   : CycleFunctor::operator()(TagCycle) + exec_per_particle
//             (model_kernel.hpp:163-217, 230-268), handle_division
//             (particles_container.hpp:559-573)
//   contribs: ContributionFunctor Tag3D/Tag0D (contribution_kernel.hpp:48-186)
//   move    : MoveFunctor TagMove + handle_move + __find_next_compartment +
//             probability_leaving<fast_tag> (move_kernel.hpp:61-103, 209-273,
//             392-437; probability_leaving.hpp:33-46)
//   leave   : MoveFunctor TagLeave + handle_exit + find_flow +
//             probability_leaving<precision_tag> (move_kernel.hpp:105-127,
//             347-359, 586-648; probability_leaving.hpp:16-30)
// Order per particle = model -> contribution (pre-move position) -> move ->
// leave (post-move position), identical to the reference's kernel order because
// particles only interact through the atomically allocated division buffer and
// the additive source terms.
//
// Structure.  A warp works on a GROUP of 32*VEC consecutive slots (lane l owns slots
// l*VEC .. l*VEC+VEC-1: 128-bit column accesses).  The per-group body is straight-line code
// over the VEC particles of a thread — load, model update, fixed-point scatter, division vote,
// the two integer tests "leaves its compartment" / "sits in a compartment with an outlet",
// write-back — so that the dependency chains of the VEC particles interleave.  Everything a
// particle does RARELY is taken out of that body:
//   * division (a warp vote guards it; a few groups per thousand);
//   * move + leave.  About 1 % of the particles change compartment in a step and 1/n_comp of them
//     sit in an outlet compartment, but with 128 slots per warp three groups out of four contain at
//     least one such particle, and a divergent branch costs the whole warp.  The body therefore only
//     QUEUES the slot index of a candidate (warp-private queue in shared memory, ballot + popcount);
//     whenever 32 candidates are waiting the warp handles them one per lane (`deferred`): Philox
//     draws, CDF search, outlet test, position / status / age-stamp updates — everything recomputed
//     from the slot index, because every random draw is a pure function of (slot, step).
// -----------------------------------------------------------------------------
__host__ __device__ constexpr int popcount_c(uint64_t x) { return x == 0u ? 0 : (int)(x & 1u) + popcount_c(x >> 1); }
__host__ __device__ constexpr bool col_flag(uint64_t mask, int k) { return ((mask >> k) & 1ull) != 0ull; }  // k < 64
// columns actually loaded per slot: every property that is not write-only
template <class M> struct ReadCols {
  static_assert(M::n_var <= 64, "at most 64 properties per particle");
  static constexpr uint64_t all = M::n_var >= 64 ? ~0ull : ((1ull << M::n_var) - 1ull);
  static constexpr int value = M::n_var - popcount_c((uint64_t)M::write_only_mask & all);
};
// entries of a warp's deferred queue: one group can add 32*VEC candidates to at most 31 waiting ones
__host__ __device__ constexpr uint32_t queue_entries(int vec) { return 32u * (uint32_t)vec + 32u; }

// Ampere-style asynchronous copies (LDGSTS): global -> shared without a register in between, completion per thread
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int VEC> struct MaskIO;
template <> struct MaskIO<4> { static __device__ __forceinline__ void st(uint32_t* p, const unsigned (&w)[4]) { *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]); } };
template <> struct MaskIO<2> { static __device__ __forceinline__ void st(uint32_t* p, const unsigned (&w)[2]) { *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]); } };
template <> struct MaskIO<1> { static __device__ __forceinline__ void st(uint32_t* p, const unsigned (&w)[1]) { *p = w[0]; } };

template <class M, int VEC, bool LAZY, int BLOCK> __device__ __forceinline__ void cycle_body(const CycleParams& p) {
  constexpr int kWarps = BLOCK / 32;
  constexpr int NV = M::n_var, NC = M::n_c, CT = 1 + M::n_pre;
  constexpr uint32_t kGroup = 32 * VEC;         // slots per group: the work unit of one warp
  constexpr unsigned kFull = 0xffffffffu;
  // Rows of 8 words (two doubles among the model terms, simple_acetate): in shared memory the table is kept PLANAR
  // (word k of compartment c at [k * n_comp + c]) — 32 lanes gathering the same word of 32 random compartments then
  // spread over all 32 banks.  (Measured on the sa_ns workload of bench.py, 1.875e8 live particles: 32-byte rows fetched with two 128-bit gathers, which put
  // every lane on one of four bank groups, 2.25 ms per step; two planes of 16-byte half rows 2.02; word planes 1.94.)
  // The model terms are fetched right before the particle's update, so that the terms of the VEC particles of a thread
  // are not all live at once.  The global-memory form of the table (n_comp too large for shared memory) stays
  // row-major: one row = one sector.
  constexpr bool kPlanar = PlanarTable<M>::value && CT == 8;
  auto planar_word = [&](uint32_t c, int k) -> uint32_t { return (uint32_t)k * p.n_comp + c; };
  // dynamic shared memory: [n_species * n_comp] 64-bit source bins when bins_in_smem, the compartment table when
  // ctab_in_smem, one deferred queue per warp
  extern __shared__ __align__(128) unsigned long long s_dyn[];
  __shared__ unsigned s_cnt[4];                // move, exit, new, overflow of this block (rare events: counted with shared atomics)
  __shared__ unsigned s_fxmax[8];              // largest |contribution| per species seen by this block (float bits)
  // work distribution (see below): ticket counter + a ring of chunk descriptors
  constexpr unsigned kRing = 8;
  // groups a block draws from the device-wide counter at a time = one per warp.  (8-group chunks let the blocks run out
  // of work closer together — 802 -> 797 us per step at 1.25e8 particles — but the case without prefetch staging (10 000
  // compartments) then hung in the ring hand-off: with a chunk smaller than the number of warps, the fetcher of chunk
  // c + kRing can find the slot of chunk c free BEFORE the fetcher of chunk c has published it.  One chunk per round of
  // the block's warps rules that out: chunk c + kRing is 8 x kWarps tickets later, more than the warps can hold.)
  constexpr int kChunk = kWarps;
  __shared__ unsigned s_ticket;
  __shared__ unsigned s_chunk_base[kRing];     // first group of chunk c (slot c % kRing) ...
  __shared__ unsigned s_chunk_seq[kRing];      // ... valid when == c + 1
  __shared__ unsigned s_chunk_left[kRing];     // tickets of the slot's chunk not yet resolved (0 = slot free)

  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_used = (uint32_t)p.st->n_used;  // a context holds fewer than 2^32 slots
  // Work distribution: the slots are cut into groups of 32*VEC; every warp of the (persistent, fully
  // resident) grid starts with the group of its own index and then draws further groups dynamically, so
  // that blocks which start late or run on a slower SM simply process fewer groups.  Two levels: a
  // BLOCK draws chunks of kChunk (= kWarps) consecutive groups from the device-wide counter, its warps draw
  // single groups of the chunk with a shared-memory ticket.  (One L2 atomic per GROUP on one address
  // was the limiter of the whole pass: same-address atomics retire at ~0.8 per ns on a B200, i.e.
  // 78 125 groups of a 1e7-particle step could not be handed out in less than ~95 us — the time per
  // step divided by the number of groups was the same 1.2 ns for 1e7 and for 1.25e8 particles.)
  const uint32_t n_groups = (uint32_t)(((unsigned long long)n_used + kGroup - 1) / kGroup);
  const uint32_t total_warps = gridDim.x * kWarps;
  const uint32_t n_bins = p.n_species * p.n_comp;
  const bool single_comp = (p.n_comp == 1);
  const bool smem_bins = p.bins_in_smem && !single_comp;
  // bin k = low word s_bins[k], high word s_bins[n_bins + k]: the low words, which take (nearly) all the atomics, are
  // contiguous, so that the 32 lanes of a warp spread over all 32 banks (interleaved low/high words used 16 of them)
  unsigned int* const s_bins = reinterpret_cast<unsigned int*>(s_dyn);
  unsigned int* const s_bins_hi = s_bins + n_bins;
  uint32_t* const s_ctab = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_dyn) + p.ctab_offset);
  uint32_t* const s_queue = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(s_dyn) + p.queue_offset) + warp * queue_entries(VEC);
  // Prefetch (VEC == 4): while a warp computes group n, the columns of its group n+1 are already on their way into a
  // warp-private staging buffer (cp.async: no registers are held by loads in flight, which is what lets a 64-register
  // kernel keep two groups per warp in flight).  Every lane copies exactly the bytes it will read back itself, so the
  // only synchronisation is the lane's own cp.async.wait_group.  A compute-free kernel with this access pattern needs
  // ~48 KB of loads in flight per SM to saturate HBM (tools/dbg/streams_bw_occ.cu); without the prefetch the warps of
  // an SM spend only part of their time waiting for loads and fall short of that.
  const bool pf = VEC == 4 && p.stage_warp_bytes != 0u;
  unsigned char* const s_stage = reinterpret_cast<unsigned char*>(s_dyn) + p.stage_offset + (size_t)warp * p.stage_warp_bytes + lane * 16u;
  constexpr int kColBytes = 32 * 4 * VEC;  // one staged column of a group

  BMC_STAMP(p.st, 0);
#if defined(BMC_TIMELINE)
  if (threadIdx.x == 0) {  // per-block start of the kernel (step parity selects the half): words [4G + 8b + 4*(step&1) ..]
    unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    uint32_t* w = p.post.src + 4 * gridDim.x + 8 * blockIdx.x + 4 * (p.step & 1u);
    w[0] = (uint32_t)t_; w[1] = (uint32_t)(t_ >> 32);
  }
#endif
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
  if (threadIdx.x < 8) s_fxmax[threadIdx.x] = 0u;
  if (threadIdx.x < kRing) { s_chunk_seq[threadIdx.x] = 0u; s_chunk_left[threadIdx.x] = 0u; s_chunk_base[threadIdx.x] = 0u; }
  if (threadIdx.x == 0) s_ticket = 0u;
  if (smem_bins)
    for (uint32_t k = threadIdx.x; k < n_bins; k += BLOCK) s_dyn[k] = 0ull;
  if (p.ctab_in_smem) {
    for (uint32_t c = threadIdx.x; c < p.n_comp; c += BLOCK) {
      float row[CT];
      compartment_row<M>(p.diag, p.vol, p.dt, p.conc, p.n_species, p.enable_move, p.outlets, p.n_flows, c, row);
#pragma unroll
      for (int k = 0; k < CT; ++k) s_ctab[kPlanar ? planar_word(c, k) : c * CT + k] = __float_as_uint(row[k]);
    }
  }
  // Scatter (c): fixed-point accumulation.  sm_100a has a native shared-memory atomic for 32-bit integers only
  // (ATOMS.ADD; fp32/fp64/u64 adds are LDS + ATOMS.CAST.SPIN retry loops), so a contribution c enters its bin as the
  // 64-bit integer rn(c * 2^k): ATOMS.ADD on the low word, the carry out of it (from the value the atomic returns)
  // added with the high word by a second ATOMS.ADD.  Integer sums do not depend on the order of the additions: the
  // source terms are bit-identical from run to run and for any grid size.  Scale 2^k and admission bound per species
  // come from the previous step (commit thread of post_cycle_body): the bound is 4x the largest |contribution| seen
  // and k is chosen so that n_used contributions at the bound cannot overflow 2^62; a contribution at the bound keeps
  // >= 60 - log2(n_used) bits, i.e. more than its 24-bit mantissa for any population a GPU can hold, smaller ones are
  // rounded to 2^-k.  A contribution above the bound (first step, or a 4x jump from one step to the next) goes to the
  // fp64 accumulator with an L2 atomic instead; the published value is the sum of the two accumulators.
  float fx_bound[NC], fx_scale[NC], fx_max[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    fx_bound[j] = (smem_bins && j < 8) ? __ldcg(&p.st->src_bound[j < 8 ? j : 7]) : -1.0f;  // -1: nothing is admitted
    fx_scale[j] = j < 8 ? __ldcg(&p.st->src_scale[j < 8 ? j : 7]) : 1.0f;
    fx_max[j] = 0.0f;
  }
  __syncthreads();

  // multi-GPU: the all-reduce of the PREVIOUS step's sources is finished here, by block 0, while the other blocks are
  // already streaming particles (the dynamic work distribution absorbs the few microseconds block 0 arrives late)
  if (p.post.px.world > 1 && p.post.px.consume_epoch && blockIdx.x == 0) p2p_consume(p.post.px, p.sources, n_bins, &p.st->error);

  BMC_STAMP(p.st, 1);
  double acc0d[NC];  // single-compartment accumulation lives in registers
#pragma unroll
  for (int j = 0; j < NC; ++j) acc0d[j] = 0.0;
  const double w = (double)p.weight;  // `const double weight = get_weight(p)` contribution_kernel.hpp:179
  const BufRows bufrows{p.buf_props, p.buf_stride};
  // LAZY: the age columns hold step stamps instead of floats (see "Step-stamped ages" above)
  uint32_t* const stamps_hyd = reinterpret_cast<uint32_t*>(p.age_hyd);
  uint32_t* const stamps_div = reinterpret_cast<uint32_t*>(p.age_div);
  (void)stamps_hyd; (void)stamps_div;

  auto ctab_word0 = [&](uint32_t c) -> uint32_t {
    return p.ctab_in_smem ? s_ctab[kPlanar ? c : c * CT] : __ldg(reinterpret_cast<const uint32_t*>(p.ctab) + (size_t)c * CT);
  };

  // ---- move + leave of one queued slot (one per lane) ------------------------------------------------
  auto deferred = [&](const uint32_t slot) {
    uint32_t c = p.pos[slot];
    if (p.enable_move) {  // handle_move (all slots, no status check: move_kernel.hpp:392-437)
      uint32_t r[4];
      philox4x32_10_idx(slot >> 2, p.ph0, r);
      const uint32_t n1 = pick4(r, slot & 3u) >> 8;
      if (n1 < (ctab_word0(c) & kThrMask)) {  // (dt*flow/volume) > rng1
        philox4x32_10_idx(slot, p.ph2, r);   // the neighbour pick draws its own block
        const float u2 = u01f(r[0]);
        const float* row = p.cdf + (size_t)c * p.m;
        int left = 0, right = p.m - 1;
        while (left < right) {  // __find_next_compartment, move_kernel.hpp:87-95
          const int mid = (left + right) >> 1;
          if (u2 > __ldg(row + mid)) left = mid + 1; else right = mid;
        }
        c = __ldg(p.neigh + (size_t)c * p.m + left);
        p.pos[slot] = c;
        atomicAdd(&s_cnt[0], 1u);  // events.wrap_incr<Move>() (Q20: aggregated per block)
      }
    }
    // leave (Idle only, post-move position: move_kernel.hpp:347-359)
    if (p.enable_leave && p.status[slot] == (uint8_t)Idle) {
      int f = -1;
      for (int k = 0; k < p.n_flows; ++k)  // find_flow: first match wins
        if (p.outlets[k].index == c) { if (p.outlets[k].flow != 0.) f = k; break; }
      if (f >= 0) {
        uint32_t r[4];
        philox4x32_10_idx(slot >> 2, p.ph1, r);
        const float u3 = u01f(pick4(r, slot & 3u));
        const float lnu = (float)log((double)u3);  // Kokkos::log(float), see oracle ln_f32
        if (p.outlets[f].dt_flow > (double)(-lnu) * p.outlets[f].volume) {  // probability_leaving<precision_tag>
          if constexpr (LAZY) {
            // the particle stops ageing: freeze the step counts (age_hyd = 0, age_div as of this step)
            stamps_hyd[slot] = kFrozen;
            const uint32_t sd = stamps_div[slot];
            stamps_div[slot] = (sd & kFrozen) ? sd : (kFrozen | (p.step + 1u - sd));
          } else {
            p.age_hyd[slot] = p.age_hyd[slot] * 0.0f;  // ages(idx,0) *= (1 - leave_mask), after this step's increment
          }
          p.status[slot] = (uint8_t)Exit;
          atomicAdd(&s_cnt[1], 1u);
        }
      }
    }
  };
  uint32_t qn = 0;  // entries waiting in this warp's queue (warp-uniform)

  // Every group but the last is entirely below n_used: the body is instantiated twice so that the
  // common case carries no per-slot range checks (FULL), the ragged tail keeps them.
  // issue the asynchronous copies of group g's columns into this warp's staging buffer (order: pos, [ages], read props, status)
  auto stage_issue = [&](const uint32_t g) {
    const uint32_t i_raw = g * kGroup + lane * VEC;
    const uint32_t i0 = i_raw < n_used ? i_raw : 0u;  // dead lanes of the ragged last group shadow slot 0
    unsigned char* dst = s_stage;
    cp_async16(dst, p.pos + i0); dst += kColBytes;
    if constexpr (!LAZY) {
      cp_async16(dst, p.age_div + i0); dst += kColBytes;
      if (p.enable_leave) { cp_async16(dst, p.age_hyd + i0); dst += kColBytes; }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (!col_flag(M::write_only_mask, k)) { cp_async16(dst, p.props + (size_t)k * p.cap + i0); dst += kColBytes; }
    cp_async4(dst - lane * 12u, p.status + i0);  // status: 4 bytes per lane, packed behind the columns
    cp_async_commit();
  };

  // `mid` runs once per group after the model update has consumed every loaded value: it resolves the warp's next
  // group and (with the prefetch) starts its copies, which then overlap the rest of this group
  auto body = [&](auto full_tag, const uint32_t g, auto&& mid) {
    constexpr bool FULL = decltype(full_tag)::value;
    const uint32_t i_raw = g * kGroup + lane * VEC;
    const bool live = FULL || i_raw < n_used;    // false only in the ragged end of the last group
    const uint32_t i0 = live ? i_raw : 0u;       // dead lanes shadow slot 0 (loads stay in range, nothing is stored)

    uint32_t pos[VEC]; float adiv[VEC], ahyd[VEC]; float v[VEC][NV], old[VEC][NV];
    uint32_t stw;
    if (pf) {
      // ---- the columns were staged by this lane's own asynchronous copies, issued one group ago ----
      cp_async_wait_all();
      const unsigned char* src = s_stage;
      VecIO<VEC>::ldu_plain(reinterpret_cast<const uint32_t*>(src), pos); src += kColBytes;
      if constexpr (!LAZY) {
        VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src), adiv); src += kColBytes;
        if (p.enable_leave) { VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src), ahyd); src += kColBytes; }
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float col[VEC];
        if (col_flag(M::write_only_mask, k)) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) col[q] = 0.f;
        } else {
          VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src), col); src += kColBytes;
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) { v[q][k] = col[q]; old[q][k] = col[q]; }
      }
      stw = *reinterpret_cast<const uint32_t*>(src - lane * 12u);
    } else {
      // ---- front-batched global loads (all independent) ----
      stw = VecIO<VEC>::ldb(p.status + i0);
      VecIO<VEC>::ldu(p.pos + i0, pos);
      if constexpr (!LAZY) {
        VecIO<VEC>::ldf(p.age_div + i0, adiv);
        if (p.enable_leave) VecIO<VEC>::ldf(p.age_hyd + i0, ahyd);
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float col[VEC];
        if (col_flag(M::write_only_mask, k)) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) col[q] = 0.f;
        } else {
          VecIO<VEC>::ldf(p.props + (size_t)k * p.cap + i0, col);
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) { v[q][k] = col[q]; old[q][k] = col[q]; }
      }
    }
    if (LAZY || !p.enable_leave) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) ahyd[q] = 0.f;
    }
    if constexpr (LAZY) {
#pragma unroll
      for (int q = 0; q < VEC; ++q) adiv[q] = 0.f;
    }
    bool idle[VEC];
    unsigned valid_m = 0, idle_m = 0;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const bool valid = FULL || (live && (i0 + q) < n_used);
      idle[q] = valid && (((stw >> (8 * q)) & 0xffu) == (unsigned)Idle);
      valid_m |= (unsigned)valid << q; idle_m |= (unsigned)idle[q] << q;
      if (!FULL && !valid) pos[q] = 0;  // slots past n_used hold unspecified bytes: keep the gathers in range
    }
    constexpr unsigned kAll = (1u << VEC) - 1u;

    // ---- compartment rows: leave threshold / outlet flag + model terms, one gather per particle
    constexpr bool kLateTerms = kPlanar;  // only word 0 is fetched here
    uint32_t cw0[VEC]; float cterm[VEC][CT];  // cterm[q][1..] = compartment_terms (index 0 unused)
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      if (p.ctab_in_smem) {
        const uint32_t* row = s_ctab + pos[q] * CT;
        if constexpr (CT == 2) { const uint2 t = *reinterpret_cast<const uint2*>(row); cw0[q] = t.x; cterm[q][1] = __uint_as_float(t.y); }
        else if constexpr (kLateTerms) cw0[q] = s_ctab[pos[q]];
        else {
          cw0[q] = row[0];
#pragma unroll
          for (int k = 1; k < CT; ++k) cterm[q][k] = __uint_as_float(row[k]);
        }
      } else {
        const uint32_t* row = reinterpret_cast<const uint32_t*>(p.ctab) + (size_t)pos[q] * CT;
        if constexpr (CT == 2) { const uint2 t = __ldg(reinterpret_cast<const uint2*>(row)); cw0[q] = t.x; cterm[q][1] = __uint_as_float(t.y); }
        else if constexpr (CT == 4) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(row));
          cw0[q] = t.x; cterm[q][1] = __uint_as_float(t.y); cterm[q][2] = __uint_as_float(t.z); cterm[q][3] = __uint_as_float(t.w);
        } else if constexpr (kLateTerms) cw0[q] = __ldg(row);
        else {
          cw0[q] = __ldg(row);
#pragma unroll
          for (int k = 1; k < CT; ++k) cterm[q][k] = __uint_as_float(__ldg(row + k));
        }
      }
    }

    // ---- u1: ONE Philox block per group of four slots (draw_block 0) ---------
    uint32_t rw1[4] = {0u, 0u, 0u, 0u};
    if (p.enable_move) philox4x32_10_idx((uint32_t)(i0 >> 2), p.ph0, rw1);

    // ---- move / leave candidates: two integer tests per slot, the work itself is deferred ----
    //   word 0 of the compartment row = leave threshold (bits 0..24) | kOutletBit
    unsigned cand = 0;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      const uint32_t n1 = (VEC == 4 ? rw1[q] : pick4(rw1, (unsigned)((i0 + q) & 3))) >> 8;
      const bool leaves = n1 < (cw0[q] & kThrMask);                 // (dt*flow/volume) > rng1, all slots
      const bool in_outlet = (cw0[q] & kOutletBit) != 0u && idle[q];  // leave test applies to Idle particles only
      cand |= (unsigned)(leaves || in_outlet) << q;
    }
    cand &= valid_m;

    // ---- model update: unconditional straight-line code over the VEC particles;
    // results of non-idle slots are never stored (model_kernel.hpp:186-196)
    float contrib[VEC][NC];
    unsigned div_nib = 0;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
      if constexpr (!LAZY) adiv[q] = idle[q] ? adiv[q] + p.dt_f : adiv[q];  // ages(i,1) += _d_t  (model_kernel.hpp:191)
      Gen gen(p.seed_lo, p.seed_hi, p.rank, (uint32_t)(i0 + q), p.step, 2u);
      if constexpr (kLateTerms) {
        if (p.ctab_in_smem) {
#pragma unroll
          for (int k = 1; k < CT - 1; ++k) cterm[q][k] = __uint_as_float(s_ctab[planar_word(pos[q], k)]);  // word 7 is padding
          cterm[q][CT - 1] = 0.f;
        } else {
          const uint4* row = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(p.ctab) + (size_t)pos[q] * CT);
          const uint4 t = __ldg(row), u = __ldg(row + 1);
          cterm[q][1] = __uint_as_float(t.y); cterm[q][2] = __uint_as_float(t.z); cterm[q][3] = __uint_as_float(t.w);
          cterm[q][4] = __uint_as_float(u.x); cterm[q][5] = __uint_as_float(u.y); cterm[q][6] = __uint_as_float(u.z); cterm[q][7] = __uint_as_float(u.w);
        }
      }
      const ConcView conc{p.conc, p.n_species, &cterm[q][1]};
      const Status s = M::update(gen, p.dt_f, i0 + q, RegRow{v[q]}, RegRow{contrib[q]}, (size_t)pos[q], conc);
      div_nib |= (unsigned)(idle[q] && s == Division) << q;
    }
    if constexpr (VEC == 4) mid();  // every loaded value has been consumed: the staging buffer is free for the next group

    // ---- contribution scatter at the PRE-move position (Q15) -----------------
    if (single_comp) {
#pragma unroll
      for (int q = 0; q < VEC; ++q)
#pragma unroll
        for (int j = 0; j < NC; ++j) acc0d[j] += idle[q] ? w * (double)contrib[q][j] : 0.0;
    } else {
      unsigned slow = 0;  // slots with a contribution that was not admitted to the integer bins (rare)
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
#pragma unroll
        for (int j = 0; j < NC; ++j) {
          const float c = contrib[q][j];
          const float a = idle[q] ? fabsf(c) : 0.0f;
          fx_max[j] = fmaxf(fx_max[j], a);
          const bool adm = idle[q] && a <= fx_bound[j];
          if (adm) {
            const long long fix = __float2ll_rn(c * fx_scale[j]);
            const uint32_t lo = (uint32_t)fix, hi = (uint32_t)((unsigned long long)fix >> 32);
            const uint32_t b = (uint32_t)j + p.n_species * pos[q];
            const uint32_t was = atomicAdd(s_bins + b, lo);
            atomicAdd(s_bins_hi + b, hi + (uint32_t)((uint32_t)(was + lo) < lo));
          }
          slow |= (unsigned)(idle[q] && !adm) << q;
        }
      }
      if (slow) {  // first step, a 4x jump, NaN, or bins too large for shared memory: fp64 accumulator in L2 (RED.F64)
#pragma unroll
        for (int q = 0; q < VEC; ++q)
          if ((slow >> q) & 1u) {
#pragma unroll
            for (int j = 0; j < NC; ++j)
              if (!(fabsf(contrib[q][j]) <= fx_bound[j]))
                atomicAdd(p.acc + (size_t)j + (size_t)p.n_species * pos[q], w * (double)contrib[q][j]);
          }
      }
    }

    // ---- division: handle_division (particles_container.hpp:559-573) -------
    // warp-aggregated slot allocation: ONE atomic per warp that has a dividing
    // mother (reference: one per mother, a6).  Rows are allocated in ascending
    // particle order inside the warp; final newborn placement is re-ranked by
    // mother index in the post-cycle phase, so the result does not depend on the
    // order warps hit the atomic.
    // Division mask: VEC ballot words per group (word q, bit l <-> slot g*32*VEC + l*VEC + q), written for EVERY
    // group of every step by lane 0 (0.125 B/slot), so no bit ever needs clearing.
    unsigned mask_w[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) mask_w[q] = 0u;
    if (__any_sync(kFull, div_nib != 0u)) {
      unsigned ok_nib = 0;
      const unsigned long long buf_cap = __ldcg(&p.st->buf_cap_eff);
      const unsigned cnt = __popc(div_nib);
      unsigned total;
      const unsigned excl = warp_excl_scan(cnt, total);
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(&p.st->buf_index, (unsigned long long)total);
      base = __shfl_sync(kFull, base, 0);
      unsigned r = 0;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        if ((div_nib >> q) & 1u) {
          const unsigned long long j = base + excl + r;
          ++r;
          atomicAdd(&s_cnt[2], 1u);  // NewParticle++ even on overflow (model_kernel.hpp:259, Q6)
          if (j < buf_cap) {
            Gen gen(p.seed_lo, p.seed_hi, p.rank, (uint32_t)(i0 + q), p.step, 0x40000000u);
            M::division(gen, i0 + q, (size_t)j, RegRow{v[q]}, bufrows);
            p.buf_pos[j] = pos[q];               // buffer_position(idx2) = position(idx1): pre-move
            p.buf_mother[j] = (uint32_t)(i0 + q);
            if constexpr (LAZY) stamps_div[i0 + q] = p.step + 1u;  // ages(idx1,1) = 0: counts from the next step
            else adiv[q] = 0.f;                                    // ages(idx1,1) = 0
            ok_nib |= 1u << q;
          } else {
            atomicAdd(&s_cnt[3], 1u);  // waiting_allocation_particle / Overflow (model_kernel.hpp:253-258)
          }
        }
      }
#pragma unroll
      for (int q = 0; q < VEC; ++q) mask_w[q] = __ballot_sync(kFull, (ok_nib >> q) & 1u);
    }
    if (lane == 0) MaskIO<VEC>::st(p.div_mask + (size_t)g * VEC, mask_w);

    // ---- write back ---------------------------------------------------------
    // Columns the model assigns on every update (always_written_mask, which includes the
    // write-only ones) are stored whenever the thread has an idle slot; the others only if a
    // value changed bitwise (the comparison folds away for columns the hooks never assign).
    // Position, status and (stamped) ages are written by `deferred` / the division branch only.
    const bool all_idle = (idle_m == kAll);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float col[VEC];
#pragma unroll
      for (int q = 0; q < VEC; ++q) col[q] = v[q][k];
      float* dst = p.props + (size_t)k * p.cap + i0;
      bool ch = false;
      if (col_flag(M::write_only_mask | M::always_written_mask, k)) ch = idle_m != 0u;
      else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) ch = ch || (idle[q] && __float_as_uint(v[q][k]) != __float_as_uint(old[q][k]));
      }
      if (ch) {
        if (all_idle) VecIO<VEC>::stf(dst, col);
        else {  // group with exited / out-of-range slots: their columns stay untouched
#pragma unroll
          for (int q = 0; q < VEC; ++q) if (idle[q]) dst[q] = col[q];
        }
      }
    }
    if constexpr (!LAZY) {
      // eager ages change for every idle particle; non-idle lanes rewrite the value they loaded
      if (p.enable_leave) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) ahyd[q] = idle[q] ? (float)((double)ahyd[q] + p.dt) : ahyd[q];  // ages(idx,0) += d_t (double)
      }
      if (idle_m) {
        VecIO<VEC>::stf(p.age_div + i0, adiv);
        if (p.enable_leave) VecIO<VEC>::stf(p.age_hyd + i0, ahyd);
      }
    }

    // ---- queue the candidates; handle 32 at a time ----------------------------
    if (__any_sync(kFull, cand != 0u)) {
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const bool mine = (cand >> q) & 1u;
        const unsigned b = __ballot_sync(kFull, mine);
        if (mine) s_queue[qn + __popc(b & lt)] = (uint32_t)(i0 + q);
        qn += __popc(b);
      }
      __syncwarp();  // queue entries and this group's age / stamp stores are visible to the whole warp
      while (qn >= 32u) {
        qn -= 32u;
        const uint32_t slot = s_queue[qn + lane];
        __syncwarp();
        deferred(slot);
      }
    }
  };

  {
    // Ticket t of a block -> chunk c = t / kChunk, offset o = t % kChunk.  The warp that draws o == 0 fetches the
    // chunk from the device-wide counter (the only L2 atomic: one per kChunk groups) and publishes its first group
    // in ring slot c % kRing; the other warps of the chunk read it when they need their group (normally long after).
    // A slot is reused only when every ticket of its previous chunk has been resolved (s_chunk_left), so a warp can
    // never read the base of a later chunk.  Every warp holds at most one unresolved ticket and resolves it before
    // it draws again or leaves the loop, which makes the scheme deadlock-free.
    volatile unsigned* const v_seq = s_chunk_seq;
    volatile unsigned* const v_left = s_chunk_left;
    volatile unsigned* const v_base = s_chunk_base;
    // Static head, dynamic tail.  The tickets cost ~45 instructions per group and thread; most of the pass does not need
    // them: the first `rounds` groups of a warp are its grid-stride groups w, w + W, w + 2W ... (W = warps of the grid:
    // the groups in flight still form one window moving through the columns), and only the last groups — an eighth of
    // the pass, at least four per warp — are drawn dynamically, which is what evens out blocks that started late or
    // run slower.  rounds == 1 is the fully dynamic scheme.
    const uint32_t per_warp = n_groups / total_warps;
    const uint32_t dyn = max(p.dyn_min, per_warp >> p.dyn_shift);
    const uint32_t rounds = per_warp > dyn ? per_warp - dyn : 1u;
    const uint32_t n_static = rounds * total_warps;  // groups below this index are assigned statically
    uint32_t s = blockIdx.x * kWarps + warp;  // first group: the warp's own index
    if (pf && s < n_groups) stage_issue(s);
#pragma unroll 1
    while (s < n_groups) {
      const bool stat = s + total_warps < n_static;  // warp-uniform: the next group is this warp's next grid-stride group
      unsigned t = 0;
      if (!stat && lane == 0) {  // draw the ticket of the NEXT group now, resolve it in the middle of this group
        t = atomicAdd(&s_ticket, 1u);
        const unsigned c = t / kChunk, o = t - c * kChunk, slot = c % kRing;
        if (o == 0u) {
          while (v_left[slot] != 0u) { }  // previous chunk of this slot still has unresolved tickets (practically never)
          const unsigned base = n_static + atomicAdd(&p.st->next_group, (unsigned)kChunk);
          v_base[slot] = base;
          v_left[slot] = (unsigned)kChunk;
          __threadfence_block();
          v_seq[slot] = c + 1u;
        }
      }
      if constexpr (VEC == 4) {
        uint32_t s_next = 0;
        auto resolve = [&]() {
          if (stat) { s_next = s + total_warps; return; }
          if (lane == 0) {
            const unsigned c = t / kChunk, o = t - c * kChunk, slot = c % kRing;
            while (v_seq[slot] != c + 1u) { }
            __threadfence_block();
            s_next = v_base[slot] + o;
            atomicSub(&s_chunk_left[slot], 1u);
          }
          s_next = __shfl_sync(kFull, s_next, 0);
        };
        // with the prefetch the next group must be known in the middle of this one (its copies overlap the rest of the
        // body); without it the ticket is resolved at the end, when the shared-memory atomic has long returned
        auto mid = [&]() {
          if (pf) {
            resolve();
            if (s_next < n_groups) stage_issue(s_next);
          }
        };
        if ((unsigned long long)(s + 1u) * kGroup <= n_used) body(FullTile{}, s, mid);
        else body(RaggedTile{}, s, mid);
        if (!pf) resolve();
        s = s_next;
      } else {
        auto nothing = []() {};
        if ((unsigned long long)(s + 1u) * kGroup <= n_used) body(FullTile{}, s, nothing);
        else body(RaggedTile{}, s, nothing);
        if (stat) { s += total_warps; continue; }
        if (lane == 0) {
          const unsigned c = t / kChunk, o = t - c * kChunk, slot = c % kRing;
          while (v_seq[slot] != c + 1u) { }
          __threadfence_block();
          s = v_base[slot] + o;
          atomicSub(&s_chunk_left[slot], 1u);
        }
        s = __shfl_sync(kFull, s, 0);
      }
    }
    __syncwarp();
    if (lane < qn) deferred(s_queue[lane]);  // what is left in the queue (< 32 entries)
  }

  BMC_STAMP(p.st, 2);
#if defined(BMC_TIMELINE)
  if (threadIdx.x == 0) {  // per-block end of the particle pass + number of tiles, into the (idle) compaction scratch
    unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.post.src[4 * blockIdx.x] = (uint32_t)t_; p.post.src[4 * blockIdx.x + 1] = (uint32_t)(t_ >> 32);
    p.post.src[4 * blockIdx.x + 2] = 0u; p.post.src[4 * blockIdx.x + 3] = smid;
  }
#endif
  // ---- block epilogue: counters, source flush ---------------------------------
#pragma unroll
  for (int j = 0; j < NC && j < 8; ++j) {  // non-negative floats order like their bit patterns
    const unsigned m = __reduce_max_sync(kFull, __float_as_uint(fx_max[j]));
    if (lane == 0 && m) atomicMax(&s_fxmax[j], m);
  }
  if (single_comp) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      double a = acc0d[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(kFull, a, o);
      if (lane == 0 && a != 0.0) atomicAdd(p.acc + j, a);
    }
  }
  __syncthreads();  // s_cnt, s_fxmax and the bins of every warp are complete (shared memory: no device-wide fence needed)
  if (threadIdx.x == 0) {
    const unsigned long long cm = s_cnt[0], ce = s_cnt[1], cn = s_cnt[2], co = s_cnt[3];
    if (cm) atomicAdd(&p.st->events[2], cm);                                             // Move
    if (ce) { atomicAdd(&p.st->events[1], ce); atomicAdd(&p.st->step_exit, ce); }        // Exit
    if (cn) atomicAdd(&p.st->events[0], cn);                                             // NewParticle
    if (co) { atomicAdd(&p.st->events[4], co); atomicAdd(&p.st->step_waiting, co); }     // Overflow
  }
  if (threadIdx.x < 8 && s_fxmax[threadIdx.x]) atomicMax(&p.st->src_max[threadIdx.x], s_fxmax[threadIdx.x]);
  if (smem_bins) {
    // every block starts at a different bin, so that the blocks (which finish together) do not hit
    // the same L2 addresses at the same time; integer adds (RED.ADD.64): order-independent
    const uint32_t rot = (uint32_t)(((unsigned long long)blockIdx.x * n_bins) / gridDim.x);
    for (uint32_t k0 = threadIdx.x; k0 < n_bins; k0 += BLOCK) {
      uint32_t k = k0 + rot; if (k >= n_bins) k -= n_bins;
      const unsigned long long a = (unsigned long long)s_bins[k] | ((unsigned long long)s_bins_hi[k] << 32);
      if (a != 0ull) atomicAdd(p.acc_fix + k, a);
    }
  }
  // ---- second phase of the step: every block's state, buffer rows and counters are complete
  BMC_STAMP(p.st, 3);
  if (p.fuse_post) {
    grid_barrier(p.st);
    BMC_STAMP(p.st, 4);
    post_cycle_body(p.post);
    BMC_STAMP(p.st, 8);
#if defined(BMC_TIMELINE)
    if (threadIdx.x == 0) {  // per-block end of the kernel
      unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
      uint32_t* w = p.post.src + 4 * gridDim.x + 8 * blockIdx.x + 4 * (p.step & 1u);
      w[2] = (uint32_t)t_; w[3] = (uint32_t)(t_ >> 32);
    }
#endif
  }
}

// -----------------------------------------------------------------------------
// mc_init_first: InitFunctor (mc/src/unit.cpp:102-144): M::init, random
// compartment, total-mass reduce.
// -----------------------------------------------------------------------------
struct InitParams {
  float* props; size_t cap; uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  unsigned long long n; uint32_t n_comp_hi; const float* linit; uint32_t seed_lo, seed_hi, rank; DevState* st;
};
template <class M> __device__ __forceinline__ void init_body(const InitParams& p) {
  double m = 0.0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float v[M::n_var];
    Gen gen(p.seed_lo, p.seed_hi, p.rank, (uint32_t)i, 0xFFFFFFFFu, 2u);
    M::init(gen, (size_t)i, RegRow{v}, ConfigView{p.linit});
    m += M::mass((size_t)i, RegRow{v});
    const uint32_t c = (uint32_t)gen.urand64(0ull, (unsigned long long)p.n_comp_hi);
#pragma unroll
    for (int k = 0; k < M::n_var; ++k) p.props[(size_t)k * p.cap + i] = v[k];
    p.pos[i] = c; p.status[i] = (uint8_t)Idle; p.age_hyd[i] = 0.f; p.age_div[i] = 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if ((threadIdx.x & 31) == 0 && m != 0.0) atomicAdd(&p.st->init_mass, m);
}

// -----------------------------------------------------------------------------
// get_properties: GetPropertiesFunctor (apps/core/public/core/post_process.hpp:33-118, 173-250).  For every
// Idle particle i of the chunk [first, first+count): particle_values(k, i) = property (double), row n_exp =
// M::mass; spatial_values(k, position) += value (per-compartment sums); ages.  `indices` selects the exported
// properties (HasExportPropertiesPartial, get_number()), nullptr = all (HasExportPropertiesFull).
// Non-idle particles keep zeros (the reference's views are zero-initialised and the functor returns early).
// -----------------------------------------------------------------------------
struct ExportParams {
  const float* props; size_t cap; const uint32_t* pos; const uint8_t* status;
  unsigned long long first, count;
  const uint32_t* indices; uint32_t n_exp;   // exported property columns (n_exp <= n_var)
  double* particle_values;                   // device chunk buffer, (n_exp + 1) rows of `count`
  double* spatial_values; uint32_t n_comp;   // (n_exp + 1) x n_comp, accumulated over chunks
};
template <class M> __device__ __forceinline__ void export_body(const ExportParams& p) {
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < p.count;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = p.first + j;
    const bool idle = p.status[i] == (uint8_t)Idle;
    float v[M::n_var];
#pragma unroll
    for (int k = 0; k < M::n_var; ++k) v[k] = p.props[(size_t)k * p.cap + i];
    const uint32_t c = p.pos[i];
    for (uint32_t e = 0; e < p.n_exp; ++e) {
      const uint32_t k = p.indices ? p.indices[e] : e;
      float cur = 0.f;
#pragma unroll
      for (int q = 0; q < M::n_var; ++q) if ((uint32_t)q == k) cur = v[q];
      p.particle_values[(size_t)e * p.count + j] = idle ? (double)cur : 0.0;
      if (idle) atomicAdd(p.spatial_values + (size_t)e * p.n_comp + c, (double)cur);
    }
    const double m = M::mass((size_t)i, RegRow{v});
    p.particle_values[(size_t)p.n_exp * p.count + j] = idle ? m : 0.0;
    if (idle) atomicAdd(p.spatial_values + (size_t)p.n_exp * p.n_comp + c, m);
  }
}

// __global__ entry points of the built-in models (the NVRTC path of user models wraps the same
// bodies in extern "C" kernels, see bmc_udf.cu)
template <class M> __global__ void __launch_bounds__(256) pre_step_kernel(const __grid_constant__ PreParams p) { pre_step_body<M>(p); }
// WB = 256-thread units per block (one block per SM): 4 -> 1024 threads x <=64 registers, 3 -> 768 x <=80,
// 2 -> 512 x <=128.  One big block per SM shares one set of shared-memory source bins among all its warps.
template <class M, int VEC, int WB, bool LAZY>
__global__ void __launch_bounds__(kBlock * WB, 1) cycle_kernel(const __grid_constant__ CycleParams p) {
  cycle_body<M, VEC, LAZY, kBlock * WB>(p);
}
template <class M> __global__ void __launch_bounds__(256) init_kernel(const __grid_constant__ InitParams p) { init_body<M>(p); }
template <class M> __global__ void __launch_bounds__(256) export_kernel(const __grid_constant__ ExportParams p) { export_body<M>(p); }

}  // namespace bmc
