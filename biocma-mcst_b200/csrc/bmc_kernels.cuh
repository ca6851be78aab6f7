// Hand-written sm_100a kernels of the Monte-Carlo particle step.
//
// One time step of SimulationUnit::cycleProcess
// (apps/libs/simulation/public/simulation/simulation.hpp:183-239) is, on the
// reference, four full passes over the particle arrays (cycle_model,
// cycle_model_contribs, cycle_move, cycle_move_leave: kernels.hpp:123-224) plus a
// host synchronisation.  Here it is ONE streaming pass (`cycle_kernel`) over
// structure-of-arrays state followed by O(events) bookkeeping kernels, with no
// host synchronisation:
//
//   pre_step     zero the source accumulators, build the per-compartment table,
//                clear last step's division bits, fix this step's buffer capacity
//   cycle        fused model update + division + contribution scatter + move +
//                outlet exit                      [HBM-bound, dominant kernel];
//                its last block decides update_and_remove_inactive (the "plan")
//   compact_*    deterministic stream compaction of exited particles (launched only
//                when inactive particles can exist; early-exit when not triggered)
//   post         merge_buffer: append newborns in ascending-mother order, commit
//
// Work distribution: a persistent grid (multiple of the SM count); block b owns
// the contiguous 1024-particle tiles [b*T/G, (b+1)*T/G).  Contiguous ownership
// lets the block (i) accumulate the per-compartment source terms in shared
// memory for its whole range and flush once, and (ii) produce block-local
// prefix sums of division counts so that newborn placement is deterministic
// without a global scan.
#pragma once
#include "bmc_models.cuh"
#include "bmc_rng.cuh"

namespace bmc {

#ifndef BMC_SCATTER_MODE
#define BMC_SCATTER_MODE 1  // 1 = shared-memory fp64 bins (product); others are timing experiments
#endif
// Cache-streaming ld/st hints (ld.global.cs / st.global.cs) on the particle columns were measured to
// HURT: with 32 resident warps per SM and more than ~3e7 particles the step becomes 60-90 % slower
// (1e8 particles, monod: 1620 us with the hints, 855 us without), so plain accesses are the default.
#ifndef BMC_STREAM_HINTS
#define BMC_STREAM_HINTS 0
#endif
#if BMC_STREAM_HINTS
#define BMC_LD(p) __ldcs(p)
#define BMC_ST(p, v) __stcs(p, v)
#else
#define BMC_LD(p) (*(p))
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
  unsigned int next_group, pad2;    // work counter of the particle pass (groups drawn beyond each warp's first)
  unsigned long long dbg[16];       // BMC_TIMELINE builds: %globaltimer stamps of block 0 (tuning only)
};

struct Outlet { uint32_t index; uint32_t pad; double flow; double dt_flow; double volume; };

// -----------------------------------------------------------------------------
// Step-stamped ages.  The reference adds d_t to both ages of every idle particle on every
// step (ages(i,1) += float(d_t) model_kernel.hpp:191; ages(i,0) += d_t move_kernel.hpp:596) —
// 16 bytes of HBM traffic per particle-step for values no kernel ever reads.  While d_t and
// the outlet configuration are constant and every age started at zero, the age of a particle
// is a pure function of the number of steps since it was last reset, BITWISE: the k-fold
// floating-point accumulation A[k] = fl(A[k-1] + d_t) is the same for every particle.  The age
// columns then hold 32-bit step stamps
//     idle particle :  s            age = A[now - s]   (s = first step that ages the particle)
//     frozen        :  kFrozen | k  age = A[k]         (exited particle: no longer updated)
// which are written only when an age is reset (division, birth, exit).  A_div / A_hyd are
// extended by one entry per step on the device (post_kernel) and applied when ages are read
// (bmc_get_particles) — bit-identical to the eager accumulation.  If d_t or the outlet
// configuration changes, or the caller supplies non-zero ages, the columns are converted to
// floats in place and the eager kernel variant (LAZY = false) takes over.
// -----------------------------------------------------------------------------
constexpr uint32_t kFrozen = 0x80000000u;

#if defined(BMC_TIMELINE)
#define BMC_STAMP(st, i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (st)->dbg[i] = t_; } } while (0)
#else
#define BMC_STAMP(st, i) do { } while (0)
#endif

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
  unsigned long long min_removal; double dead_ratio;  // RuntimeParameters of update_and_remove_inactive
  int count_step;  // 1 when called from a cycle, 0 from force_remove_dead
  // step-stamped ages (bmc_kernels.cuh): stamp given to newborns (0 = eager float ages, bits of 0.f)
  // and the per-step extension of the age tables A_div / A_hyd
  uint32_t newborn_stamp;
  float* tab_div; float* tab_hyd; uint32_t tab_idx; int tab_extend; int enable_leave; float dt_f; double dt;
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
  double* sources;   // published by the last block: sources = acc, acc = 0
  // compartment table built by every block in shared memory (small n_comp) instead of pre_step
  const double* diag; const double* vol; int ctab_in_smem; uint32_t ctab_offset;
  float weight;
  double dt; float dt_f;
  uint32_t step, rank, seed_lo, seed_hi;
  int enable_move, enable_leave, bins_in_smem;
  uint32_t stage_offset;  // byte offset of the cp.async staging buffers inside dynamic shared memory
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
// Compartment table, one row per compartment (n_comp rows — 500 .. 10k — not N):
//     col 0        ceil_f32(dt * diag_transition / liquid_volume)   leave threshold
//     col 1..n_pre M::compartment_terms(c, compartment)             optional model hook
// A model whose update starts with a function of the local concentration only
// (Monod: mu = mu_max*s/(k_s+s)) hoists it here: the IEEE division then runs once
// per compartment instead of once per particle, with bit-identical results.
// Small tables are built by every cycle block in shared memory (no extra launch);
// pre_step builds large ones in global memory.
// -----------------------------------------------------------------------------
struct PreParams {
  const double* diag; const double* vol; double dt; const double* conc; uint32_t n_species; float* ctab; uint32_t n_comp;
  int enable_move;
};

template <class M> __device__ __forceinline__ void compartment_row(const double* diag, const double* vol, double dt, const double* conc,
                                                                   uint32_t n_species, int enable_move, uint32_t c, float* row) {
  row[0] = enable_move ? __double2float_ru(dt * diag[c] / vol[c]) : 0.0f;
  if constexpr (M::n_pre > 0) M::compartment_terms(ConcView{conc, n_species, nullptr}, (size_t)c, row + 1);
}

template <class M> __device__ __forceinline__ void pre_step_body(const PreParams& p) {
  constexpr int CT = 1 + M::n_pre;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  for (uint32_t c = i; c < p.n_comp; c += nthreads) {
    float row[CT];
    compartment_row<M>(p.diag, p.vol, p.dt, p.conc, p.n_species, p.enable_move, c, row);
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

// flags of the slots a thread owns in `tile` (slot = r*blockDim + thread, r < 4: blocks of 256..1024
// threads) + per-virtual-warp counts in s_w[32] (virtual warp = 32 consecutive slots); returns the
// ballots; ends with a barrier so that s_w is complete
template <bool WANT_GAP>
__device__ __forceinline__ void tile_flags(const PostParams& p, uint32_t tile, unsigned long long old_n, unsigned long long new_n,
                                           unsigned (&bal)[4], bool (&flag)[4], unsigned* s_w) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const unsigned slot = (unsigned)r * blockDim.x + threadIdx.x;  // uniform per warp: blockDim is a multiple of 32
    flag[r] = false; bal[r] = 0u;
    if (slot < (unsigned)kTile) {
      const unsigned long long i = (unsigned long long)tile * kTile + slot;
      bool f = false;
      if (i < old_n) {
        const bool is_idle = p.status[i] == (uint8_t)Idle;
        f = WANT_GAP ? (i < new_n && !is_idle) : (i >= new_n && is_idle);
      }
      flag[r] = f;
      bal[r] = __ballot_sync(0xffffffffu, f);
      if (lane == 0) s_w[slot >> 5] = __popc(bal[r]);
    }
  }
  __syncthreads();
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
  reinterpret_cast<uint32_t*>(p.age_hyd)[dst] = p.newborn_stamp;
  reinterpret_cast<uint32_t*>(p.age_div)[dst] = p.newborn_stamp;
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
  // ---- publish the source terms of this step; the accumulator is left zeroed for the next one
  // (not from force_remove_dead: the sources of the last cycle stay what they are)
  if (p.count_step) {
    for (unsigned long long k = gtid; k < p.n_bins; k += gstride) {
      p.sources[k] = __ldcg(p.acc + k);
      p.acc[k] = 0.0;
    }
  }
  BMC_STAMP(st, 5);
  // ---- plan (read-only on the counters: every block derives the same values; one warp per block
  // reads them, so that the few counter lines are not hammered by every thread of the grid)
  __shared__ unsigned long long s_plan[6];
  if (threadIdx.x < 6) {
    const unsigned long long* src = threadIdx.x == 0 ? &st->step_exit : threadIdx.x == 1 ? &st->n_used : threadIdx.x == 2 ? &st->inactive
                                  : threadIdx.x == 3 ? &st->buf_index : threadIdx.x == 4 ? &st->buf_cap_eff : nullptr;
    s_plan[threadIdx.x] = src ? __ldcg(src) : (unsigned long long)__ldcg(&st->force_compact);
  }
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
    const uint32_t n_tiles = (uint32_t)((old_n + kTile - 1) / kTile);
    const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
    const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
    // ---- count ----
    {
      unsigned run_g = 0, run_i = 0;
      for (uint32_t tile = t0; tile < t1; ++tile) {
        unsigned bal[4]; bool fl[4];
        tile_flags<true>(p, tile, old_n, new_n, bal, fl, s_w);
        unsigned tg = 0;
        if (warp == 0) tg = __reduce_add_sync(0xffffffffu, s_w[lane]);
        __syncthreads();
        tile_flags<false>(p, tile, old_n, new_n, bal, fl, s_w);
        if (warp == 0) {
          const unsigned ti = __reduce_add_sync(0xffffffffu, s_w[lane]);
          if (lane == 0) { p.tile_gap_off[tile] = run_g; p.tile_idle_off[tile] = run_i; }
          run_g += tg; run_i += ti;
        }
        __syncthreads();
      }
      if (threadIdx.x == 0) { p.blk_gap[blockIdx.x] = run_g; p.blk_idle[blockIdx.x] = run_i; }
    }
    grid_barrier(st);
    // ---- src: k-th idle tail particle counted from the end ----
    unsigned total_idle;
    {
      const unsigned blk_off = block_prefix_of(p.blk_idle, gridDim.x, blockIdx.x, s_pref, total_idle);
      const uint32_t first_tail_tile = (uint32_t)(new_n / kTile);
      for (uint32_t tile = (t0 > first_tail_tile ? t0 : first_tail_tile); tile < t1; ++tile) {
        unsigned bal[4]; bool fl[4];
        tile_flags<false>(p, tile, old_n, new_n, bal, fl, s_w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (fl[r]) {
            unsigned woff = 0;
            const unsigned slot = (unsigned)r * blockDim.x + threadIdx.x;
            for (unsigned k = 0; k < (slot >> 5); ++k) woff += s_w[k];
            const unsigned asc = blk_off + p.tile_idle_off[tile] + woff + __popc(bal[r] & ((1u << lane) - 1u));
            p.src[total_idle - 1u - asc] = (uint32_t)((unsigned long long)tile * kTile + slot);
          }
        }
        __syncthreads();
      }
    }
    grid_barrier(st);
    // ---- move: gaps below new_n pull their replacement ----
    {
      unsigned total_gap;
      const unsigned blk_off = block_prefix_of(p.blk_gap, gridDim.x, blockIdx.x, s_pref, total_gap);
      const uint32_t last_low_tile = (uint32_t)((new_n + kTile - 1) / kTile);  // exclusive
      for (uint32_t tile = t0; tile < t1 && tile < last_low_tile; ++tile) {
        unsigned bal[4]; bool fl[4];
        tile_flags<true>(p, tile, old_n, new_n, bal, fl, s_w);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (fl[r]) {
            unsigned woff = 0;
            const unsigned slot = (unsigned)r * blockDim.x + threadIdx.x;
            for (unsigned k = 0; k < (slot >> 5); ++k) woff += s_w[k];
            const unsigned k = blk_off + p.tile_gap_off[tile] + woff + __popc(bal[r] & ((1u << lane) - 1u));
            if (k >= total_idle) {
              atomicOr(&st->error, 2u);  // inactive counter inconsistent with the status column
            } else {
              const size_t i = (size_t)tile * kTile + slot;
              const size_t s2 = p.src[k];
              p.status[i] = (uint8_t)Idle;
              p.pos[i] = p.pos[s2];
              for (int c = 0; c < p.n_var; ++c) p.props[(size_t)c * p.cap + i] = p.props[(size_t)c * p.cap + s2];
              p.age_hyd[i] = p.age_hyd[s2];
              p.age_div[i] = p.age_div[s2];
            }
          }
        }
        __syncthreads();
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
    if ((unsigned long long)blockIdx.x * blockDim.x < n_add) {  // blocks that have a newborn to place
      for (unsigned j = threadIdx.x; j < (unsigned)n_add; j += blockDim.x) s_pref[j] = __ldcg(p.buf_mother + j);
      __syncthreads();
      const unsigned long long j = gtid;
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
    const unsigned long long words_valid = (old_n + 31ull) / 32ull;  // words beyond the last slot are stale
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
      // rank of the mother among the dividing mothers of its tile: all 32 mask words in one round trip
      const uint4* w4 = reinterpret_cast<const uint4*>(p.div_mask + (size_t)tile * (kTile / 32));
      const unsigned wi = (mother & (kTile - 1)) >> 5, bit = mother & 31u;
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
          const unsigned idx = (unsigned)(4 * q + c);
          const unsigned m = idx < wi ? 0xffffffffu : (idx == wi ? ((1u << bit) - 1u) : 0u);
          rank += __popc(ww[c] & m);
        }
      }
      insert_newborn(p, j, new_n + s_pref[b] + toff + rank, npos);
    }
  }
  BMC_STAMP(st, 7);
  // commit: the last block to get here (every other block is done reading the counters)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(&st->done_blocks, 1u) == gridDim.x - 1) {
    __threadfence();
    st->done_blocks = 0; st->next_group = 0;
    st->last_out = out; st->last_dead = 0; st->last_waiting = st->step_waiting;
    st->total_out += out;
    st->step_exit = 0; st->step_waiting = 0; st->buf_index = 0; st->force_compact = 0;
    st->inactive = do_compact ? 0ull : inactive;
    if (do_compact) st->n_compactions += 1;
    st->do_compact = do_compact ? 1u : 0u; st->cmp_old_n = old_n; st->cmp_new_n = new_n; st->n_add = n_add;  // for inspection
    const unsigned long long n = new_n + n_add;
    st->n_used = n;
    st->total_new += n_add;
    st->step += (unsigned long long)p.count_step;
    // room of the next step's division buffer: min(B, capacity - n_used); the device can never write
    // past the capacity, growth is done lazily by the host
    const unsigned long long room = p.cap > n ? p.cap - n : 0ull;
    st->buf_cap_eff = p.buf_cap < room ? p.buf_cap : room;
    if (p.tab_extend) {  // A[k+1] = fl(A[k] + d_t): exactly the accumulation an eagerly updated age goes through
      p.tab_div[p.tab_idx + 1] = p.tab_div[p.tab_idx] + p.dt_f;                       // model_kernel.hpp:191 (float d_t)
      p.tab_hyd[p.tab_idx + 1] = p.enable_leave ? (float)((double)p.tab_hyd[p.tab_idx] + p.dt)  // move_kernel.hpp:596 (double d_t)
                                                : p.tab_hyd[p.tab_idx];
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
// Structure of the per-thread body (VEC particles per thread, 128-bit column
// accesses): the common path — load, model update, age updates, leave/outlet
// tests — is straight-line code over the VEC particles so that their dependency
// chains interleave; everything rare (division, the neighbour pick of a mover,
// the outlet exit draw, partially idle groups) sits behind warp-level votes.
// -----------------------------------------------------------------------------
__host__ __device__ constexpr int popcount_c(uint64_t x) { return x == 0u ? 0 : (int)(x & 1u) + popcount_c(x >> 1); }
__host__ __device__ constexpr bool col_flag(uint64_t mask, int k) { return ((mask >> k) & 1ull) != 0ull; }  // k < 64
// columns actually loaded per slot: every property that is not write-only
template <class M> struct ReadCols {
  static_assert(M::n_var <= 64, "at most 64 properties per particle");
  static constexpr uint64_t all = M::n_var >= 64 ? ~0ull : ((1ull << M::n_var) - 1ull);
  static constexpr int value = M::n_var - popcount_c((uint64_t)M::write_only_mask & all);
};
// One staging buffer of the bulk-copy pipeline holds one GROUP = 32*VEC consecutive slots (the work
// unit of a warp): pos and the read columns (32*VEC*4 bytes each), then the status bytes.  Every warp
// has kStages private buffers.  (The pipeline is built for step-stamped ages only: no age column.)
template <class M, int VEC> struct StageBytes {
  static constexpr size_t col = (size_t)32 * 4 * VEC;
  static constexpr size_t warp_stage = (size_t)(1 + ReadCols<M>::value) * col + (size_t)32 * VEC;
};
constexpr int kStages = 2;

// ---- mbarrier + TMA bulk copy (cp.async.bulk, 1-D; SASS: UBLKCP) --------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

template <class M, int VEC, bool PIPE, bool LAZY, int BLOCK> __device__ __forceinline__ void cycle_body(const CycleParams& p) {
  constexpr int kWarps = BLOCK / 32;
  static_assert(!PIPE || LAZY, "the bulk-copy pipeline is built for step-stamped ages only");
  constexpr int NV = M::n_var, NC = M::n_c, CT = 1 + M::n_pre;
  constexpr uint32_t kGroup = 32 * VEC;         // slots per group: the work unit of one warp
  constexpr int kColStride = 32 * 4 * VEC;      // bytes between staged columns of a warp's buffer
  constexpr size_t kWarpStage = StageBytes<M, VEC>::warp_stage;
  extern __shared__ __align__(128) double s_bins[];  // [n_species * n_comp] when bins_in_smem, then kStages staging buffers
  __shared__ unsigned long long s_cnt[4];      // move, exit, new, overflow

  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long n_used = p.st->n_used;
  const unsigned long long buf_cap = p.st->buf_cap_eff;
  // Work distribution: the slots are cut into groups of 32*VEC; every warp of the (persistent,
  // fully resident) grid starts with the group of its own index and then draws further groups from
  // a device-wide counter — one L2 atomic per group, issued one group ahead so that its latency is
  // hidden.  Blocks that start late or run on a slower SM simply process fewer groups.
  const uint32_t n_groups = (uint32_t)((n_used + kGroup - 1) / kGroup);
  const uint32_t total_warps = gridDim.x * kWarps;
  const uint32_t n_bins = p.n_species * p.n_comp;
  const bool single_comp = (p.n_comp == 1);
  const bool smem_bins = p.bins_in_smem && !single_comp;

  BMC_STAMP(p.st, 0);
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
  if (smem_bins)
    for (uint32_t k = threadIdx.x; k < n_bins; k += BLOCK) s_bins[k] = 0.0;
  float* const s_ctab = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(s_bins) + p.ctab_offset);
  if (p.ctab_in_smem) {
    for (uint32_t c = threadIdx.x; c < p.n_comp; c += BLOCK) {
      float row[CT];
      compartment_row<M>(p.diag, p.vol, p.dt, p.conc, p.n_species, p.enable_move, c, row);
#pragma unroll
      for (int k = 0; k < CT; ++k) s_ctab[c * CT + k] = row[k];
    }
  }
  __syncthreads();

  BMC_STAMP(p.st, 1);
  unsigned c_move = 0, c_exit = 0, c_new = 0, c_over = 0;
  double acc0d[NC];  // single-compartment accumulation lives in registers
#pragma unroll
  for (int j = 0; j < NC; ++j) acc0d[j] = 0.0;
  const double w = (double)p.weight;  // `const double weight = get_weight(p)` contribution_kernel.hpp:179
  const BufRows bufrows{p.buf_props, p.buf_stride};
  // LAZY: the age columns hold step stamps instead of floats (see "Step-stamped ages" below)
  uint32_t* const stamps_hyd = reinterpret_cast<uint32_t*>(p.age_hyd);
  uint32_t* const stamps_div = reinterpret_cast<uint32_t*>(p.age_div);
  (void)stamps_hyd; (void)stamps_div;
  const uint32_t outlet0 = p.n_flows > 0 ? p.outlets[0].index : 0xffffffffu;
  const bool outlet0_live = p.n_flows > 0 && p.outlets[0].flow != 0.;

  // PIPE: two-stage bulk-copy pipeline, private to each warp.  Lane 0 arms the stage's mbarrier with
  // the byte count and issues one cp.async.bulk (TMA, 1-D) per column for the warp's NEXT group; the
  // bytes land in shared memory while the warp computes its current group, so the HBM latency is off
  // the critical path and no registers are held by loads in flight.  Lane l then reads bytes
  // [l*4*VEC, (l+1)*4*VEC) of every staged column (conflict-free LDS.128).  No block-wide barrier.
  unsigned char* const s_stage = reinterpret_cast<unsigned char*>(s_bins) + p.stage_offset;
  __shared__ __align__(8) unsigned long long s_bar[kStages][kWarps];
  constexpr unsigned kColBytes = (unsigned)kColStride;
  auto issue = [&](uint32_t g, int buf) {  // executed by lane 0 of the warp
    const size_t g0 = (size_t)g * kGroup;  // first slot of the group (the whole group is below the capacity)
    unsigned char* dst = s_stage + ((size_t)buf * kWarps + warp) * kWarpStage;
    unsigned long long* bar = &s_bar[buf][warp];
    mbar_expect_tx(bar, (unsigned)kWarpStage);
    bulk_g2s(dst, p.pos + g0, kColBytes, bar);
    int c = 1;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (!col_flag(M::write_only_mask, k)) { bulk_g2s(dst + c * kColStride, p.props + (size_t)k * p.cap + g0, kColBytes, bar); ++c; }
    bulk_g2s(dst + (size_t)(1 + ReadCols<M>::value) * kColStride, p.status + g0, (unsigned)kGroup, bar);
  };

  {
    // Every group but the last is entirely below n_used: the body is instantiated twice so that the
    // common case carries no per-slot range checks (FULL), the ragged tail keeps them.
    auto body = [&](auto full_tag, const uint32_t g, const int buf) {
      constexpr bool FULL = decltype(full_tag)::value;
      const uint32_t i_raw = g * kGroup + lane * VEC;
      const bool live = FULL || i_raw < n_used;    // false only in the ragged end of the last group
      const size_t i0 = live ? i_raw : 0u;         // dead lanes shadow slot 0 (loads stay in range, nothing is stored)

      uint32_t pos[VEC]; float adiv[VEC], ahyd[VEC]; float v[VEC][NV], old[VEC][NV];
      uint32_t stw;
      if constexpr (PIPE) {
        // ---- operands were staged in shared memory by the bulk copies issued one iteration ago ----
        const unsigned char* stage = s_stage + ((size_t)buf * kWarps + warp) * kWarpStage;
        const unsigned char* src = stage + lane * (4 * VEC);
        stw = VecIO<VEC>::ldb_plain(stage + (size_t)(1 + ReadCols<M>::value) * kColStride + lane * VEC);
        VecIO<VEC>::ldu_plain(reinterpret_cast<const uint32_t*>(src), pos);
        int c = 1;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          float col[VEC];
          if (col_flag(M::write_only_mask, k)) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) col[q] = 0.f;
          } else {
            VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src + c * kColStride), col);
            ++c;
          }
#pragma unroll
          for (int q = 0; q < VEC; ++q) { v[q][k] = col[q]; old[q][k] = col[q]; }
        }
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

      // ---- compartment rows: leave threshold + model terms, one gather per particle
      float ctab[VEC][CT];
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        if (p.ctab_in_smem) {
          const float* row = s_ctab + pos[q] * CT;
          if constexpr (CT == 2) { const float2 t = *reinterpret_cast<const float2*>(row); ctab[q][0] = t.x; ctab[q][1] = t.y; }
          else {
#pragma unroll
            for (int k = 0; k < CT; ++k) ctab[q][k] = row[k];
          }
        } else {
          const float* row = p.ctab + (size_t)pos[q] * CT;
          if constexpr (CT == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(row)); ctab[q][0] = t.x; ctab[q][1] = t.y; }
          else if constexpr (CT == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(row)); ctab[q][0] = t.x; ctab[q][1] = t.y; ctab[q][2] = t.z; ctab[q][3] = t.w; }
          else {
#pragma unroll
            for (int k = 0; k < CT; ++k) ctab[q][k] = __ldg(row + k);
          }
        }
      }

      // ---- u1: ONE Philox block per group of four slots (draw_block 0) ---------
      uint32_t rw1[4] = {0u, 0u, 0u, 0u};
      if (p.enable_move) philox4x32_10((uint32_t)(i0 >> 2), p.step, 0u, p.rank, p.seed_lo, p.seed_hi, rw1);

      // ---- model update: unconditional straight-line code over the VEC particles;
      // results of non-idle slots are never stored (model_kernel.hpp:186-196)
      float contrib[VEC][NC];
      unsigned div_nib = 0;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        if constexpr (!LAZY) adiv[q] = idle[q] ? adiv[q] + p.dt_f : adiv[q];  // ages(i,1) += _d_t  (model_kernel.hpp:191)
        Gen gen(p.seed_lo, p.seed_hi, p.rank, (uint32_t)(i0 + q), p.step, 2u);
        const ConcView conc{p.conc, p.n_species, &ctab[q][1]};
        const Status s = M::update(gen, p.dt_f, i0 + q, RegRow{v[q]}, RegRow{contrib[q]}, (size_t)pos[q], conc);
        div_nib |= (unsigned)(idle[q] && s == Division) << q;
      }

      // ---- contribution scatter at the PRE-move position (Q15) -----------------
      if (single_comp) {
#pragma unroll
        for (int q = 0; q < VEC; ++q)
#pragma unroll
          for (int j = 0; j < NC; ++j) acc0d[j] += idle[q] ? w * (double)contrib[q][j] : 0.0;
      } else if (smem_bins) {  // block-private fp64 bins (LDS/DADD/ATOMS.CAST.SPIN), flushed once per block
#pragma unroll
        for (int q = 0; q < VEC; ++q)
          if (idle[q]) {
#pragma unroll
            for (int j = 0; j < NC; ++j) atomicAdd(&s_bins[(uint32_t)j + p.n_species * pos[q]], w * (double)contrib[q][j]);
          }
      } else {  // table too large for shared memory: L2 atomics (RED.F64)
#pragma unroll
        for (int q = 0; q < VEC; ++q)
          if (idle[q]) {
#pragma unroll
            for (int j = 0; j < NC; ++j) atomicAdd(p.acc + (size_t)j + (size_t)p.n_species * pos[q], w * (double)contrib[q][j]);
          }
      }

      // ---- division: handle_division (particles_container.hpp:559-573) -------
      // warp-aggregated slot allocation: ONE atomic per warp that has a dividing
      // mother (reference: one per mother, a6).  Rows are allocated in ascending
      // particle order inside the warp; final newborn placement is re-ranked by
      // mother index in insert_kernel, so the result does not depend on the
      // order warps hit the atomic.
      unsigned ok_nib = 0;
      if (__any_sync(0xffffffffu, div_nib != 0u)) {
        const unsigned cnt = __popc(div_nib);
        unsigned total;
        const unsigned excl = warp_excl_scan(cnt, total);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&p.st->buf_index, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        unsigned r = 0;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if ((div_nib >> q) & 1u) {
            const unsigned long long j = base + excl + r;
            ++r;
            ++c_new;  // NewParticle++ even on overflow (model_kernel.hpp:259, Q6)
            if (j < buf_cap) {
              Gen gen(p.seed_lo, p.seed_hi, p.rank, (uint32_t)(i0 + q), p.step, 0x40000000u);
              M::division(gen, i0 + q, (size_t)j, RegRow{v[q]}, bufrows);
              p.buf_pos[j] = pos[q];               // buffer_position(idx2) = position(idx1): pre-move
              p.buf_mother[j] = (uint32_t)(i0 + q);
              if constexpr (LAZY) stamps_div[i0 + q] = p.step + 1u;  // ages(idx1,1) = 0: counts from the next step
              else adiv[q] = 0.f;                                    // ages(idx1,1) = 0
              ok_nib |= 1u << q;
            } else {
              ++c_over;  // waiting_allocation_particle / Overflow (model_kernel.hpp:253-258)
            }
          }
        }
      }
      {
        // division bitmask: bit (slot & 31) of word (slot >> 5); a word is owned by 32/VEC lanes.
        // Written for EVERY group of every step (0.125 B/slot), so no bit ever needs clearing.
        constexpr int LPW = 32 / VEC;
        unsigned word = ok_nib << (VEC * (lane % LPW));
#pragma unroll
        for (int o = 1; o < LPW; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
#if !defined(BMC_EXP_NO_MASKSTORE)
        if ((lane % LPW) == 0) p.div_mask[i_raw >> 5] = word;  // i_raw: also the ragged end of the last tile (zeros)
#else
        if ((lane % LPW) == 0 && word) p.div_mask[i_raw >> 5] = word;
#endif
      }

      // ---- move (all slots, no status check: move_kernel.hpp:392-437) --------
      unsigned moved = 0;
      if (p.enable_move) {
        unsigned mv = 0;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const float u1 = u01f(pick4(rw1, (unsigned)((i0 + q) & 3)));
          mv |= (unsigned)(u1 < ctab[q][0]) << q;  // (dt*flow/volume) > rng1
        }
        mv &= valid_m;
        moved = mv;
        // movers are rare (dt*F/V ~ 1e-2): their neighbour pick draws its own block
        while (mv) {
          const int q = __ffs(mv) - 1;
          mv &= mv - 1;
          uint32_t c = pos[0];
#pragma unroll
          for (int qq = 1; qq < VEC; ++qq) if (qq == q) c = pos[qq];
          uint32_t r2[4];
          philox4x32_10((uint32_t)(i0 + q), p.step, 2u, p.rank, p.seed_lo, p.seed_hi, r2);
          const float u2 = u01f(r2[0]);
          const float* row = p.cdf + (size_t)c * p.m;
          int left = 0, right = p.m - 1;
          while (left < right) {  // __find_next_compartment, move_kernel.hpp:87-95
            const int mid = (left + right) >> 1;
            if (u2 > __ldg(row + mid)) left = mid + 1; else right = mid;
          }
          const uint32_t np = __ldg(p.neigh + (size_t)c * p.m + left);
#pragma unroll
          for (int qq = 0; qq < VEC; ++qq) if (qq == q) pos[qq] = np;
          ++c_move;  // events.wrap_incr<Move>() (Q20: aggregated)
        }
      }

      // ---- leave (Idle only, post-move position: move_kernel.hpp:347-359) ----
      unsigned exit_nib = 0;
      if (p.enable_leave) {
        unsigned in_outlet = 0; int fsel[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          if constexpr (!LAZY) ahyd[q] = idle[q] ? (float)((double)ahyd[q] + p.dt) : ahyd[q];  // ages(idx,0) += d_t (double)
          fsel[q] = 0;
        }
        if (p.n_flows == 1) {  // the usual case (0D reactor or a single outlet, move_kernel.hpp:113)
#pragma unroll
          for (int q = 0; q < VEC; ++q) in_outlet |= (unsigned)(outlet0_live && pos[q] == outlet0) << q;
        } else {
#pragma unroll
          for (int q = 0; q < VEC; ++q)
            for (int f = 0; f < p.n_flows; ++f)  // find_flow: first match wins
              if (p.outlets[f].index == pos[q]) { if (p.outlets[f].flow != 0.) { in_outlet |= 1u << q; fsel[q] = f; } break; }
        }
        in_outlet &= idle_m;
        if (in_outlet) {  // u3: draw_block 1 of the group of four
          uint32_t rw3[4];
          philox4x32_10((uint32_t)(i0 >> 2), p.step, 1u, p.rank, p.seed_lo, p.seed_hi, rw3);
#pragma unroll
          for (int q = 0; q < VEC; ++q) {
            if ((in_outlet >> q) & 1u) {
              const float u3 = u01f(pick4(rw3, (unsigned)((i0 + q) & 3)));
              const float lnu = (float)log((double)u3);  // Kokkos::log(float), see oracle ln_f32
              const Outlet& o = p.outlets[fsel[q]];
              if (o.dt_flow > (double)(-lnu) * o.volume) {  // probability_leaving<precision_tag>
                if constexpr (LAZY) {
                  // the particle stops ageing: freeze the step counts (age_hyd = 0, age_div as of this step)
                  stamps_hyd[i0 + q] = kFrozen;
                  const uint32_t sd = stamps_div[i0 + q];
                  stamps_div[i0 + q] = (sd & kFrozen) ? sd : (kFrozen | (p.step + 1u - sd));
                } else {
                  ahyd[q] = ahyd[q] * 0.0f;                 // ages(idx,0) *= (1 - leave_mask)
                }
                exit_nib |= 1u << q;
                ++c_exit;
              }
            }
          }
        }
      }

      // ---- write back ---------------------------------------------------------
      // Columns the model assigns on every update (always_written_mask, which includes the
      // write-only ones) are stored whenever the thread has an idle slot; the others only if a
      // value changed bitwise (the comparison folds away for columns the hooks never assign).
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
        if (idle_m) {
          VecIO<VEC>::stf(p.age_div + i0, adiv);
          if (p.enable_leave) VecIO<VEC>::stf(p.age_hyd + i0, ahyd);
        }
      }
      if (moved) {  // Q14: position written only for movers (never for slots >= n_used: mv is masked)
#pragma unroll
        for (int q = 0; q < VEC; ++q) if ((moved >> q) & 1u) p.pos[i0 + q] = pos[q];
      }
      if (exit_nib) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) if ((exit_nib >> q) & 1u) p.status[i0 + q] = (uint8_t)Exit;
      }
    };
    // Draw number s -> group s: the warps that run at the same time work on adjacent groups, i.e. the
    // whole grid streams through ONE moving window of every column.  (Spreading the draws over many
    // interleaved address streams was measured to make no difference.)
    const unsigned long long n_draws = n_groups;
    auto group_of = [&](unsigned long long s) -> uint32_t { return (uint32_t)s; };
    unsigned long long s = (unsigned long long)blockIdx.x * kWarps + warp;  // first draw: the warp's own index
    if constexpr (PIPE) {
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < kStages; ++b) mbar_init(&s_bar[b][warp], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncwarp();
      uint32_t nxt = 0;
      uint32_t g = s < n_draws ? group_of(s) : n_groups;
      if (lane == 0) {
        if (g < n_groups) issue(g, 0);
        nxt = atomicAdd(&p.st->next_group, 1u);  // the pipeline needs the next group one iteration ahead
      }
      uint32_t it = 0;  // counts the groups that were staged (buffer / barrier phase)
#pragma unroll 1
      while (s < n_draws) {
        const int buf = (int)(it & 1u);
        const unsigned long long s_next = (unsigned long long)total_warps + __shfl_sync(0xffffffffu, nxt, 0);
        const uint32_t g_next = s_next < n_draws ? group_of(s_next) : n_groups;
        const bool have = g < n_groups, have_next = g_next < n_groups;
        __syncwarp();  // every lane is done reading the other buffer (previous iteration): it may be refilled
        if (lane == 0) {
          if (have_next) issue(g_next, have ? buf ^ 1 : buf);
          nxt = atomicAdd(&p.st->next_group, 1u);
        }
        if (have) {
          mbar_wait(&s_bar[buf][warp], (it >> 1) & 1u);
          if ((unsigned long long)(g + 1) * kGroup <= n_used) body(FullTile{}, g, buf);
          else body(RaggedTile{}, g, buf);
          ++it;
        }
        g = g_next; s = s_next;
      }
    } else {
#pragma unroll 1
      while (s < n_draws) {
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(&p.st->next_group, 1u);  // consumed after this group: latency hidden
        const uint32_t g = group_of(s);
        if (g < n_groups) {
          if ((unsigned long long)(g + 1) * kGroup <= n_used) body(FullTile{}, g, 0);
          else body(RaggedTile{}, g, 0);
        }
        s = (unsigned long long)total_warps + __shfl_sync(0xffffffffu, nxt, 0);
      }
    }
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
  // ---- block epilogue: counters, source flush, block-local tile prefix -------
  const unsigned cm = __reduce_add_sync(0xffffffffu, c_move), ce = __reduce_add_sync(0xffffffffu, c_exit);
  const unsigned cn = __reduce_add_sync(0xffffffffu, c_new), co = __reduce_add_sync(0xffffffffu, c_over);
  if (lane == 0) {
    if (cm) atomicAdd(&s_cnt[0], (unsigned long long)cm);
    if (ce) atomicAdd(&s_cnt[1], (unsigned long long)ce);
    if (cn) atomicAdd(&s_cnt[2], (unsigned long long)cn);
    if (co) atomicAdd(&s_cnt[3], (unsigned long long)co);
  }
  if (single_comp) {
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      double a = acc0d[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && a != 0.0) atomicAdd(p.acc + j, a);
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_cnt[0]) atomicAdd(&p.st->events[2], s_cnt[0]);                                               // Move
    if (s_cnt[1]) { atomicAdd(&p.st->events[1], s_cnt[1]); atomicAdd(&p.st->step_exit, s_cnt[1]); }     // Exit
    if (s_cnt[2]) atomicAdd(&p.st->events[0], s_cnt[2]);                                               // NewParticle
    if (s_cnt[3]) { atomicAdd(&p.st->events[4], s_cnt[3]); atomicAdd(&p.st->step_waiting, s_cnt[3]); }  // Overflow
  }
  if (smem_bins) {
    // every block starts at a different bin, so that the blocks (which finish together) do not hit
    // the same L2 addresses at the same time
    const uint32_t rot = (uint32_t)(((unsigned long long)blockIdx.x * n_bins) / gridDim.x);
    for (uint32_t k0 = threadIdx.x; k0 < n_bins; k0 += BLOCK) {
      uint32_t k = k0 + rot; if (k >= n_bins) k -= n_bins;
      const double a = s_bins[k];
      if (a != 0.0) atomicAdd(p.acc + k, a);
    }
  }
  // ---- second phase of the step: every block's state, buffer rows and counters are complete
  BMC_STAMP(p.st, 3);
  if (p.fuse_post) {
    grid_barrier(p.st);
    BMC_STAMP(p.st, 4);
    post_cycle_body(p.post);
    BMC_STAMP(p.st, 8);
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
template <class M, int VEC, int WB, bool PIPE, bool LAZY>
__global__ void __launch_bounds__(kBlock * WB, 1) cycle_kernel(const __grid_constant__ CycleParams p) {
  cycle_body<M, VEC, PIPE, LAZY, kBlock * WB>(p);
}
template <class M> __global__ void __launch_bounds__(256) init_kernel(const __grid_constant__ InitParams p) { init_body<M>(p); }
template <class M> __global__ void __launch_bounds__(256) export_kernel(const __grid_constant__ ExportParams p) { export_body<M>(p); }

}  // namespace bmc
