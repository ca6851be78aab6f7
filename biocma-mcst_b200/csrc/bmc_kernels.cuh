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
#include <cstdint>
#include <cuda_runtime.h>
#include "bmc_models.cuh"
#include "bmc_rng.cuh"

namespace bmc {

#ifndef BMC_SCATTER_MODE
#define BMC_SCATTER_MODE 1  // 1 = shared-memory fp64 bins (product); others are timing experiments
#endif
#ifndef BMC_STREAM_HINTS
#define BMC_STREAM_HINTS 1
#endif
#if BMC_STREAM_HINTS
#define BMC_LD(p) __ldcs(p)
#define BMC_ST(p, v) __stcs(p, v)
#else
#define BMC_LD(p) (*(p))
#define BMC_ST(p, v) (*(p) = (v))
#endif

constexpr int kTile = 1024;         // particles per rank-tile (bitmask / prefix granularity)
constexpr int kBlock = 256;         // threads per block
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
  unsigned long long clear_n;       // division records whose bitmask bits the next pre_step clears
};

struct Outlet { uint32_t index; uint32_t pad; double flow; double dt_flow; double volume; };

struct CycleParams {
  // particle SoA columns (ParticlesContainer views, particles_container.hpp:82-88)
  float* props; size_t cap;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  // division buffer (particles_container.hpp:222-227)
  float* buf_props; size_t buf_stride; uint32_t* buf_pos; uint32_t* buf_mother;
  uint32_t* div_mask;   // 1 bit / slot: mother divided this step (allocated a buffer row)
  uint32_t* tile_div;   // per tile: number of such mothers
  uint32_t* tile_off;   // per tile: exclusive prefix inside the owning block
  uint32_t* blk_total;  // per block: divisions in its range
  // domain (DomainState, domain.hpp:28-35) in derived single-precision form
  const float* ctab;       // compartment table rows {leave threshold, model terms...}
  const float* cdf;        // floor_f32(cumulative_probability), row-major n_comp x m
  const uint32_t* neigh;   // neighbors, row-major n_comp x m
  int m; uint32_t n_comp;
  int n_flows; Outlet outlets[kMaxFlows];
  // liquid coupling
  const double* conc; uint32_t n_species; double* sources;
  float weight;
  double dt; float dt_f;
  uint32_t step, rank, seed_lo, seed_hi;
  int enable_move, enable_leave, bins_in_smem;
  int prefetch_ahead;  // tiles of L2 prefetch distance (0 = off)
  unsigned long long min_removal; double dead_ratio;  // RuntimeParameters used by the post-cycle plan
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

// cp.async.bulk.prefetch.L2: one thread asks the memory system to pull a whole
// column chunk of the NEXT tile into L2 while the current tile is being computed
// (SASS: UBLKPF).  Address and size must be multiples of 16 B.
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
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
};
template <> struct VecIO<1> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[1]) { v[0] = BMC_LD(p); }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[1]) { BMC_ST(p, v[0]); }
  static __device__ __forceinline__ void ldu(const uint32_t* p, uint32_t (&v)[1]) { v[0] = BMC_LD(p); }
  static __device__ __forceinline__ uint32_t ldb(const uint8_t* p) { return BMC_LD(p); }
};

// -----------------------------------------------------------------------------
// pre_step: everything that must happen before the particle pass, in one launch:
//   * contribs_scatter.reset() (simulation.hpp:201): zero the source accumulators
//   * compartment table, one row per compartment (n_comp rows — 500 .. 10k — not N):
//       col 0        ceil_f32(dt * diag_transition / liquid_volume)   leave threshold
//       col 1..n_pre M::compartment_terms(c, compartment)             optional model hook
//     A model whose update starts with a function of the local concentration only
//     (Monod: mu = mu_max*s/(k_s+s)) hoists it here: the IEEE division then runs once
//     per compartment instead of once per particle, with bit-identical results.
//   * clear the division bitmask bits of the previous step's newborn records
//   * this step's usable buffer capacity and the per-step counters
// -----------------------------------------------------------------------------
struct PreParams {
  DevState* st; double* sources; uint32_t n_bins;
  unsigned long long cap, buf_cap; unsigned int grid_cycle;
  const double* diag; const double* vol; double dt; const double* conc; uint32_t n_species; float* ctab; uint32_t n_comp;
  int enable_move;
  const uint32_t* buf_mother; uint32_t* div_mask; uint32_t* tile_div;
};

template <class M>
__global__ void __launch_bounds__(256) pre_step_kernel(const __grid_constant__ PreParams p) {
  constexpr int CT = 1 + M::n_pre;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  for (uint32_t k = i; k < p.n_bins; k += nthreads) p.sources[k] = 0.0;
  for (uint32_t c = i; c < p.n_comp; c += nthreads) {
    float row[CT];
    row[0] = p.enable_move ? __double2float_ru(p.dt * p.diag[c] / p.vol[c]) : 0.0f;
    if constexpr (M::n_pre > 0) M::compartment_terms(ConcView{p.conc, p.n_species, nullptr}, (size_t)c, row + 1);
#pragma unroll
    for (int k = 0; k < CT; ++k) p.ctab[(size_t)c * CT + k] = row[k];
  }
  const unsigned long long n_clear = p.st->clear_n;
  for (unsigned long long j = i; j < n_clear; j += nthreads) {
    const uint32_t mother = p.buf_mother[j];
    p.div_mask[mother >> 5] = 0u;
    p.tile_div[mother >> 10] = 0u;
  }
  if (i == 0) {
    DevState* st = p.st;
    const unsigned long long n = st->n_used;
    const unsigned long long room = p.cap > n ? p.cap - n : 0ull;
    st->buf_cap_eff = p.buf_cap < room ? p.buf_cap : room;
    st->buf_index = 0; st->step_exit = 0; st->step_waiting = 0;
    st->cyc_n_used = n;
    st->cyc_tiles = (unsigned int)((n + kTile - 1) / kTile);
    st->cyc_grid = p.grid_cycle;
  }
}

// update_and_remove_inactive (particles_container.hpp:539-557) and the merge_buffer size
// (:575-581), decided on the device by ONE thread once every particle has been processed.
__device__ __forceinline__ void make_plan(DevState* st, unsigned long long min_removal, double dead_ratio) {
  const unsigned long long out = st->step_exit;
  st->last_out = out; st->last_dead = 0; st->last_waiting = st->step_waiting;
  st->total_out += out;
  st->inactive += out;  // inactive_counter += out; += dead (always 0, Q3)
  st->step_exit = 0;
  const unsigned long long n = st->n_used;
  unsigned long long thr = (unsigned long long)((double)n * dead_ratio);
  if (min_removal > thr) thr = min_removal;
  const bool trig = (st->inactive > thr) || (st->force_compact && st->inactive > 0);
  st->force_compact = 0;
  st->cmp_old_n = n;
  if (trig) {
    st->do_compact = 1;
    st->cmp_new_n = n - st->inactive;
    st->cmp_tiles = (unsigned int)((n + kTile - 1) / kTile);
  } else {
    st->do_compact = 0; st->cmp_new_n = n;
  }
  const unsigned long long bi = st->buf_index;
  st->n_add = bi < st->buf_cap_eff ? bi : st->buf_cap_eff;
  st->buf_index = 0;
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
template <class M, int VEC, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) cycle_kernel(const __grid_constant__ CycleParams p) {
  constexpr int NV = M::n_var, NC = M::n_c, NP = M::n_pre, CT = 1 + M::n_pre;
  constexpr int SUB = kTile / (kBlock * VEC);  // sub-iterations per tile
  extern __shared__ double s_bins[];           // [n_species * n_comp] when bins_in_smem
  __shared__ unsigned long long s_cnt[4];      // move, exit, new, overflow

  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long n_used = p.st->n_used;
  const unsigned long long buf_cap = p.st->buf_cap_eff;
  const uint32_t n_tiles = (uint32_t)((n_used + kTile - 1) / kTile);
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  const uint32_t n_bins = p.n_species * p.n_comp;
  const bool single_comp = (p.n_comp == 1);
  const bool smem_bins = p.bins_in_smem && !single_comp;

  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0ull;
  if (smem_bins)
    for (uint32_t k = threadIdx.x; k < n_bins; k += kBlock) s_bins[k] = 0.0;
  __syncthreads();

  unsigned c_move = 0, c_exit = 0, c_new = 0, c_over = 0;
  double acc0d[NC];  // single-compartment accumulation lives in registers
#pragma unroll
  for (int j = 0; j < NC; ++j) acc0d[j] = 0.0;
  const double w = (double)p.weight;  // `const double weight = get_weight(p)` contribution_kernel.hpp:179
  const BufRows bufrows{p.buf_props, p.buf_stride};
  const uint32_t outlet0 = p.n_flows > 0 ? p.outlets[0].index : 0xffffffffu;
  const bool outlet0_live = p.n_flows > 0 && p.outlets[0].flow != 0.;

  for (uint32_t tile = t0; tile < t1; ++tile) {
    if (p.prefetch_ahead && tile + p.prefetch_ahead < t1) {  // optional L2 bulk prefetch of a later tile
      const size_t nb = (size_t)(tile + p.prefetch_ahead) * kTile;
      const int col = (int)threadIdx.x;
      if (col < 4 + NV) {
        if (col == 0) l2_prefetch_bulk(p.status + nb, kTile);
        else if (col == 1) l2_prefetch_bulk(p.pos + nb, kTile * 4);
        else if (col == 2) l2_prefetch_bulk(p.age_div + nb, kTile * 4);
        else if (col == 3) { if (p.enable_leave) l2_prefetch_bulk(p.age_hyd + nb, kTile * 4); }
        else if (!((M::write_only_mask >> (col - 4)) & 1u)) l2_prefetch_bulk(p.props + (size_t)(col - 4) * p.cap + nb, kTile * 4);
      }
    }
#pragma unroll 1
    for (int sub = 0; sub < SUB; ++sub) {
      const size_t i_raw = (size_t)tile * kTile + ((size_t)sub * (kBlock / 32) + warp) * (32 * VEC) + (size_t)lane * VEC;
      const bool live = i_raw < n_used;    // false only in the ragged end of the last tile
      const size_t i0 = live ? i_raw : 0;  // dead lanes shadow slot 0 (loads stay in range, nothing is stored)

      // ---- front-batched loads (all independent; MLP = 4 + #columns read) ----
      uint32_t pos[VEC]; float adiv[VEC], ahyd[VEC]; float v[VEC][NV], old[VEC][NV];
      const uint32_t stw = VecIO<VEC>::ldb(p.status + i0);
      VecIO<VEC>::ldu(p.pos + i0, pos);
      VecIO<VEC>::ldf(p.age_div + i0, adiv);
      if (p.enable_leave) VecIO<VEC>::ldf(p.age_hyd + i0, ahyd);
      else {
#pragma unroll
        for (int q = 0; q < VEC; ++q) ahyd[q] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float col[VEC];
        if ((M::write_only_mask >> k) & 1u) {
#pragma unroll
          for (int q = 0; q < VEC; ++q) col[q] = 0.f;
        } else {
          VecIO<VEC>::ldf(p.props + (size_t)k * p.cap + i0, col);
        }
#pragma unroll
        for (int q = 0; q < VEC; ++q) { v[q][k] = col[q]; old[q][k] = col[q]; }
      }
      uint32_t pos_old[VEC]; float adiv_old[VEC], ahyd_old[VEC];
      bool idle[VEC];
      unsigned valid_m = 0, idle_m = 0;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        pos_old[q] = pos[q]; adiv_old[q] = adiv[q]; ahyd_old[q] = ahyd[q];
        const bool valid = live && (i0 + q) < n_used;
        idle[q] = valid && (((stw >> (8 * q)) & 0xffu) == (unsigned)Idle);
        valid_m |= (unsigned)valid << q; idle_m |= (unsigned)idle[q] << q;
        if (!valid) pos[q] = 0;  // slots past n_used hold unspecified bytes: keep the gathers in range
      }
      constexpr unsigned kAll = (1u << VEC) - 1u;

      // ---- compartment rows: leave threshold + model terms, one gather per particle
      float ctab[VEC][CT];
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        const float* row = p.ctab + (size_t)pos[q] * CT;
        if constexpr (CT == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(row)); ctab[q][0] = t.x; ctab[q][1] = t.y; }
        else if constexpr (CT == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(row)); ctab[q][0] = t.x; ctab[q][1] = t.y; ctab[q][2] = t.z; ctab[q][3] = t.w; }
        else {
#pragma unroll
          for (int k = 0; k < CT; ++k) ctab[q][k] = __ldg(row + k);
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
        adiv[q] = idle[q] ? adiv[q] + p.dt_f : adiv[q];  // ages(i,1) += _d_t  (model_kernel.hpp:191)
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
            for (int j = 0; j < NC; ++j) atomicAdd(p.sources + (size_t)j + (size_t)p.n_species * pos[q], w * (double)contrib[q][j]);
          }
      }

      // ---- division: handle_division (particles_container.hpp:559-573) -------
      // warp-aggregated slot allocation: ONE atomic per warp that has a dividing
      // mother (reference: one per mother, a6).  Rows are allocated in ascending
      // particle order inside the warp; final newborn placement is re-ranked by
      // mother index in insert_kernel, so the result does not depend on the
      // order warps hit the atomic.
      if (__any_sync(0xffffffffu, div_nib != 0u)) {
        const unsigned cnt = __popc(div_nib);
        unsigned total;
        const unsigned excl = warp_excl_scan(cnt, total);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&p.st->buf_index, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        unsigned ok_nib = 0, r = 0;
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
              adiv[q] = 0.f;                       // ages(idx1,1) = 0
              ok_nib |= 1u << q;
            } else {
              ++c_over;  // waiting_allocation_particle / Overflow (model_kernel.hpp:253-258)
            }
          }
        }
        // division bitmask: bit (slot & 31) of word (slot >> 5); a word is owned by 32/VEC lanes
        constexpr int LPW = 32 / VEC;
        unsigned word = ok_nib << (VEC * (lane % LPW));
#pragma unroll
        for (int o = 1; o < LPW; o <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, o);
        const unsigned n_ok = __reduce_add_sync(0xffffffffu, __popc(ok_nib));
        if (live && (lane % LPW) == 0 && word != 0u) p.div_mask[(i0 >> 5)] = word;
        if (lane == 0 && n_ok) atomicAdd(&p.tile_div[tile], n_ok);
      }

      // ---- move (all slots, no status check: move_kernel.hpp:392-437) --------
      if (p.enable_move) {
        unsigned mv = 0;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
          const float u1 = u01f(pick4(rw1, (unsigned)((i0 + q) & 3)));
          mv |= (unsigned)(u1 < ctab[q][0]) << q;  // (dt*flow/volume) > rng1
        }
        mv &= valid_m;
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
          ahyd[q] = idle[q] ? (float)((double)ahyd[q] + p.dt) : ahyd[q];  // ages(idx,0) += d_t (double)
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
                ahyd[q] = ahyd[q] * 0.0f;                   // ages(idx,0) *= (1 - leave_mask)
                exit_nib |= 1u << q;
                ++c_exit;
              }
            }
          }
        }
      }

      // ---- write back only what changed -------------------------------------
      const bool all_idle = (idle_m == kAll);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        float col[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) col[q] = v[q][k];
        float* dst = p.props + (size_t)k * p.cap + i0;
        bool ch = false;
        if ((M::write_only_mask >> k) & 1u) ch = idle_m != 0u;
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
      bool ch_ad = false, ch_ah = false, ch_pos = false;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        ch_ad = ch_ad || (__float_as_uint(adiv[q]) != __float_as_uint(adiv_old[q]));
        ch_ah = ch_ah || (__float_as_uint(ahyd[q]) != __float_as_uint(ahyd_old[q]));
        ch_pos = ch_pos || (pos[q] != pos_old[q]);
      }
      if (ch_ad) VecIO<VEC>::stf(p.age_div + i0, adiv);  // unchanged lanes rewrite their own value
      if (ch_ah) VecIO<VEC>::stf(p.age_hyd + i0, ahyd);
      if (ch_pos) {  // Q14: position written only when it changed (never for slots >= n_used: mv is masked)
#pragma unroll
        for (int q = 0; q < VEC; ++q) if (pos[q] != pos_old[q] && ((valid_m >> q) & 1u)) p.pos[i0 + q] = pos[q];
      }
      if (exit_nib) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) if ((exit_nib >> q) & 1u) p.status[i0 + q] = (uint8_t)Exit;
      }
    }
  }

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
      if (lane == 0 && a != 0.0) atomicAdd(p.sources + j, a);
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
    for (uint32_t k = threadIdx.x; k < n_bins; k += kBlock) {
      const double a = s_bins[k];
      if (a != 0.0) atomicAdd(p.sources + k, a);
    }
  }
  // block-local exclusive prefix of tile_div over [t0,t1) -> tile_off, blk_total
  if (warp == 0) {
    unsigned run = 0;
    for (uint32_t base = t0; base < t1; base += 32) {
      const uint32_t t = base + lane;
      const unsigned cnt = (t < t1) ? __ldcg(p.tile_div + t) : 0u;
      unsigned tot;
      const unsigned ex = warp_excl_scan(cnt, tot);
      if (t < t1) p.tile_off[t] = run + ex;
      run += tot;
    }
    if (lane == 0) p.blk_total[blockIdx.x] = run;
  }
  // last block to finish: every block's counters are visible (fence + ticket) -> write the
  // post-cycle plan (compaction trigger, newborn count) for the kernels that follow
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned ticket = atomicAdd(&p.st->done_blocks, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      p.st->done_blocks = 0;
      p.st->clear_n = 0;
      make_plan(p.st, p.min_removal, p.dead_ratio);
    }
  }
}

// plan without a particle pass (ParticlesContainer::force_remove_dead path)
__global__ void plan_kernel(DevState* st, unsigned long long min_removal, double dead_ratio) {
  if (blockIdx.x || threadIdx.x) return;
  make_plan(st, min_removal, dead_ratio);
}

// -----------------------------------------------------------------------------
// Compaction: remove_inactive_particles + CompactParticlesFunctor
// (particles_container.hpp:735-796, 292-385), made exact and deterministic
// (SURVEY Q4): the k-th non-idle slot below new_n (ascending) receives the k-th
// idle particle of the tail [new_n, old_n) counted from the end — the pairing a
// serial execution of the reference functor produces.
//   compact_count : per-tile counts (gaps below new_n, idle in the tail), with
//                   block-local prefix (contiguous tile ranges)
//   compact_src   : tail tiles -> src[k] = slot of the k-th idle from the end
//   compact_move  : low tiles  -> gap with rank k pulls src[k]
// -----------------------------------------------------------------------------
struct CompactParams {
  float* props; size_t cap; int n_var;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  uint32_t* tile_gap_off; uint32_t* tile_idle_off; uint32_t* blk_gap; uint32_t* blk_idle;
  uint32_t* src;
};

__device__ __forceinline__ void compact_tile_flags(const CompactParams& p, uint32_t tile, unsigned long long old_n,
                                                   unsigned long long new_n, unsigned q, bool& gap, bool& tail_idle) {
  const unsigned long long i = (unsigned long long)tile * kTile + q;
  gap = false; tail_idle = false;
  if (i < old_n) {
    const bool is_idle = p.status[i] == (uint8_t)Idle;
    if (i < new_n) gap = !is_idle; else tail_idle = is_idle;
  }
}

// blocks of 1024 threads: thread q handles slot q of the tile
__global__ void __launch_bounds__(1024) compact_count_kernel(const __grid_constant__ CompactParams p) {
  if (!p.st->do_compact) return;
  __shared__ unsigned s_g[32], s_i[32];
  const unsigned long long old_n = p.st->cmp_old_n, new_n = p.st->cmp_new_n;
  const uint32_t n_tiles = p.st->cmp_tiles;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned run_g = 0, run_i = 0;
  for (uint32_t tile = t0; tile < t1; ++tile) {
    bool gap, ti;
    compact_tile_flags(p, tile, old_n, new_n, threadIdx.x, gap, ti);
    const unsigned bg = __popc(__ballot_sync(0xffffffffu, gap)), bi = __popc(__ballot_sync(0xffffffffu, ti));
    if (lane == 0) { s_g[warp] = bg; s_i[warp] = bi; }
    __syncthreads();
    unsigned tg = 0, tii = 0;
    if (warp == 0) {
      tg = __reduce_add_sync(0xffffffffu, s_g[lane]); tii = __reduce_add_sync(0xffffffffu, s_i[lane]);
      if (lane == 0) { p.tile_gap_off[tile] = run_g; p.tile_idle_off[tile] = run_i; }
      run_g += tg; run_i += tii;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { p.blk_gap[blockIdx.x] = run_g; p.blk_idle[blockIdx.x] = run_i; }
}

// exclusive prefix of per-block totals in shared memory (grid <= kMaxGrid): warp 0 scans 32
// entries per step with shuffles; executed by the whole block
__device__ __forceinline__ unsigned block_prefix_of(const uint32_t* blk_tot, unsigned nblk, unsigned b, unsigned* s_tmp,
                                                    unsigned& grand_total) {
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

__global__ void __launch_bounds__(1024) compact_src_kernel(const __grid_constant__ CompactParams p) {
  if (!p.st->do_compact) return;
  __shared__ unsigned s_pref[kMaxGrid + 1];
  __shared__ unsigned s_w[32];
  const unsigned long long old_n = p.st->cmp_old_n, new_n = p.st->cmp_new_n;
  const uint32_t n_tiles = p.st->cmp_tiles;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  unsigned total_idle;
  const unsigned blk_off = block_prefix_of(p.blk_idle, gridDim.x, blockIdx.x, s_pref, total_idle);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.st->cmp_total_idle = total_idle;
  const uint32_t first_tail_tile = (uint32_t)(new_n / kTile);
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t tile = (t0 > first_tail_tile ? t0 : first_tail_tile); tile < t1; ++tile) {
    bool gap, ti;
    compact_tile_flags(p, tile, old_n, new_n, threadIdx.x, gap, ti);
    const unsigned bal = __ballot_sync(0xffffffffu, ti);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    unsigned woff = 0;
    for (unsigned k = 0; k < warp; ++k) woff += s_w[k];
    if (ti) {
      const unsigned asc = blk_off + p.tile_idle_off[tile] + woff + __popc(bal & ((1u << lane) - 1u));
      p.src[total_idle - 1u - asc] = (uint32_t)((unsigned long long)tile * kTile + threadIdx.x);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) compact_move_kernel(const __grid_constant__ CompactParams p) {
  if (!p.st->do_compact) return;
  __shared__ unsigned s_pref[kMaxGrid + 1];
  __shared__ unsigned s_w[32];
  const unsigned long long old_n = p.st->cmp_old_n, new_n = p.st->cmp_new_n;
  const uint32_t n_tiles = p.st->cmp_tiles;
  const uint32_t t0 = (uint32_t)(((unsigned long long)blockIdx.x * n_tiles) / gridDim.x);
  const uint32_t t1 = (uint32_t)(((unsigned long long)(blockIdx.x + 1) * n_tiles) / gridDim.x);
  unsigned total_gap;
  const unsigned blk_off = block_prefix_of(p.blk_gap, gridDim.x, blockIdx.x, s_pref, total_gap);
  const uint32_t last_low_tile = (uint32_t)((new_n + kTile - 1) / kTile);  // exclusive
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t tile = t0; tile < t1; ++tile) {
    const unsigned long long i = (unsigned long long)tile * kTile + threadIdx.x;
    if (tile < last_low_tile) {
      bool gap, ti;
      compact_tile_flags(p, tile, old_n, new_n, threadIdx.x, gap, ti);
      const unsigned bal = __ballot_sync(0xffffffffu, gap);
      if (lane == 0) s_w[warp] = __popc(bal);
      __syncthreads();
      unsigned woff = 0;
      for (unsigned k = 0; k < warp; ++k) woff += s_w[k];
      if (gap) {
        const unsigned k = blk_off + p.tile_gap_off[tile] + woff + __popc(bal & ((1u << lane) - 1u));
        if (k >= p.st->cmp_total_idle) {
          atomicOr(&p.st->error, 2u);  // inactive counter inconsistent with the status column
        } else {
          const size_t r = p.src[k];
          p.status[i] = (uint8_t)Idle;
          p.pos[i] = p.pos[r];
          for (int c = 0; c < p.n_var; ++c) p.props[(size_t)c * p.cap + i] = p.props[(size_t)c * p.cap + r];
          p.age_hyd[i] = p.age_hyd[r];
          p.age_div[i] = p.age_div[r];
        }
      }
      __syncthreads();
    }
  }
}

// -----------------------------------------------------------------------------
// insert: merge_buffer + InsertFunctor (particles_container.hpp:575-599,
// 403-443).  Newborn of mother i goes to new_n + (number of dividing mothers with
// a smaller slot index) — the order the reference's buffer has under serial
// execution.
// -----------------------------------------------------------------------------
struct InsertParams {
  float* props; size_t cap; int n_var;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  const float* buf_props; size_t buf_stride; const uint32_t* buf_pos; const uint32_t* buf_mother;
  uint32_t* div_mask; uint32_t* tile_div; const uint32_t* tile_off; const uint32_t* blk_total;
  int count_step;  // 1 when called from a cycle, 0 from force_remove_dead
};

__global__ void __launch_bounds__(256) post_kernel(const __grid_constant__ InsertParams p) {
  __shared__ unsigned s_pref[kMaxGrid + 1];
  const unsigned long long n_add = p.st->n_add;
  const unsigned long long base = p.st->cmp_new_n;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  // slots [new_n, old_n) left the container in a compaction: mark them Idle so that appended
  // newborns never inherit a stale status (the reference relies on zero-initialised storage,
  // particles_container.hpp:403-443).  Newborn slots below are written Idle as well, so the
  // two writers agree where they overlap.
  if (p.st->do_compact) {
    const unsigned long long old_n = p.st->cmp_old_n;
    for (unsigned long long i = base + gtid; i < old_n; i += gstride) p.status[i] = (uint8_t)Idle;
  }
  if (n_add) {  // uniform across the grid
    const unsigned G = p.st->cyc_grid;
    const unsigned T = p.st->cyc_tiles;
    unsigned total;
    block_prefix_of(p.blk_total, G, 0, s_pref, total);
    for (unsigned long long j = gtid; j < n_add; j += gstride) {
      const uint32_t mother = p.buf_mother[j];
      const uint32_t tile = mother >> 10;
      const unsigned b = (unsigned)((((unsigned long long)tile + 1ull) * G - 1ull) / T);  // owner block of the tile
      const uint32_t* words = p.div_mask + (size_t)tile * (kTile / 32);
      const unsigned wi = (mother & (kTile - 1)) >> 5, bit = mother & 31u;
      unsigned rank = 0;
      for (unsigned k = 0; k < wi; ++k) rank += __popc(words[k]);
      rank += __popc(words[wi] & ((1u << bit) - 1u));
      const unsigned long long dst = base + s_pref[b] + p.tile_off[tile] + rank;
      for (int c = 0; c < p.n_var; ++c) p.props[(size_t)c * p.cap + dst] = p.buf_props[(size_t)c * p.buf_stride + j];
      p.pos[dst] = p.buf_pos[j];
      p.age_hyd[dst] = 0.f; p.age_div[dst] = 0.f;  // InsertFunctor: both ages reset
      p.status[dst] = (uint8_t)Idle;
    }
  }
  // commit (one thread).  Only fields no other thread of this kernel reads are modified.
  if (gtid == 0) {
    DevState* st = p.st;
    if (st->do_compact) { st->inactive -= (st->cmp_old_n - st->cmp_new_n); st->n_compactions += 1; }
    st->n_used = base + n_add;
    st->total_new += n_add;
    if (n_add > st->clear_n) st->clear_n = n_add;  // bits cleared by the next pre_step
    st->step += (unsigned long long)p.count_step;
  }
}

// -----------------------------------------------------------------------------
// Domain tables: ReactorDomain::update (mc/src/domain.cpp:43-74) -> derived
// single-precision tables that reproduce the double-precision comparisons
// bit-exactly for float uniforms:
//   (dt*flow/volume) > (double)u   <=>  u < ceil_f32(dt*flow/volume)   (compartment_table_kernel)
//   (double)u > cdf                <=>  u > floor_f32(cdf)
// -----------------------------------------------------------------------------
__global__ void derive_cdf_table_kernel(const double* cdf, float* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __double2float_rd(cdf[i]);
}

// -----------------------------------------------------------------------------
// mc_init_first: InitFunctor (mc/src/unit.cpp:102-144): M::init, random
// compartment, total-mass reduce.
// -----------------------------------------------------------------------------
template <class M>
__global__ void __launch_bounds__(256) init_kernel(float* props, size_t cap, uint32_t* pos, uint8_t* status, float* age_hyd,
                                                   float* age_div, unsigned long long n, uint32_t n_comp_hi, const float* linit,
                                                   uint32_t seed_lo, uint32_t seed_hi, uint32_t rank, DevState* st) {
  double m = 0.0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float v[M::n_var];
    Gen gen(seed_lo, seed_hi, rank, (uint32_t)i, 0xFFFFFFFFu, 2u);
    M::init(gen, (size_t)i, RegRow{v}, ConfigView{linit});
    m += M::mass((size_t)i, RegRow{v});
    const uint32_t c = (uint32_t)gen.urand64(0ull, (unsigned long long)n_comp_hi);
#pragma unroll
    for (int k = 0; k < M::n_var; ++k) props[(size_t)k * cap + i] = v[k];
    pos[i] = c; status[i] = (uint8_t)Idle; age_hyd[i] = 0.f; age_div[i] = 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if ((threadIdx.x & 31) == 0 && m != 0.0) atomicAdd(&st->init_mass, m);
}

// get_repartition: NcellFunctor (mc/src/unit.cpp:48-100, 190-230)
__global__ void __launch_bounds__(256) repartition_kernel(const uint32_t* pos, const uint8_t* status, const DevState* st,
                                                          unsigned long long* out) {
  const unsigned long long n = st->n_used;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    if (status[i] == (uint8_t)Idle) atomicAdd(out + pos[i], 1ull);
}

// u64 <-> u32 position conversion for the host boundary
__global__ void pos_narrow_kernel(const unsigned long long* in, uint32_t* out, size_t n, uint32_t n_comp, unsigned int* err) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const unsigned long long v = in[i]; if (v >= n_comp) atomicOr(err, 1u); out[i] = (uint32_t)v; }
}
__global__ void pos_widen_kernel(const uint32_t* in, unsigned long long* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void count_inactive_kernel(const uint8_t* status, size_t n, DevState* st) {
  unsigned c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += status[i] != (uint8_t)Idle;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&st->inactive, (unsigned long long)c);
}
__global__ void fill_u8_kernel(uint8_t* p, size_t n, uint8_t v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace bmc
