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

struct FullTile { static constexpr bool value = true; };
struct RaggedTile { static constexpr bool value = false; };

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
  uint32_t stage_offset;  // byte offset of the cp.async staging buffers inside dynamic shared memory
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
  static __device__ __forceinline__ void ldf_plain(const float* p, float (&v)[1]) { v[0] = *p; }
  static __device__ __forceinline__ void ldu_plain(const uint32_t* p, uint32_t (&v)[1]) { v[0] = *p; }
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

template <class M> __device__ __forceinline__ void pre_step_body(const PreParams& p) {
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
__host__ __device__ constexpr int popcount_c(uint32_t x) { return x == 0u ? 0 : (int)(x & 1u) + popcount_c(x >> 1); }
// columns actually loaded per slot: every property that is not write-only
template <class M> struct ReadCols {
  static constexpr uint32_t all = M::n_var >= 32 ? 0xffffffffu : ((1u << M::n_var) - 1u);
  static constexpr int value = M::n_var - popcount_c(M::write_only_mask & all);
};
// bytes of one staging buffer of the software pipeline: pos, age_div, age_hyd + the read columns
template <class M, int VEC> struct StageBytes { static constexpr size_t value = (size_t)(3 + ReadCols<M>::value) * kBlock * 4 * VEC; };

template <int BYTES> __device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class M, int VEC, bool PIPE> __device__ __forceinline__ void cycle_body(const CycleParams& p) {
  constexpr int NV = M::n_var, NC = M::n_c, CT = 1 + M::n_pre;
  constexpr int SUB = kTile / (kBlock * VEC);  // sub-iterations per tile
  constexpr int kColStride = kBlock * 4 * VEC;  // bytes between staged columns
  constexpr size_t kStage = StageBytes<M, VEC>::value;
  extern __shared__ double s_bins[];           // [n_species * n_comp] when bins_in_smem, then 2 staging buffers
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

  // PIPE: two-stage software pipeline.  Each thread copies ITS OWN next group of slots
  // global -> shared with cp.async (LDGSTS: no registers held while the bytes are in flight),
  // then computes the current group out of shared memory.  A thread only ever reads what it
  // copied itself, so cp.async.wait_group is the only synchronisation needed.
  unsigned char* const s_stage = reinterpret_cast<unsigned char*>(s_bins) + p.stage_offset;
  const uint32_t n_used32 = (uint32_t)n_used;
  auto slot_base = [&](uint32_t tile, int sub) -> uint32_t {
    return tile * (uint32_t)kTile + ((uint32_t)sub * (kBlock / 32) + warp) * (32 * VEC) + lane * VEC;
  };
  auto issue = [&](uint32_t it, int buf) -> uint32_t {  // returns the status bytes of that group
    const uint32_t i_raw = slot_base(t0 + it / SUB, (int)(it % SUB));
    const size_t i0 = i_raw < n_used32 ? i_raw : 0u;
    unsigned char* dst = s_stage + (size_t)buf * kStage + threadIdx.x * (4 * VEC);
    cp_async<4 * VEC>(dst, p.pos + i0);
    cp_async<4 * VEC>(dst + kColStride, p.age_div + i0);
    if (p.enable_leave) cp_async<4 * VEC>(dst + 2 * kColStride, p.age_hyd + i0);
    int c = 3;
#pragma unroll
    for (int k = 0; k < NV; ++k)
      if (!((M::write_only_mask >> k) & 1u)) { cp_async<4 * VEC>(dst + c * kColStride, p.props + (size_t)k * p.cap + i0); ++c; }
    cp_async_commit();
    return VecIO<VEC>::ldb(p.status + i0);
  };

  {
    // Every tile but the last is entirely below n_used: the body is instantiated twice so that the
    // common case carries no per-slot range checks (FULL), the ragged tail keeps them.
    auto body = [&](auto full_tag, const uint32_t tile, const int sub, const int buf, const uint32_t stw_in) {
      constexpr bool FULL = decltype(full_tag)::value;
      const uint32_t i_raw = slot_base(tile, sub);
      const bool live = FULL || i_raw < n_used32;  // false only in the ragged end of the last tile
      const size_t i0 = live ? i_raw : 0u;         // dead lanes shadow slot 0 (loads stay in range, nothing is stored)

      uint32_t pos[VEC]; float adiv[VEC], ahyd[VEC]; float v[VEC][NV], old[VEC][NV];
      uint32_t stw;
      if constexpr (PIPE) {
        // ---- operands were staged in shared memory by this thread one iteration ago ----
        const unsigned char* src = s_stage + (size_t)buf * kStage + threadIdx.x * (4 * VEC);
        stw = stw_in;
        VecIO<VEC>::ldu_plain(reinterpret_cast<const uint32_t*>(src), pos);
        VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src + kColStride), adiv);
        if (p.enable_leave) VecIO<VEC>::ldf_plain(reinterpret_cast<const float*>(src + 2 * kColStride), ahyd);
        else {
#pragma unroll
          for (int q = 0; q < VEC; ++q) ahyd[q] = 0.f;
        }
        int c = 3;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          float col[VEC];
          if ((M::write_only_mask >> k) & 1u) {
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
        // ---- front-batched global loads (all independent; MLP = 4 + #columns read) ----
        stw = VecIO<VEC>::ldb(p.status + i0);
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
      }
      uint32_t pos_old[VEC]; float adiv_old[VEC], ahyd_old[VEC];
      bool idle[VEC];
      unsigned valid_m = 0, idle_m = 0;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        pos_old[q] = pos[q]; adiv_old[q] = adiv[q]; ahyd_old[q] = ahyd[q];
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
    };
    if constexpr (PIPE) {
      const uint32_t n_it = (t1 - t0) * SUB;
      uint32_t stw_next = n_it ? issue(0, 0) : 0u;
#pragma unroll 1
      for (uint32_t it = 0; it < n_it; ++it) {
        const uint32_t stw_cur = stw_next;
        if (it + 1 < n_it) { stw_next = issue(it + 1, (int)((it + 1) & 1u)); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        const uint32_t tile = t0 + it / SUB;
        if ((unsigned long long)(tile + 1) * kTile <= n_used) body(FullTile{}, tile, (int)(it % SUB), (int)(it & 1u), stw_cur);
        else body(RaggedTile{}, tile, (int)(it % SUB), (int)(it & 1u), stw_cur);
      }
    } else {
#pragma unroll 1
      for (uint32_t tile = t0; tile < t1; ++tile) {
        if ((unsigned long long)(tile + 1) * kTile <= n_used) {
#pragma unroll 1
          for (int sub = 0; sub < SUB; ++sub) body(FullTile{}, tile, sub, 0, 0u);
        } else {
#pragma unroll 1
          for (int sub = 0; sub < SUB; ++sub) body(RaggedTile{}, tile, sub, 0, 0u);
        }
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

// __global__ entry points of the built-in models (the NVRTC path of user models wraps the same
// bodies in extern "C" kernels, see bmc_udf.cu)
template <class M> __global__ void __launch_bounds__(256) pre_step_kernel(const __grid_constant__ PreParams p) { pre_step_body<M>(p); }
template <class M, int VEC, int MINB, bool PIPE>
__global__ void __launch_bounds__(kBlock, MINB) cycle_kernel(const __grid_constant__ CycleParams p) { cycle_body<M, VEC, PIPE>(p); }
template <class M> __global__ void __launch_bounds__(256) init_kernel(const __grid_constant__ InitParams p) { init_body<M>(p); }

}  // namespace bmc
