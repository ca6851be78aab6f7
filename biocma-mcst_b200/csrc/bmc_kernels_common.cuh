// Non-template kernels of the particle step (compaction, spawn/commit, domain tables,
// repartition, host-boundary conversions).  Included by bmc_api.cu only; the per-model
// translation units include bmc_kernels.cuh (template kernels) alone.
#pragma once
#include <cooperative_groups.h>
#include "bmc_kernels.cuh"

namespace bmc {

// plan without a particle pass (ParticlesContainer::force_remove_dead path)
__global__ void plan_kernel(DevState* st, unsigned long long min_removal, double dead_ratio) {
  if (blockIdx.x || threadIdx.x) return;
  make_plan(st, min_removal, dead_ratio);
}

// -----------------------------------------------------------------------------
// post_cycle: everything of SimulationUnit::post_cycle (simulation.hpp:213-239) after the
// particle pass, in ONE cooperative launch:
//
//   [only when the plan written by the cycle kernel's last block says so]
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
//   commit: container counters, the next step's buffer room, one more entry of the age tables.
//
// Blocks of 256 threads own contiguous ranges of 1024-slot tiles; thread t handles slots
// t, t+256, t+512, t+768 of a tile ("virtual warp" vw = 8*r + warp covers 32 consecutive slots).
// -----------------------------------------------------------------------------
struct PostParams {
  float* props; size_t cap; int n_var;
  uint32_t* pos; uint8_t* status; float* age_hyd; float* age_div;
  DevState* st;
  // compaction scratch
  uint32_t* tile_gap_off; uint32_t* tile_idle_off; uint32_t* blk_gap; uint32_t* blk_idle; uint32_t* src;
  // division buffer + ranking data written by the cycle kernel
  const float* buf_props; size_t buf_stride; const uint32_t* buf_pos; const uint32_t* buf_mother;
  const uint32_t* div_mask; const uint32_t* tile_off; const uint32_t* blk_total;
  unsigned long long buf_cap;
  int count_step;  // 1 when called from a cycle, 0 from force_remove_dead
  // step-stamped ages (bmc_kernels.cuh): stamp given to newborns (0 = eager float ages, bits of 0.f)
  // and the per-step extension of the age tables A_div / A_hyd
  uint32_t newborn_stamp;
  float* tab_div; float* tab_hyd; uint32_t tab_idx; int tab_extend; int enable_leave; float dt_f; double dt;
};

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

// flags of the four slots a thread owns in `tile` + per-virtual-warp counts in s_w[32];
// returns (by reference) the ballots; ends with a barrier so that s_w is complete
template <bool WANT_GAP>
__device__ __forceinline__ void tile_flags(const PostParams& p, uint32_t tile, unsigned long long old_n, unsigned long long new_n,
                                           unsigned (&bal)[4], bool (&flag)[4], unsigned* s_w) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const unsigned long long i = (unsigned long long)tile * kTile + (unsigned)r * 256u + threadIdx.x;
    bool f = false;
    if (i < old_n) {
      const bool is_idle = p.status[i] == (uint8_t)Idle;
      f = WANT_GAP ? (i < new_n && !is_idle) : (i >= new_n && is_idle);
    }
    flag[r] = f;
    bal[r] = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[r * 8 + warp] = __popc(bal[r]);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) post_cycle_kernel(const __grid_constant__ PostParams p) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  __shared__ unsigned s_pref[kMaxGrid + 1];
  __shared__ unsigned s_w[32];
  DevState* const st = p.st;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long gstride = (unsigned long long)gridDim.x * blockDim.x;
  // plan of this post-cycle: written by the last block of the cycle kernel (or plan_kernel), constant here
  const bool do_compact = st->do_compact != 0;
  const unsigned long long old_n = st->cmp_old_n, new_n = st->cmp_new_n, n_add = st->n_add;

  if (do_compact) {  // uniform across the grid
    const uint32_t n_tiles = st->cmp_tiles;
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
    grid.sync();
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
            for (unsigned k = 0; k < (unsigned)r * 8u + warp; ++k) woff += s_w[k];
            const unsigned asc = blk_off + p.tile_idle_off[tile] + woff + __popc(bal[r] & ((1u << lane) - 1u));
            p.src[total_idle - 1u - asc] = (uint32_t)((unsigned long long)tile * kTile + (unsigned)r * 256u + threadIdx.x);
          }
        }
        __syncthreads();
      }
    }
    grid.sync();
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
            for (unsigned k = 0; k < (unsigned)r * 8u + warp; ++k) woff += s_w[k];
            const unsigned k = blk_off + p.tile_gap_off[tile] + woff + __popc(bal[r] & ((1u << lane) - 1u));
            if (k >= total_idle) {
              atomicOr(&st->error, 2u);  // inactive counter inconsistent with the status column
            } else {
              const size_t i = (size_t)tile * kTile + (unsigned)r * 256u + threadIdx.x;
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
    grid.sync();
    // slots [new_n, old_n) left the container: mark them Idle so that appended newborns never inherit
    // a stale status (the reference relies on zero-initialised storage, particles_container.hpp:403-443).
    // Newborn slots below are written Idle as well, so the two writers agree where they overlap.
    for (unsigned long long i = new_n + gtid; i < old_n; i += gstride) p.status[i] = (uint8_t)Idle;
  }

  if (n_add) {  // uniform across the grid
    const unsigned G = st->cyc_grid;
    const unsigned T = st->cyc_tiles;
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
      const unsigned long long dst = new_n + s_pref[b] + p.tile_off[tile] + rank;
      for (int c = 0; c < p.n_var; ++c) p.props[(size_t)c * p.cap + dst] = p.buf_props[(size_t)c * p.buf_stride + j];
      p.pos[dst] = p.buf_pos[j];
      // InsertFunctor: both ages reset (eager: 0.f; stamped: the newborn ages from the next step on)
      reinterpret_cast<uint32_t*>(p.age_hyd)[dst] = p.newborn_stamp;
      reinterpret_cast<uint32_t*>(p.age_div)[dst] = p.newborn_stamp;
      p.status[dst] = (uint8_t)Idle;
    }
  }
  // commit (one thread).  Only fields no other thread of this kernel reads are modified.
  if (gtid == 0) {
    if (do_compact) { st->inactive -= (old_n - new_n); st->n_compactions += 1; }
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

// host events that change n_used or the capacity (set/init particles, resize): next step's buffer room
__global__ void prepare_kernel(DevState* st, unsigned long long cap, unsigned long long buf_cap) {
  if (blockIdx.x || threadIdx.x) return;
  const unsigned long long n = st->n_used;
  const unsigned long long room = cap > n ? cap - n : 0ull;
  st->buf_cap_eff = buf_cap < room ? buf_cap : room;
  st->buf_index = 0; st->step_exit = 0; st->step_waiting = 0;
}

// ---- step-stamped ages -> floats -------------------------------------------------
__device__ __forceinline__ float age_from_stamp(uint32_t s, const float* tab, uint32_t now) {
  uint32_t k = (s & kFrozen) ? (s & ~kFrozen) : now - s;
  if (k > now) k = now;  // slots beyond n_used hold unspecified stamps: stay inside the table
  return tab[k];
}
__global__ void ages_read_kernel(const float* col, const float* tab, uint32_t now, float* out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = age_from_stamp(reinterpret_cast<const uint32_t*>(col)[i], tab, now);
}
__global__ void ages_to_eager_kernel(float* col, const float* tab, uint32_t now, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) col[i] = age_from_stamp(reinterpret_cast<const uint32_t*>(col)[i], tab, now);
}
__global__ void ages_init_stamps_kernel(const uint8_t* status, float* age_hyd, float* age_div, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const uint32_t s = status[i] == (uint8_t)Idle ? 0u : kFrozen;  // non-idle particles never age
    reinterpret_cast<uint32_t*>(age_hyd)[i] = s;
    reinterpret_cast<uint32_t*>(age_div)[i] = s;
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
// Liquid phase: ScalarSimulation::performStep (implScalar.cpp:251-266) preceded by the
// scalar part of update_feed (simulation.model.cpp:55-69) and followed by clearContribution
// (sync.cpp:89-116).  One thread per (species, compartment); the transition matrix is stored
// by destination column (CSC, entries in COO order) so that every element accumulates its
// inflow terms in the same order a sequential COO sweep does -> bit-identical to the oracle.
// -----------------------------------------------------------------------------
struct FeedDev { uint32_t species, input_position, output_position; int has_output, first_of_feed; double flow, concentration; };
struct LiquidParams {
  const double* c_old; double* c_new; double* mass; const double* vol; double* sources;
  const uint32_t* csc_ptr; const uint32_t* csc_row; const double* csc_val;
  uint32_t n_species, n_comp; double dt; int n_feeds; FeedDev feeds[kMaxFlows];
};
__global__ void __launch_bounds__(256) liquid_step_kernel(const __grid_constant__ LiquidParams p) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n_species * p.n_comp) return;
  const uint32_t s = k % p.n_species, j = k / p.n_species;
  double src = p.sources[k], sink = 0.0;
  for (int f = 0; f < p.n_feeds; ++f) {  // set_feed / set_sink
    if (p.feeds[f].input_position == j && p.feeds[f].species == s) src += p.feeds[f].flow * p.feeds[f].concentration;
    if (p.feeds[f].has_output && p.feeds[f].first_of_feed && p.feeds[f].output_position == j) sink += p.feeds[f].flow;
  }
  double dm = 0.0;
  for (uint32_t e = p.csc_ptr[j]; e < p.csc_ptr[j + 1]; ++e) dm += p.c_old[s + p.n_species * p.csc_row[e]] * p.csc_val[e];
  const double c = p.c_old[k];
  dm += -c * sink + src;
  const double m = p.mass[k] + p.dt * dm;
  p.mass[k] = m;
  p.c_new[k] = m * (1.0 / p.vol[j]);
  p.sources[k] = 0.0;  // clearContribution
}
__global__ void liquid_mass_kernel(const double* c, const double* vol, double* mass, uint32_t n_species, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) mass[k] = c[k] * vol[k / n_species];
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
