// Non-template kernels of the particle step (compaction, spawn/commit, domain tables,
// repartition, host-boundary conversions).  Included by bmc_api.cu only; the per-model
// translation units include bmc_kernels.cuh (template kernels) alone.
#pragma once
#include "bmc_kernels.cuh"

namespace bmc {

// ParticlesContainer::force_remove_dead (particles_container.hpp:463-468) and any other post-cycle
// without a particle pass: the second phase of the step kernel on its own (cooperative launch)
__global__ void __launch_bounds__(kBlock) post_only_kernel(const __grid_constant__ PostParams p) { post_cycle_body(p); }

// host events that change n_used or the capacity (set/init/load particles, resize): next step's buffer room.
// `reset_logical`: the container was (re)constructed with n_used particles -> the reference's constructor extents
// (_resize(n_particle) + __allocate_buffer__ on an empty container, particles_container.hpp:692-727)
__global__ void prepare_kernel(DevState* st, unsigned long long cap, unsigned long long buf_cap, int reset_logical,
                               double allocation_factor, double buffer_ratio, PinState* pin) {
  if (blockIdx.x || threadIdx.x) return;
  const unsigned long long n = st->n_used;
  if (reset_logical) {
    unsigned long long la = 0ull, lb = 0ull;
    if (n) logical_grow(n, allocation_factor, buffer_ratio, la, lb);
    st->logical_alloc = la; st->logical_buf = lb;
  }
  const unsigned long long room = cap > n ? cap - n : 0ull;
  const unsigned long long eff = st->logical_buf < buf_cap ? st->logical_buf : buf_cap;
  st->buf_cap_eff = eff < room ? eff : room;
  st->buf_index = 0; st->step_exit = 0; st->step_waiting = 0;
  if (pin) pin_write(pin, st->step, n, 0ull, st->logical_alloc, st->logical_buf, st->error);
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
  PeerExchange px;  // multi-GPU: a pending all-reduce of the sources is finished HERE (consume_epoch != 0): every element
                    // is the sum over the ranks of the peers' published buffers, in rank order
};
// source term of element k: the local vector, or — when the all-reduce of the last cycle is still pending — the sum of
// what every rank published (bmc_kernels.cuh: PeerExchange).  Called by EVERY thread of the block (it contains a barrier).
__device__ __forceinline__ double liquid_source(const LiquidParams& p, uint32_t k, bool active, unsigned int* error) {
  if (p.px.world > 1 && p.px.consume_epoch) {
    if ((int)threadIdx.x < p.px.world && (int)threadIdx.x != p.px.rank) {
      volatile unsigned long long* f = p2p_flag(p.px.base[threadIdx.x]);
      const long long t0 = clock64();
      while (*f < p.px.consume_epoch) {
        if (clock64() - t0 > p.px.spin_limit) { atomicOr(error, 4u); break; }
      }
      __threadfence_system();
    }
    __syncthreads();
    const unsigned par = (unsigned)(p.px.consume_epoch & 1ull);
    double a = 0.0;
    if (active)
      for (int r = 0; r < p.px.world; ++r) a += *reinterpret_cast<volatile double*>(p2p_buf(p.px.base[r], p.n_species * p.n_comp, par) + k);
    p2p_check_not_overrun(p.px, error);
    return a;
  }
  return active ? p.sources[k] : 0.0;
}
__global__ void __launch_bounds__(256) liquid_step_kernel(const __grid_constant__ LiquidParams p, unsigned int* error) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = k < p.n_species * p.n_comp;
  double src = liquid_source(p, k, active, error), sink = 0.0;
  if (!active) return;
  const uint32_t s = k % p.n_species, j = k / p.n_species;
  for (int f = 0; f < p.n_feeds; ++f) {  // set_feed / set_sink
    if (p.feeds[f].input_position == j && p.feeds[f].species == s) src += p.feeds[f].flow * p.feeds[f].concentration;
    if (p.feeds[f].has_output && p.feeds[f].first_of_feed && p.feeds[f].output_position == j) sink += p.feeds[f].flow;
  }
  double dm = 0.0;
  for (uint32_t e = p.csc_ptr[j]; e < p.csc_ptr[j + 1]; ++e) dm += p.c_old[s + p.n_species * p.csc_row[e]] * p.csc_val[e];
  const double c = p.c_old[k];
  dm = (dm - c * sink) + src;  // `c * m_transition - c * sink + _sources`, coefficient-wise, left to right (implScalar.cpp:259)
  const double m = p.mass[k] + p.dt * dm;
  p.mass[k] = m;
  p.c_new[k] = m * (1.0 / p.vol[j]);
  p.sources[k] = 0.0;  // clearContribution
}
// -----------------------------------------------------------------------------
// Two-phase step: SimulationUnit::ode_step with a gas phase (simulation.model.cpp:131-154):
//   mtr = (kla o (Cg o Henry - Cl)) * diag(V_liquid)                      MassTransferModel::gas_liquid_mass_transfer
//                                                                         (hydro/mass_transfer.cpp:143-161)
//   gas    : dm/dt = Cg*Mg - Cg*sink_g + sources_g + (-1) * mtr           performStepGL(d_t, mtr, Sign::GasToLiquid)
//   liquid : dm/dt = Cl*Ml - Cl*sink_l + sources_l + (+1) * mtr           performStepGL(d_t, mtr, Sign::LiquidToGas)
//            (implScalar.cpp:229-249), then clearNegs on the liquid (:270-296): values in (-5e-7, 0) become 0
// Everything is evaluated from the concentrations BEFORE the step (both phases are double-buffered), one thread per
// (species, compartment), inflow terms accumulated in COO order like the single-phase kernel.
// -----------------------------------------------------------------------------
struct GasLiquidParams {
  LiquidParams liq;                 // liquid phase: as in liquid_step_kernel (sources = particle source terms + feeds)
  const double* g_old; double* g_new; double* g_mass; const double* g_vol;
  const uint32_t* g_csc_ptr; const uint32_t* g_csc_row; const double* g_csc_val;
  int n_gas_feeds; FeedDev gas_feeds[kMaxFlows];
  const double* kla; const double* henry; double* mtr;
};
__global__ void __launch_bounds__(256) gas_liquid_step_kernel(const __grid_constant__ GasLiquidParams p, unsigned int* error) {
  const LiquidParams& l = p.liq;
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = k < l.n_species * l.n_comp;
  const double src_particles = liquid_source(l, k, active, error);
  if (!active) return;
  const uint32_t s = k % l.n_species, j = k / l.n_species;
  const double cl = l.c_old[k], cg = p.g_old[k];
  const double mtr = p.kla[k] * (cg * p.henry[s] - cl) * l.vol[j];
  p.mtr[k] = mtr;
  auto feed_terms = [&](const FeedDev* feeds, int n, double& src, double& sink) {
    for (int f = 0; f < n; ++f) {  // set_feed / set_sink
      if (feeds[f].input_position == j && feeds[f].species == s) src += feeds[f].flow * feeds[f].concentration;
      if (feeds[f].has_output && feeds[f].first_of_feed && feeds[f].output_position == j) sink += feeds[f].flow;
    }
  };
  {  // gas
    double src = 0.0, sink = 0.0;
    feed_terms(p.gas_feeds, p.n_gas_feeds, src, sink);
    double dm = 0.0;
    for (uint32_t e = p.g_csc_ptr[j]; e < p.g_csc_ptr[j + 1]; ++e) dm += p.g_old[s + l.n_species * p.g_csc_row[e]] * p.g_csc_val[e];
    dm = (dm - cg * sink) + src;  // `c * m_transition - c * sink + _sources + float(sign) * mtr` (implScalar.cpp:237-238)
    dm += -1.0 * mtr;
    const double m = p.g_mass[k] + l.dt * dm;
    p.g_mass[k] = m;
    p.g_new[k] = m * (1.0 / p.g_vol[j]);
  }
  {  // liquid
    double src = src_particles, sink = 0.0;
    feed_terms(l.feeds, l.n_feeds, src, sink);
    double dm = 0.0;
    for (uint32_t e = l.csc_ptr[j]; e < l.csc_ptr[j + 1]; ++e) dm += l.c_old[s + l.n_species * l.csc_row[e]] * l.csc_val[e];
    dm = (dm - cl * sink) + src;
    dm += 1.0 * mtr;
    const double m = l.mass[k] + l.dt * dm;
    l.mass[k] = m;
    double c = m * (1.0 / l.vol[j]);
    if (c < 0.0 && fabs(c) < 1e-4 * 5e-3) c = 0.0;  // clearNegs: TOL = scheme_relative_error * max_species_value
    l.c_new[k] = c;
    l.sources[k] = 0.0;  // clearContribution
  }
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



// Peer exchange outside the step kernel (bmc_kernels.cuh: PeerExchange): publish the current sources when no cycle
// has published them, and/or finish the all-reduce when something other than the next cycle reads the sources first.
__global__ void __launch_bounds__(1024) p2p_exchange_kernel(const __grid_constant__ PeerExchange x, double* sources, uint32_t n, DevState* st,
                                                            unsigned long long* mirror, unsigned long long tag) {
  if (x.publish_epoch) p2p_publish(x, sources, n);
  if (x.consume_epoch) { __syncthreads(); p2p_consume(x, sources, n, &st->error, mirror, tag); }
}

}  // namespace bmc
