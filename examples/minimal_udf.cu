// Example user-defined model for BMC_MODEL_UDF: the model of the reference's
// apps/udf_model/minimal.cpp:59-121 (a fixed-length cell with a saturating
// length increment) written against include/bmc_udf.cuh.  Selected with
//   biocma_b200 -mn udf_model   + env BIOMC_LIB_UDF=<this file>
// or bmc_config::udf_source_path.  Compiled at bmc_create() by NVRTC (sm_100a).
// tests/test_udf_gpu.py checks it bit-exactly against the oracle's restatement.
#include "bmc_udf.cuh"

namespace {
using namespace Models;
using FloatType = Models::UdfModel::FloatType;

constexpr FloatType yield_x_s = 0.5;           // glucose -> biomass
constexpr FloatType l_dot_max = 2e-6 / 3600.;  // m/s
constexpr FloatType l_max_m = 2e-6;            // m
constexpr FloatType k = 1e-3;
constexpr FloatType d_m = 0.6e-6;
constexpr FloatType lin_density = c_linear_density(static_cast<FloatType>(1000), d_m);
constexpr FloatType phi_s_max = (l_dot_max * lin_density) / yield_x_s;  // kg/s

enum class particle_var : uint8_t { length = 0, l_max, __COUNT__ };

constexpr std::size_t _set_nvar() { return static_cast<size_t>(particle_var::__COUNT__); }
constexpr std::size_t _set_nc() { return 1; }

void _init_udf(const MC::pool_type& random_pool, std::size_t idx, const UdfModel::SelfParticle& arr,
               const UdfModel::Config& config) {
  GET_PROPERTY(particle_var::length) = config(idx, 0);
  GET_PROPERTY(particle_var::l_max) = l_max_m;
}

MC::Status _update_udf(const MC::pool_type& random_pool, float d_t, std::size_t idx, const UdfModel::SelfParticle& arr,
                       const UdfModel::SelfContribs& arr_contribs, const std::size_t position_index,
                       const MC::LocalConcentration& c) {
  const auto s = static_cast<FloatType>(GET_CONCENTRATION(0));
  const FloatType g = s / (k + s);
  const FloatType ldot = l_dot_max * g;
  const FloatType d_length = d_t * ldot;
  GET_PROPERTY(particle_var::length) += d_length / (1.0 + d_t * ldot);  // double division, as in the reference
  GET_CONTRIBS(0) = -(phi_s_max * g);
  return check_div(GET_PROPERTY(particle_var::length), GET_PROPERTY(particle_var::l_max));
}

void _division_udf(const MC::pool_type& random_pool, std::size_t idx, std::size_t idx2,
                   const MC::DynParticlesModel<float>& arr, const MC::DynParticlesModel<float>& buffer_arr) {
  const FloatType half = GET_PROPERTY(particle_var::length) / 2.F;
  GET_PROPERTY_FROM(idx2, buffer_arr, particle_var::length) = half;
  GET_PROPERTY_FROM(idx2, buffer_arr, particle_var::l_max) = l_max_m;
  GET_PROPERTY(particle_var::length) = half;
  GET_PROPERTY(particle_var::l_max) = l_max_m;
}

double mass(std::size_t idx, const UdfModel::SelfParticle& arr) { return GET_PROPERTY(particle_var::length) * lin_density; }
}  // namespace

EXPORT_MODULE(module, &_init_udf, &_update_udf, &_division_udf, &mass, BMC_UDF_NONE, BMC_UDF_NONE, &_set_nvar, &_set_nc,
              BMC_UDF_NONE);
