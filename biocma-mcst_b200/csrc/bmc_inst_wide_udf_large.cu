// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_wide_udf_large(const std::string& var, int n_var, ModelVT& vt) {
  if (n_var == 32) return pick_variant<WideUdf<32>, 1, 3>(var, vt);
  if (n_var == 64) return pick_variant<WideUdf<64>, 1, 2>(var, vt);
  return false;
}
}  // namespace bmc
