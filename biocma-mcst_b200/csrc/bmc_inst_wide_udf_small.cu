// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_wide_udf_small(const std::string& var, int n_var, ModelVT& vt) {
  if (n_var == 8) return pick_variant<WideUdf<8>, 4, 4, 3>(var, vt);
  if (n_var == 16) return pick_variant<WideUdf<16>, 2, 3>(var, vt);
  return false;
}
}  // namespace bmc
