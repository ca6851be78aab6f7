// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_wide_udf_small(const std::string& var, bool large, int n_var, ModelVT& vt) {
  if (n_var == 8) return pick_variant<WideUdf<8>, 4>(var, large ? 3 : 4, vt);
  if (n_var == 16) return pick_variant<WideUdf<16>, 2>(var, 3, vt);
  return false;
}
}  // namespace bmc
