// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_monod(const std::string& var, ModelVT& vt) {
  if (var.size() == 4 && var[1] == '2') return pick_variant<Monod, 2, true>(var, 4, vt);
  return pick_variant<Monod, 4, true>(var, 4, vt);  // 64 registers -> 4 blocks (32 warps) per SM
}
}  // namespace bmc
