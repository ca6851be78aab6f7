// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_monod(const std::string& var, bool large, ModelVT& vt) {
  return pick_variant<Monod, 4>(var, large ? 3 : 4, vt);  // 1024 threads x 64 registers / 768 x 80 (see kLargePopulation)
}
}  // namespace bmc
