// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_simple_acetate(const std::string& var, ModelVT& vt) { return pick_variant<SimpleAcetate, 4, 3, 4>(var, vt); }
}  // namespace bmc
