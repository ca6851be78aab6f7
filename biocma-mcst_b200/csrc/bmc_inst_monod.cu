// Kernel instantiations of one model (own translation unit: the models build in parallel).
#include "bmc_model_vt.cuh"

namespace bmc {
bool pick_monod(const std::string& var, ModelVT& vt) { return pick_variant<Monod, 4, 4, 3>(var, vt); }  // 1024 threads x 64 registers (v4b3: 768 x 80)
}  // namespace bmc
