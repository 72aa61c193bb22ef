// Instantiations of the fused depth-step kernel (csrc/step_fused_kernel.cuh); see csrc/step_fused.cu.
#include "step_fused_kernel.cuh"

namespace dmpnn {
namespace fused {
template cudaError_t dispatch_bwd<true>(int, int, int, cudaStream_t, const CUtensorMap&, const CUtensorMap&, const Params&);
}  // namespace fused
}  // namespace dmpnn
