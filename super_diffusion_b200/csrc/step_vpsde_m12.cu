// Explicit instantiations of the fused step for M = 1, 2 models.
#include "step_vpsde_kernel.cuh"

namespace sdb {
template cudaError_t launch_m<1>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_m<2>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_small<1>(const StepParams&, cudaStream_t);
template cudaError_t launch_small<2>(const StepParams&, cudaStream_t);
template cudaError_t launch_and_stream<1>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_stream<2>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_smem<1>(const StepParams&, int, cudaStream_t);
template cudaError_t launch_and_smem<2>(const StepParams&, int, cudaStream_t);
}  // namespace sdb
