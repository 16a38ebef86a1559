// Explicit instantiations of the fused step for M = 3, 4 models.
#include "step_vpsde_kernel.cuh"

namespace sdb {
template cudaError_t launch_m<3>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_m<4>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_small<3>(const StepParams&, cudaStream_t);
template cudaError_t launch_small<4>(const StepParams&, cudaStream_t);
template cudaError_t launch_and_stream<3>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_stream<4>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_smem<3>(const StepParams&, int, cudaStream_t);
template cudaError_t launch_and_smem<4>(const StepParams&, int, cudaStream_t);
}  // namespace sdb
