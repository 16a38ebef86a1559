// Explicit instantiations of the fused step for M = 5, 6 models.
#include "step_vpsde_kernel.cuh"

namespace sdb {
template cudaError_t launch_m<5>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_m<6>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_small<5>(const StepParams&, cudaStream_t);
template cudaError_t launch_small<6>(const StepParams&, cudaStream_t);
template cudaError_t launch_and_stream<5>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_stream<6>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_smem<5>(const StepParams&, int, cudaStream_t);
template cudaError_t launch_and_smem<6>(const StepParams&, int, cudaStream_t);
}  // namespace sdb
