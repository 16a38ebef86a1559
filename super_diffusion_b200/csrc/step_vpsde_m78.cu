// Explicit instantiations of the fused step for M = 7, 8 models.
#include "step_vpsde_kernel.cuh"

namespace sdb {
template cudaError_t launch_m<7>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_m<8>(const StepParams&, int, int, int, int, cudaStream_t);
template cudaError_t launch_small<7>(const StepParams&, cudaStream_t);
template cudaError_t launch_small<8>(const StepParams&, cudaStream_t);
template cudaError_t launch_and_stream<7>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_stream<8>(const StepParams&, int, int, int, cudaStream_t);
template cudaError_t launch_and_smem<7>(const StepParams&, int, cudaStream_t);
template cudaError_t launch_and_smem<8>(const StepParams&, int, cudaStream_t);
}  // namespace sdb
