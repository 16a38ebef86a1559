// Shared device/host helpers for the superdiff_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <utility>

namespace sdb {

namespace cg = cooperative_groups;

// ---- error plumbing for the C ABI (no exceptions cross the boundary) --------
void set_last_error(const std::string& msg);
int fail(int code, const std::string& msg);
int check_cuda(cudaError_t e, const char* what);

constexpr int kErrInvalidArg = -1;
constexpr int kErrUnsupported = -2;
constexpr int kErrCuda = -3;

// Kernel function attributes (cudaFuncAttributeMaxDynamicSharedMemorySize ...) belong to a device's context: a process that drives
// several GPUs has to set them once per device, not once per process.
struct PerDeviceOnce {
  std::mutex mu;
  bool done[64] = {};
  cudaError_t err[64] = {};
  template <typename F>
  cudaError_t run(F&& fn) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(mu);
    if (!done[dev]) { err[dev] = fn(); done[dev] = true; }
    return err[dev];
  }
};

// ---- streaming 128-bit global accesses ---------------------------------------
// Inputs of the fused step are read exactly once: keep them out of L1.
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st4(float* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- programmatic dependent launch ---------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its predecessor on the
// stream is still draining: everything before pdl_wait() (barrier init, TMEM allocation, descriptor prefetch, smem
// carve-up) overlaps the predecessor's tail; pdl_wait() returns once the predecessor grid has completed and its
// memory is visible.  pdl_launch_dependents() lets the *next* kernel begin the same way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();   // SDB_PDL=1 (default off, see capi.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (pdl_enabled()) { cfg.attrs = attr; cfg.numAttrs = 1; }
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// 1 / d in fp64 without the division subroutine: fp32 reciprocal as the seed, two Newton steps (4 DFMAs) -> ~1 ulp.  fp64
// division / rsqrt compile to CALLs into slow-path routines whose latency (microseconds when they sit on a CTA's critical
// path: profiles/r02_notes.md) dwarfs the arithmetic they replace.  Falls back to the true division outside fp32's range.
__device__ __forceinline__ double fast_drcp(double d) {
  const double ad = fabs(d);
  if (!(ad > 1e-30 && ad < 1e30)) return 1.0 / d;
  double x = (double)(1.0f / (float)d);
  x = x * fma(-d, x, 2.0);
  x = x * fma(-d, x, 2.0);
  return x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp totals of K per-lane values by butterfly REDUCE-SCATTER: in the round with lane offset o a lane keeps one half of
// the values it is still responsible for and hands the other half to its partner, so a group of 32 values costs
// 16+8+4+2+1 = 31 exchanges instead of the 5*32 of one all-reduce butterfly per value, and lane l ends with the total of
// value l.  (K = 44 for AND with 8 models: 440 -> 116 SHFL per warp; at one SHFL per clock per SM the all-reduce form
// alone cost 1.9 us per CTA, more than moving the sample.)  A group with R < 32 values first runs plain butterflies for
// the offsets >= R's power-of-two ceiling.  Order of the fp64 additions is fixed: results are deterministic.
template <int R>
__device__ __forceinline__ void warp_reduce_scatter_group(double (&v)[32], int lane) {
  constexpr int N0 = R > 16 ? 32 : R > 8 ? 16 : R > 4 ? 8 : R > 2 ? 4 : R > 1 ? 2 : 1;
#pragma unroll
  for (int o = 16; o >= N0 && o >= 1; o >>= 1) {     // values replicated over the high lane bits
#pragma unroll
    for (int i = 0; i < R; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
#pragma unroll
  for (int o = N0 / 2; o >= 1; o >>= 1) {
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const double send = upper ? v[i] : v[i + o];
      const double keep = upper ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

// dst[k] = sum over the warp of part[k], k < K (dst: K doubles, written by the lanes that end up owning each value)
template <int K>
__device__ __forceinline__ void warp_reduce_scatter_store(const float (&part)[K], double* dst, int lane) {
  constexpr int G = (K + 31) / 32;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (g * 32 + i < K) ? (double)part[(g * 32 + i < K) ? g * 32 + i : 0] : 0.0;
    if (g < G - 1 || K % 32 == 0) {
      warp_reduce_scatter_group<32>(v, lane);
      dst[g * 32 + lane] = v[0];
    } else {
      constexpr int R = K % 32 == 0 ? 32 : K % 32;
      constexpr int N0 = R > 16 ? 32 : R > 8 ? 16 : R > 4 ? 8 : R > 2 ? 4 : R > 1 ? 2 : 1;
      warp_reduce_scatter_group<R>(v, lane);
      if (lane < R && lane < N0) dst[g * 32 + lane] = v[0];
    }
  }
}

// Reduce K per-thread fp32 partials over the CTA and (optionally) over the
// thread-block cluster that shares one sample.  Partials are widened to fp64
// before the first cross-thread add, so the result does not depend on the
// CTA/cluster shape beyond fp64 rounding.  Returns a shared-memory pointer to
// the K totals, valid for every thread of every CTA in the cluster.
//   scratch: shared double[ (nwarps + 2) * K ]
template <int K, bool CLUSTER>
__device__ __forceinline__ const double* block_cluster_sum(const float (&part)[K], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  double* cta_tot = scratch + nwarps * K;
  double* full = cta_tot + K;
  warp_reduce_scatter_store<K>(part, scratch + warp * K, lane);
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += scratch[w * K + threadIdx.x];
    cta_tot[threadIdx.x] = s;
    if (!CLUSTER) full[threadIdx.x] = s;
  }
  if (CLUSTER) {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();  // every CTA's cta_tot is written and visible cluster-wide
    if (threadIdx.x < K) {
      double s = 0.0;
      const unsigned n = cluster.num_blocks();
      for (unsigned r = 0; r < n; ++r) s += *cluster.map_shared_rank(&cta_tot[threadIdx.x], r);
      full[threadIdx.x] = s;
    }
    cluster.sync();  // nobody exits (or reuses cta_tot) while peers still read it
  } else {
    __syncthreads();
  }
  return full;
}

}  // namespace sdb
