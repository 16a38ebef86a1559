// Forward-mode (JVP) companions of the score-net's non-linear ops, for the Hutchinson divergence estimator of the
// deterministic SuperDiff sampler (reference cifar/dynamics.py:72-97: jax.jvp(sdlogdx_fn, (x,), (eps,)) through
// cifar/models/ddpm.py).  Every linear layer's tangent is the same tcgen05 GEMM applied to the tangent tensor; only
// GroupNorm+swish (normalization.py:38-39 + layers.py:552,557) and the attention softmax (layers.py:507) need kernels.
#include "common.cuh"
#include "../../include/superdiff_b200.h"
#include <cuda_bf16.h>

namespace sdb {

__device__ __forceinline__ void unpack8j(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8j(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct GnJvpParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;     // primal sources (channel concat)
  const __nv_bfloat16* d0; const __nv_bfloat16* d1;     // tangent sources
  int C0, C1, B, HW, nchunk, px_per_chunk;
  const float* gamma; const float* beta;
  float eps; int apply_swish;
  float* partial;                                        // [B][nchunk][4][C]: sum x, sum x^2, sum dx, sum x*dx
  __nv_bfloat16* out; __nv_bfloat16* dout;
};

// pass 1: per (sample, pixel chunk) channel sums of x, x^2, dx, x*dx
__global__ void __launch_bounds__(256) gn_jvp_stats_kernel(const __grid_constant__ GnJvpParams p) {
  extern __shared__ float sm[];   // [rows_per_pass][4*C]
  const int C = p.C0 + p.C1, VC = C / 8;
  const int sample = blockIdx.x / p.nchunk, chunk = blockIdx.x - sample * p.nchunk;
  const int cv = threadIdx.x % VC, r = threadIdx.x / VC, rows_per_pass = blockDim.x / VC;
  const int c0 = cv * 8;
  const bool from0 = c0 < p.C0;
  const int ld = from0 ? p.C0 : p.C1;
  const size_t soff = (size_t)sample * p.HW * ld + (from0 ? c0 : c0 - p.C0);
  const __nv_bfloat16* xs = (from0 ? p.x0 : p.x1) + soff;
  const __nv_bfloat16* ds = (from0 ? p.d0 : p.d1) + soff;
  const int px0 = chunk * p.px_per_chunk, px1 = min(p.HW, px0 + p.px_per_chunk);
  float s[8], q[8], sd[8], sxd[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; sd[e] = 0.f; sxd[e] = 0.f; }
  for (int px = px0 + r; px < px1; px += rows_per_pass) {
    float f[8], g[8];
    unpack8j(*reinterpret_cast<const uint4*>(xs + (size_t)px * ld), f);
    unpack8j(*reinterpret_cast<const uint4*>(ds + (size_t)px * ld), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); sd[e] += g[e]; sxd[e] = fmaf(f[e], g[e], sxd[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float* row = sm + (size_t)r * 4 * C;
    row[c0 + e] = s[e]; row[C + c0 + e] = q[e]; row[2 * C + c0 + e] = sd[e]; row[3 * C + c0 + e] = sxd[e];
  }
  __syncthreads();
  float* dst = p.partial + ((size_t)sample * p.nchunk + chunk) * 4 * C;
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows_per_pass; ++rr) a += sm[(size_t)rr * 4 * C + i];   // fixed order
    dst[i] = a;
  }
}

// pass 2: u = xhat*gamma + beta, du = rstd*gamma*(dx - mean(dx) - xhat*mean(xhat*dx));  a = swish(u), da = swish'(u)*du
__global__ void __launch_bounds__(256) gn_jvp_apply_kernel(const __grid_constant__ GnJvpParams p) {
  __shared__ float ch_tot[4 * 1024];
  __shared__ float g_stat[4 * 32];      // mean, rstd, m1 = mean(dx), m2 = mean(xhat*dx)
  const int C = p.C0 + p.C1, VC = C / 8, cpg = C / 32;
  const int sample = blockIdx.x / p.nchunk, chunk = blockIdx.x - sample * p.nchunk;
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    const float* base = p.partial + (size_t)sample * p.nchunk * 4 * C + i;
    float a = 0.f;
    for (int k = 0; k < p.nchunk; ++k) a += base[(size_t)k * 4 * C];
    ch_tot[i] = a;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double sx = 0.0, sq = 0.0, sd = 0.0, sxd = 0.0;
    for (int c = 0; c < cpg; ++c) {
      const int ch = threadIdx.x * cpg + c;
      sx += (double)ch_tot[ch]; sq += (double)ch_tot[C + ch]; sd += (double)ch_tot[2 * C + ch]; sxd += (double)ch_tot[3 * C + ch];
    }
    const double n = (double)p.HW * cpg;
    const double mean = sx / n;
    double var = sq / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)p.eps);
    g_stat[threadIdx.x * 4] = (float)mean;
    g_stat[threadIdx.x * 4 + 1] = (float)rstd;
    g_stat[threadIdx.x * 4 + 2] = (float)(sd / n);
    g_stat[threadIdx.x * 4 + 3] = (float)(rstd * (sxd - mean * sd) / n);
  }
  __syncthreads();
  const int cv = threadIdx.x % VC, r = threadIdx.x / VC, rows_per_pass = blockDim.x / VC;
  const int c0 = cv * 8;
  const bool from0 = c0 < p.C0;
  const int ld = from0 ? p.C0 : p.C1;
  const size_t soff = (size_t)sample * p.HW * ld + (from0 ? c0 : c0 - p.C0);
  const __nv_bfloat16* xs = (from0 ? p.x0 : p.x1) + soff;
  const __nv_bfloat16* ds = (from0 ? p.d0 : p.d1) + soff;
  float mu[8], rs[8], m1[8], m2[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e, g = c / cpg;
    mu[e] = g_stat[g * 4]; rs[e] = g_stat[g * 4 + 1]; m1[e] = g_stat[g * 4 + 2]; m2[e] = g_stat[g * 4 + 3];
    ga[e] = p.gamma[c]; be[e] = p.beta[c];
  }
  __nv_bfloat16* dst = p.out + (size_t)sample * p.HW * C + c0;
  __nv_bfloat16* ddst = p.dout + (size_t)sample * p.HW * C + c0;
  const int px0 = chunk * p.px_per_chunk, px1 = min(p.HW, px0 + p.px_per_chunk);
  for (int px = px0 + r; px < px1; px += rows_per_pass) {
    float f[8], g[8];
    unpack8j(*reinterpret_cast<const uint4*>(xs + (size_t)px * ld), f);
    unpack8j(*reinterpret_cast<const uint4*>(ds + (size_t)px * ld), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (f[e] - mu[e]) * rs[e];
      const float u = fmaf(xh, ga[e], be[e]);
      const float du = rs[e] * ga[e] * (g[e] - m1[e] - xh * m2[e]);
      if (p.apply_swish) {
        const float sg = 1.f / (1.f + __expf(-u));
        f[e] = u * sg;
        g[e] = du * sg * (1.f + u * (1.f - sg));
      } else {
        f[e] = u;
        g[e] = du;
      }
    }
    *reinterpret_cast<uint4*>(dst + (size_t)px * C) = pack8j(f);
    *reinterpret_cast<uint4*>(ddst + (size_t)px * C) = pack8j(g);
  }
}

// dP = P * (dS - sum_k P_k dS_k), dS = scale * (dS1 + dS2)  -- JVP of a row softmax given its output P
__global__ void __launch_bounds__(256) softmax_jvp_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dS1,
                                                           const float* __restrict__ dS2, float scale,
                                                           __nv_bfloat16* __restrict__ dP, long rows, int cols) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* pr = P + row * cols;
  const float* a = dS1 + row * cols;
  const float* b = dS2 ? dS2 + row * cols : nullptr;
  float acc = 0.f;
  for (int j = lane; j < cols; j += 32) {
    const float ds = scale * (a[j] + (b ? b[j] : 0.f));
    acc = fmaf(__bfloat162float(pr[j]), ds, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  for (int j = lane; j < cols; j += 32) {
    const float ds = scale * (a[j] + (b ? b[j] : 0.f));
    dP[row * cols + j] = __float2bfloat16_rn(__bfloat162float(pr[j]) * (ds - acc));
  }
}

// out[b * out_stride] = scale * <a[b,:], b[b,:]>   (fp32 inputs, fp64 reduction)
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int D, float scale,
                                                      float* __restrict__ out, int out_stride) {
  __shared__ double red[8];
  const size_t base = (size_t)blockIdx.x * D;
  double acc = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) acc += (double)a[base + i] * (double)b[base + i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    out[(size_t)blockIdx.x * out_stride] = (float)(t * (double)scale);
  }
}

}  // namespace sdb

extern "C" {

int sd_groupnorm_swish_jvp(const void* x0, const void* dx0, int C0, const void* x1, const void* dx1, int C1, int B, int HW,
                           const float* gamma, const float* beta, float eps, int apply_swish, float* scratch,
                           size_t scratch_floats, void* out, void* dout, void* stream) {
  using namespace sdb;
  if (!x0 || !dx0 || !gamma || !beta || !out || !dout || !scratch || (C1 > 0 && (!x1 || !dx1)))
    return fail(kErrInvalidArg, "sd_groupnorm_swish_jvp: null pointer");
  if (C1 < 0) C1 = 0;
  const int C = C0 + C1;
  if (C0 < 8 || C0 % 8 || C1 % 8 || C % 32 || C > 1024) return fail(kErrInvalidArg, "sd_groupnorm_swish_jvp: channels must be multiples of 8, total a multiple of 32, <= 1024");
  if (B < 0 || HW < 1) return fail(kErrInvalidArg, "sd_groupnorm_swish_jvp: bad shape");
  if (B == 0) return SD_OK;
  const int VC = C / 8;
  int k = 256 / VC;
  if (k < 1) k = 1;
  while (k > 1 && (VC * k) % 32) --k;
  const int T = VC * k;
  if (T % 32 || T > 256) return fail(kErrUnsupported, "sd_groupnorm_swish_jvp: unsupported channel count");
  GnJvpParams p{};
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1; p.d0 = (const __nv_bfloat16*)dx0; p.d1 = (const __nv_bfloat16*)dx1;
  p.C0 = C0; p.C1 = C1; p.B = B; p.HW = HW; p.gamma = gamma; p.beta = beta; p.eps = eps; p.apply_swish = apply_swish;
  p.out = (__nv_bfloat16*)out; p.dout = (__nv_bfloat16*)dout; p.partial = scratch;
  int nchunk = (148 * 4 + B - 1) / B;
  const int max_chunks = (HW + k - 1) / k;
  if (nchunk > max_chunks) nchunk = max_chunks;
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 64) nchunk = 64;
  p.px_per_chunk = (HW + nchunk - 1) / nchunk;
  p.nchunk = (HW + p.px_per_chunk - 1) / p.px_per_chunk;
  if ((size_t)B * p.nchunk * 4 * C > scratch_floats) return fail(kErrInvalidArg, "sd_groupnorm_swish_jvp: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  gn_jvp_stats_kernel<<<(unsigned)(B * p.nchunk), T, sizeof(float) * (size_t)k * 4 * C, st>>>(p);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return check_cuda(err, "sd_groupnorm_swish_jvp (stats) launch");
  gn_jvp_apply_kernel<<<(unsigned)(B * p.nchunk), T, 0, st>>>(p);
  return check_cuda(cudaGetLastError(), "sd_groupnorm_swish_jvp (apply) launch");
}

int sd_softmax_jvp(const void* P, const float* dS1, const float* dS2, float scale, void* dP, long rows, int cols, void* stream) {
  using namespace sdb;
  if (!P || !dS1 || !dP || rows < 0 || cols < 1) return fail(kErrInvalidArg, "sd_softmax_jvp: bad argument");
  if (rows == 0) return SD_OK;
  const long blocks = (rows + 7) / 8;
  softmax_jvp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)P, dS1, dS2, scale, (__nv_bfloat16*)dP, rows, cols);
  return check_cuda(cudaGetLastError(), "sd_softmax_jvp launch");
}

int sd_rowdot(const float* a, const float* b, int B, int D, float scale, float* out, int out_stride, void* stream) {
  using namespace sdb;
  if (!a || !b || !out || B < 0 || D < 1 || out_stride < 1) return fail(kErrInvalidArg, "sd_rowdot: bad argument");
  if (B == 0) return SD_OK;
  rowdot_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(a, b, D, scale, out, out_stride);
  return check_cuda(cudaGetLastError(), "sd_rowdot launch");
}

}  // extern "C"
