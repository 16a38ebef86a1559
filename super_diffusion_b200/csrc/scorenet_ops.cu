// Memory-bound and small ops of the CIFAR score-net forward (sm_100a), NHWC bf16 activations.
// Reference call sites: cifar/models/ddpm.py:47-101, cifar/models/layers.py:450-565,
// cifar/models/normalization.py:38-39 (flax nn.GroupNorm defaults).
#include "common.cuh"
#include "../../include/superdiff_b200.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace sdb {

__device__ __forceinline__ float swish_f(float v) { return v / (1.f + __expf(-v)); }

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<const uint32_t*>(&h2);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------
// GroupNorm(32) + swish over concat(x0, x1) along channels.
// Groups are independent, so the work is split over (sample, channel block): one CTA owns all H*W pixels of
// `Cs` consecutive channels (a whole number of groups), keeps them in registers (read once), reduces per-channel
// partials through shared memory in a fixed order (deterministic) and writes the normalised block.  No clusters,
// no cross-CTA traffic (measured on B200: cluster launches of memory-bound kernels cost 2-3x).
// ---------------------------------------------------------------------------
struct GnParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int C0, C1, B, HW;
  int Cs, nsplit;            // channels per CTA, CTAs per sample
  const float* gamma; const float* beta;
  float eps; int apply_swish;
  __nv_bfloat16* out;
};

template <int NV>
__global__ void __launch_bounds__(512) groupnorm_swish_kernel(const __grid_constant__ GnParams p) {
  extern __shared__ float gn_smem[];   // [2*Cs] channel sums / sumsq, [2*32] group mean / rstd, [rows][2*Cs] partials
  const int C = p.C0 + p.C1, Cs = p.Cs;
  const int VC = Cs / 8;               // 16-byte vectors per pixel owned by this CTA
  float* ch_sum = gn_smem;
  float* ch_sq = gn_smem + Cs;
  float* g_mean = gn_smem + 2 * Cs;
  float* g_rstd = g_mean + 32;
  float* part = g_rstd + 32;           // [rows_per_pass][2*Cs]
  const int sample = blockIdx.x / p.nsplit;
  const int c_base = (blockIdx.x - sample * p.nsplit) * Cs;
  const int cv = threadIdx.x % VC, r = threadIdx.x / VC, rows_per_pass = blockDim.x / VC;
  const int c0 = c_base + cv * 8;      // first of this thread's 8 channels (never straddles x0 | x1: C0 % 8 == 0)

  const bool from0 = c0 < p.C0;
  const __nv_bfloat16* src = from0 ? p.x0 + (size_t)sample * p.HW * p.C0 + c0
                                   : p.x1 + (size_t)sample * p.HW * p.C1 + (c0 - p.C0);
  const int src_ld = from0 ? p.C0 : p.C1;
  uint4 v[NV];
  bool ok[NV];
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int px = r + j * rows_per_pass;
    ok[j] = px < p.HW;
    if (ok[j]) v[j] = *reinterpret_cast<const uint4*>(src + (size_t)px * src_ld);
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (!ok[j]) continue;
    float f[8];
    unpack8(v[j], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
  // fixed-order (deterministic) reduction over the pixel rows
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    part[(size_t)r * 2 * Cs + cv * 8 + e] = s[e];
    part[(size_t)r * 2 * Cs + Cs + cv * 8 + e] = q[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cs; i += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows_per_pass; ++rr) a += part[(size_t)rr * 2 * Cs + i];
    gn_smem[i] = a;
  }
  __syncthreads();
  const int cpg = C / 32;
  const int ngroups = Cs / cpg;
  if (threadIdx.x < ngroups) {
    double sum = 0.0, sq = 0.0;
    for (int c = 0; c < cpg; ++c) { sum += (double)ch_sum[threadIdx.x * cpg + c]; sq += (double)ch_sq[threadIdx.x * cpg + c]; }
    const double n = (double)p.HW * cpg;
    const double mean = sum / n;
    double var = sq / n - mean * mean;      // flax: E[x^2] - E[x]^2, clipped at 0
    if (var < 0.0) var = 0.0;
    g_mean[threadIdx.x] = (float)mean;
    g_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)p.eps));
  }
  __syncthreads();
  float mu[8], rs[8], ga[8], be[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cl = cv * 8 + e;
    mu[e] = g_mean[cl / cpg];
    rs[e] = g_rstd[cl / cpg];
    ga[e] = p.gamma[c_base + cl];
    be[e] = p.beta[c_base + cl];
  }
  __nv_bfloat16* dst = p.out + (size_t)sample * p.HW * C + c0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (!ok[j]) continue;
    float f[8];
    unpack8(v[j], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float y = (f[e] - mu[e]) * rs[e] * ga[e] + be[e];
      f[e] = p.apply_swish ? swish_f(y) : y;
    }
    const int px = r + j * rows_per_pass;
    *reinterpret_cast<uint4*>(dst + (size_t)px * C) = pack8(f);
  }
}

template <int NV>
static cudaError_t launch_gn(const GnParams& p, int threads, cudaStream_t st) {
  const int rows = threads / (p.Cs / 8);
  const size_t smem = sizeof(float) * ((size_t)2 * p.Cs * (1 + rows) + 64);
  groupnorm_swish_kernel<NV><<<(unsigned)(p.B * p.nsplit), threads, smem, st>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Small-S attention (S <= 64 tokens, C = 256): one CTA per sample, fp32 math.
// ---------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) attention_small_kernel(const __nv_bfloat16* __restrict__ qkv, int C,
                                                              __nv_bfloat16* __restrict__ out) {
  extern __shared__ float att_smem[];
  // layout: q [S][C+1], k [S][C+1], v [S][C], w [S][S+1]
  float* sq = att_smem;
  float* sk = sq + S * (C + 1);
  float* sv = sk + S * (C + 1);
  float* sw = sv + S * C;
  const int b = blockIdx.x;
  const __nv_bfloat16* base = qkv + (size_t)b * S * 3 * C;
  for (int i = threadIdx.x; i < S * 3 * C / 8; i += blockDim.x) {
    const int row = i / (3 * C / 8), col = (i % (3 * C / 8)) * 8;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(base + (size_t)row * 3 * C + col), f);
    float* dst = col < C ? sq + row * (C + 1) + col : (col < 2 * C ? sk + row * (C + 1) + (col - C) : sv + row * C + (col - 2 * C));
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[e] = f[e];
  }
  __syncthreads();
  const float scale = rsqrtf((float)C);
  for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
    const int qi = i / S, kj = i % S;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(sq[qi * (C + 1) + c], sk[kj * (C + 1) + c], acc);
    sw[qi * (S + 1) + kj] = acc * scale;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = warp; qi < S; qi += blockDim.x / 32) {
    float mx = -INFINITY;
    for (int j = lane; j < S; j += 32) mx = fmaxf(mx, sw[qi * (S + 1) + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) { const float e = expf(sw[qi * (S + 1) + j] - mx); sw[qi * (S + 1) + j] = e; sum += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int j = lane; j < S; j += 32) sw[qi * (S + 1) + j] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < S; ++j) acc = fmaf(sw[qi * (S + 1) + j], sv[j * C + c], acc);
    out[(size_t)b * S * C + i] = __float2bfloat16_rn(acc);
  }
}

// Row softmax: P[r, :] = softmax(scale * X[r, :]), fp32 in, bf16 out; one warp per row.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           long rows, int cols, float scale) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * cols;
  float mx = -INFINITY;
  for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, xr[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < cols; j += 32) sum += expf(xr[j] * scale - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int j = lane; j < cols; j += 32) out[row * cols + j] = __float2bfloat16_rn(expf(xr[j] * scale - mx) * inv);
}

// Nearest x2 upsample, 16-byte vectors.
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                         int B, int H, int W, int VC) {
  const size_t total = (size_t)B * 4 * H * W * VC;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = i % VC;
    size_t t = i / VC;
    const int wo = t % (2 * W); t /= (2 * W);
    const int ho = t % (2 * H);
    const int b = t / (2 * H);
    out[i] = x[(((size_t)b * H + ho / 2) * W + wo / 2) * VC + cv];
  }
}

// Gather for the stride-2 SAME conv (pad (0,1)): out[b,ho,wo,tap,c] = x[b,2ho+kh,2wo+kw,c].
__global__ void __launch_bounds__(256) im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out,
                                                        int B, int H, int W, int VC) {
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)B * Ho * Wo * 9 * VC;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cv = i % VC;
    size_t t = i / VC;
    const int tap = t % 9; t /= 9;
    const int wo = t % Wo; t /= Wo;
    const int ho = t % Ho;
    const int b = t / Ho;
    const int hi = 2 * ho + tap / 3, wi = 2 * wo + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hi < H && wi < W) v = x[(((size_t)b * H + hi) * W + wi) * VC + cv];
    out[i] = v;
  }
}

// First conv: fp32 NHWC [B,H,W,Cin<=4] -> bf16 [B,H,W,Cout].  Weights (9*Cin x Cout fp32, <= 18 KB) sit in
// shared memory; a thread owns one pixel x 32 output channels (its 9*Cin inputs in registers) and writes
// 64 contiguous bytes, so a pixel's 256-byte row is covered by 4 adjacent threads (coalesced).
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, int B, int H, int W, int Cin,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      int Cout, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float cin_smem[];   // [9*Cin][Cout] weights, then [Cout] bias
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) cin_smem[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) cin_smem[K * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout / 32;
  const size_t total = (size_t)B * H * W * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg0 = (int)(i % groups) * 32;
    size_t pix = i / groups;
    const int wo = pix % W; pix /= W;
    const int ho = pix % H;
    const int b = (int)(pix / H);
    float in[36];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int hi = ho + kh - 1, wi = wo + kw - 1;
        const bool okp = hi >= 0 && hi < H && wi >= 0 && wi < W;
        for (int c = 0; c < Cin; ++c)
          in[(kh * 3 + kw) * 4 + c] = okp ? x[(((size_t)b * H + hi) * W + wi) * Cin + c] : 0.f;
      }
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = cin_smem[K * Cout + cg0 + j];
#pragma unroll
    for (int t = 0; t < 9; ++t)
      for (int c = 0; c < Cin; ++c) {
        const float v = in[t * 4 + c];
        const float4* wr = reinterpret_cast<const float4*>(cin_smem + (size_t)(t * Cin + c) * Cout + cg0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 w4 = wr[j];
          acc[4 * j] = fmaf(v, w4.x, acc[4 * j]); acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
        }
      }
    uint4* op = reinterpret_cast<uint4*>(out + ((((size_t)b * H + ho) * W + wo) * Cout + cg0));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = acc[8 * j + e];
      op[j] = pack8(f);
    }
  }
}

// temb = Dense1(swish(Dense0(sinusoidal(t))))  for n_t distinct times -> fp32 scratch [n_t, 4nf]
__global__ void __launch_bounds__(512) temb_dense_kernel(const float* __restrict__ t_dev, int t_stride,
                                                         const float* __restrict__ sched, const int* __restrict__ step_counter,
                                                         int nf, const float* __restrict__ w0, const float* __restrict__ b0,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         float* __restrict__ temb) {
  extern __shared__ float te_smem[];    // emb[nf], h1[4nf]
  float* emb = te_smem;
  float* h1 = te_smem + nf;
  const int i = blockIdx.x, nh = 4 * nf;
  float t;
  if (sched) t = sched[4 * (size_t)(step_counter ? *step_counter : 0) + 2];   // sigma column == t (sigma_t = t)
  else t = t_dev[(size_t)i * t_stride];
  const int half = nf / 2;
  const float lf = logf(10000.f) / (float)(half - 1);
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float arg = t * expf((float)k * -lf);
    emb[k] = sinf(arg);
    emb[half + k] = cosf(arg);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nh; j += blockDim.x) {
    float acc = b0[j];
    for (int k = 0; k < nf; ++k) acc = fmaf(emb[k], w0[(size_t)k * nh + j], acc);
    h1[j] = swish_f(acc);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < nh; j += blockDim.x) {
    float acc = b1[j];
    for (int k = 0; k < nh; ++k) acc = fmaf(h1[k], w1[(size_t)k * nh + j], acc);
    temb[(size_t)i * nh + j] = acc;
  }
}

// act_temb[b, :] = swish(temb[b or 0, :] + class_emb[label_b, :])  -> bf16
__global__ void __launch_bounds__(256) temb_act_kernel(const float* __restrict__ temb, int temb_stride,
                                                       const float* __restrict__ class_emb, const int* __restrict__ labels,
                                                       int B, int nh, __nv_bfloat16* __restrict__ out) {
  const size_t total = (size_t)B * nh;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = i / nh, j = i % nh;
    float v = temb[(size_t)b * temb_stride + j];
    if (class_emb) v += class_emb[(size_t)labels[b] * nh + j];
    out[i] = __float2bfloat16_rn(swish_f(v));
  }
}

static unsigned grid_for(size_t total, int threads) {
  const size_t want = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace sdb

extern "C" {

int sd_groupnorm_swish(const void* x0, int C0, const void* x1, int C1, int B, int HW, const float* gamma,
                       const float* beta, float eps, int apply_swish, void* out, void* stream) {
  using namespace sdb;
  if (!x0 || !gamma || !beta || !out || (C1 > 0 && !x1)) return fail(kErrInvalidArg, "sd_groupnorm_swish: null pointer");
  if (C1 < 0) C1 = 0;
  const int C = C0 + C1;
  if (C0 < 8 || C0 % 8 || C1 % 8 || C % 32 || C > 2048) return fail(kErrInvalidArg, "sd_groupnorm_swish: channels must be multiples of 8, total a multiple of 32");
  if (B < 0 || HW < 1) return fail(kErrInvalidArg, "sd_groupnorm_swish: bad shape");
  if (B == 0) return SD_OK;
  GnParams p{(const __nv_bfloat16*)x0, (const __nv_bfloat16*)x1, C0, C1, B, HW, 0, 0, gamma, beta, eps, apply_swish,
             (__nv_bfloat16*)out};
  const int cpg = C / 32;
  // channel block per CTA: a multiple of lcm(cpg, 8) channels dividing C; threads = (Cs/8) * k pixel rows per pass.
  // Prefer <= 8 vectors per thread (3+ CTAs resident per SM), then wide channel blocks (longer contiguous segments).
  int unit = cpg;
  while (unit % 8) unit += cpg;
  static const int min_cs = [] { const char* e = getenv("SDB_GN_MIN_CS"); return e ? atoi(e) : 32; }();   // tuning knob
  int best_cs = 0, best_t = 0, best_nv = 0;
  long best_cost = -1;
  for (int Cs = unit; Cs <= C; Cs += unit) {
    if (C % Cs) continue;
    const int VC = Cs / 8;
    for (int k = 1; VC * k <= 512; ++k) {
      const int T = VC * k;
      if (T % 32) continue;
      const int need = (HW + k - 1) / k;
      if (need > 16) continue;
      const int nvt = need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : need <= 8 ? 8 : 16;
      const size_t smem = sizeof(float) * ((size_t)2 * Cs * (1 + k) + 64);
      if (smem > 96 * 1024) continue;
      long cost = (nvt > 8 ? 4000 : 0) + (T < 128 ? 1500 : 0) + (T > 384 ? 300 : 0) + (long)(nvt * k - HW) * 4  // idle lanes
                  + 2048 / Cs + (Cs < min_cs ? 1000 : 0) + (smem > 32 * 1024 ? 500 : 0);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_cs = Cs; best_t = T; best_nv = need; }
    }
  }
  if (best_cost < 0) return fail(kErrUnsupported, "sd_groupnorm_swish: H*W too large for the register-resident path (max 8192 pixels)");
  p.Cs = best_cs;
  p.nsplit = C / best_cs;
  const int nv = best_nv, T = best_t;
  cudaError_t err;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(groupnorm_swish_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(groupnorm_swish_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(groupnorm_swish_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(groupnorm_swish_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(groupnorm_swish_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_done = true;
  }
  if (nv <= 1) err = launch_gn<1>(p, T, st);
  else if (nv <= 2) err = launch_gn<2>(p, T, st);
  else if (nv <= 4) err = launch_gn<4>(p, T, st);
  else if (nv <= 8) err = launch_gn<8>(p, T, st);
  else err = launch_gn<16>(p, T, st);
  return check_cuda(err, "sd_groupnorm_swish launch");
}

int sd_attention(const void* qkv, int B, int S, int C, void* out, void* stream) {
  using namespace sdb;
  if (!qkv || !out) return fail(kErrInvalidArg, "sd_attention: null pointer");
  if (C % 8 || C < 8) return fail(kErrInvalidArg, "sd_attention: C must be a multiple of 8");
  if (B == 0) return SD_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t err;
#define SDB_ATT(Sv)                                                                                        \
  {                                                                                                        \
    const size_t smem = sizeof(float) * ((size_t)2 * Sv * (C + 1) + (size_t)Sv * C + (size_t)Sv * (Sv + 1)); \
    if (smem > 227 * 1024) return fail(kErrUnsupported, "sd_attention: S*C too large for the small-S kernel");  \
    err = cudaFuncSetAttribute(attention_small_kernel<Sv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (err == cudaSuccess) {                                                                              \
      attention_small_kernel<Sv><<<B, 256, smem, st>>>((const __nv_bfloat16*)qkv, C, (__nv_bfloat16*)out); \
      err = cudaGetLastError();                                                                            \
    }                                                                                                      \
  }
  if (S == 16) SDB_ATT(16)
  else if (S == 64) SDB_ATT(64)
  else if (S == 4) SDB_ATT(4)
  else return fail(kErrUnsupported, "sd_attention: small-S kernel supports S in {4, 16, 64}; larger S uses the batched tcgen05 GEMM path");
#undef SDB_ATT
  return check_cuda(err, "sd_attention launch");
}

int sd_softmax_rows(const float* x, void* out, long rows, int cols, float scale, void* stream) {
  using namespace sdb;
  if (!x || !out || rows < 0 || cols < 1) return fail(kErrInvalidArg, "sd_softmax_rows: bad argument");
  if (rows == 0) return SD_OK;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  softmax_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, rows, cols, scale);
  return check_cuda(cudaGetLastError(), "sd_softmax_rows launch");
}

int sd_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream) {
  using namespace sdb;
  if (!x || !out || C % 8) return fail(kErrInvalidArg, "sd_upsample2x: bad argument (C must be a multiple of 8)");
  if (B == 0) return SD_OK;
  const size_t total = (size_t)B * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, B, H, W, C / 8);
  return check_cuda(cudaGetLastError(), "sd_upsample2x launch");
}

int sd_im2col_s2(const void* x, int B, int H, int W, int C, void* out, void* stream) {
  using namespace sdb;
  if (!x || !out || C % 8 || H % 2 || W % 2) return fail(kErrInvalidArg, "sd_im2col_s2: bad argument");
  if (B == 0) return SD_OK;
  const size_t total = (size_t)B * (H / 2) * (W / 2) * 9 * (C / 8);
  im2col_s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)out, B, H, W, C / 8);
  return check_cuda(cudaGetLastError(), "sd_im2col_s2 launch");
}

int sd_conv_in(const float* x, int B, int H, int W, int Cin, const float* w_hwio, const float* bias, int Cout,
               void* out, void* stream) {
  using namespace sdb;
  if (!x || !w_hwio || !out || Cin < 1 || Cin > 4) return fail(kErrInvalidArg, "sd_conv_in: Cin must be in [1, 4]");
  if (Cout < 32 || Cout % 32 || Cout > 512) return fail(kErrInvalidArg, "sd_conv_in: Cout must be a multiple of 32, <= 512");
  if (B == 0) return SD_OK;
  const size_t smem = sizeof(float) * ((size_t)9 * Cin * Cout + Cout);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_in_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return check_cuda(e, "sd_conv_in: cudaFuncSetAttribute");
    attr_done = true;
  }
  const size_t total = (size_t)B * H * W * (Cout / 32);
  conv_in_kernel<<<grid_for(total, 256), 256, smem, (cudaStream_t)stream>>>(x, B, H, W, Cin, w_hwio, bias, Cout,
                                                                             (__nv_bfloat16*)out);
  return check_cuda(cudaGetLastError(), "sd_conv_in launch");
}

int sd_time_embedding(const float* t_dev, int t_stride, const float* sched, const int* step_counter, int B, int nf,
                      const float* w0, const float* b0, const float* w1, const float* b1, const float* class_emb,
                      const int* labels, float* temb_scratch, void* act_temb_out, void* stream) {
  using namespace sdb;
  if ((!t_dev && !sched) || !w0 || !b0 || !w1 || !b1 || !temb_scratch || !act_temb_out)
    return fail(kErrInvalidArg, "sd_time_embedding: null pointer");
  if (class_emb && !labels) return fail(kErrInvalidArg, "sd_time_embedding: class embedding needs labels");
  if (nf < 4 || nf % 2) return fail(kErrInvalidArg, "sd_time_embedding: nf must be even and >= 4");
  if (B == 0) return SD_OK;
  const bool shared_t = sched != nullptr || t_stride == 0;
  const int n_t = shared_t ? 1 : B;
  cudaStream_t st = (cudaStream_t)stream;
  temb_dense_kernel<<<n_t, 512, sizeof(float) * 5 * nf, st>>>(t_dev, t_stride, sched, step_counter, nf, w0, b0, w1, b1, temb_scratch);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return check_cuda(err, "sd_time_embedding (dense) launch");
  const size_t total = (size_t)B * 4 * nf;
  temb_act_kernel<<<grid_for(total, 256), 256, 0, st>>>(temb_scratch, shared_t ? 0 : 4 * nf, class_emb, labels, B, 4 * nf,
                                                        (__nv_bfloat16*)act_temb_out);
  return check_cuda(cudaGetLastError(), "sd_time_embedding (act) launch");
}

}  // extern "C"
