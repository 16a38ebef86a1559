// Memory-bound and small ops of the CIFAR score-net forward (sm_100a), NHWC bf16 activations.
// Reference call sites: cifar/models/ddpm.py:47-101, cifar/models/layers.py:450-565,
// cifar/models/normalization.py:38-39 (flax nn.GroupNorm defaults).
#include "common.cuh"
#include "../../include/superdiff_b200.h"
#include <cuda_bf16.h>
#include <cstdlib>

namespace sdb {

// swish(v) = v * sigmoid(v) = h + h * tanh(h), h = v / 2: ONE MUFU op (tanh.approx, max rel. error 2^-11, below the bf16
// rounding of the result) instead of ex2 + rcp -- the GroupNorm apply pass needs ~1.5 T elements/s to stream at HBM rate,
// and two MUFU ops per element is 70 % of the chip's 16 / clk / SM
__device__ __forceinline__ float swish_f(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

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

// 1 / sqrt(v) to fp32 accuracy: MUFU.RSQ (2 ulp) + one Newton step
__device__ __forceinline__ float rstd_f32(float v) {
  const float r = rsqrtf(v);
  return r * fmaf(-0.5f * v * r, r, 1.5f);
}

// exact form for the FP32-faithful arm (SD_GEMM_SPLIT3 activations): ex2.approx + division, ~1e-6 relative
__device__ __forceinline__ float swish_exact_f(float v) { return v / (1.f + __expf(-v)); }
template <bool SPLIT>
__device__ __forceinline__ float act_swish(float v) { return SPLIT ? swish_exact_f(v) : swish_f(v); }

// Split activations (SD_GEMM_SPLIT3): a pixel row is [hi(C) | lo(C)] bf16, value = hi + lo.  load8 / store8 move 8 channels.
template <bool SPLIT>
__device__ __forceinline__ void load8(const __nv_bfloat16* p, int lo_off, float (&f)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(p);
  if (SPLIT) {
    const uint4 l = *reinterpret_cast<const uint4*>(p + lo_off);
    float g[8];
    unpack8(h, f); unpack8(l, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] += g[e];
  } else {
    unpack8(h, f);
  }
}
template <bool SPLIT>
__device__ __forceinline__ void store8(__nv_bfloat16* p, int lo_off, const float (&f)[8]) {
  const uint4 h = pack8(f);
  *reinterpret_cast<uint4*>(p) = h;
  if (SPLIT) {
    float hf[8], l[8];
    unpack8(h, hf);
#pragma unroll
    for (int e = 0; e < 8; ++e) l[e] = f[e] - hf[e];
    *reinterpret_cast<uint4*>(p + lo_off) = pack8(l);
  }
}

// ---------------------------------------------------------------------------
// GroupNorm(32) + swish over concat(x0, x1) along channels, as two streaming passes:
//   gn_stats_kernel : per (sample, pixel chunk) per-channel sum / sum-of-squares -> partial[B][nchunk][2][C]
//   gn_apply_kernel : group mean / rstd from the partials in its prologue (fixed order => deterministic), then
//                     y = swish((x - mean) * rstd * gamma + beta) streamed with 16-byte vectors.
// Whole pixel rows are read (fully coalesced), nothing is register- or cluster-resident, so occupancy stays high.
// Measured alternatives on B200 (profiles/): a cluster-per-sample register-resident kernel ran at 0.6-1.6 TB/s
// (cluster launches of memory-bound kernels cost 2-3x; one fat CTA per SM serialises load/reduce/store).
// Algorithmic bytes: 2 reads + 1 write of the activation (the second read mostly hits L2 for the
// chunk sizes used here).
// ---------------------------------------------------------------------------
struct GnStatsParams {
  const __nv_bfloat16* x;    // [B][HW][C]
  int C, B, HW, nchunk, px_per_chunk;
  float* partial;            // [B][nchunk][2][C]
};

struct GnParams {
  const __nv_bfloat16* x0; const __nv_bfloat16* x1;
  int C0, C1, B, HW;
  int nchunk, px_per_chunk;  // apply-pass chunking
  const float* gamma; const float* beta;
  float eps; int apply_swish;
  double inv_n;                   // 1 / (HW * channels per group), formed on the host: the kernels multiply instead of calling the
                                  // fp64 division subroutine (fp64 division / sqrt on the critical path cost microseconds per CTA)
  const float* part0; int nch0;   // per-source channel partials [B][nch][2][Csrc]
  const float* part1; int nch1;
  __nv_bfloat16* out;
  int reverse;                    // walk the CTAs from the last sample to the first (see sd_groupnorm_swish)
};

template <bool SPLIT>
__global__ void __launch_bounds__(256) gn_stats_kernel(const __grid_constant__ GnStatsParams p) {
  extern __shared__ float gn_smem[];   // [rows_per_pass][2*C]
  const int C = p.C, VC = C / 8;
  const int sample = blockIdx.x / p.nchunk, chunk = blockIdx.x - sample * p.nchunk;
  const int cv = threadIdx.x % VC, r = threadIdx.x / VC, rows_per_pass = blockDim.x / VC;
  const int c0 = cv * 8;
  const int ld = SPLIT ? 2 * C : C;              // split rows are [hi(C) | lo(C)]
  const __nv_bfloat16* src = p.x + (size_t)sample * p.HW * ld + c0;
  const int px0 = chunk * p.px_per_chunk, px1 = min(p.HW, px0 + p.px_per_chunk);
  pdl_wait();
  pdl_launch_dependents();
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  int px = px0 + r;
  for (; px + 3 * rows_per_pass < px1; px += 4 * rows_per_pass) {      // 4 independent 16-byte loads in flight
    float fv[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) load8<SPLIT>(src + (size_t)(px + j * rows_per_pass) * ld, C, fv[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += fv[j][e]; q[e] = fmaf(fv[j][e], fv[j][e], q[e]); }
    }
  }
  for (; px < px1; px += rows_per_pass) {
    float f[8];
    load8<SPLIT>(src + (size_t)px * ld, C, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    gn_smem[(size_t)r * 2 * C + c0 + e] = s[e];
    gn_smem[(size_t)r * 2 * C + C + c0 + e] = q[e];
  }
  __syncthreads();
  float* dst = p.partial + ((size_t)sample * p.nchunk + chunk) * 2 * C;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    float a = 0.f;
    for (int rr = 0; rr < rows_per_pass; ++rr) a += gn_smem[(size_t)rr * 2 * C + i];   // fixed order
    dst[i] = a;
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(256) gn_apply_kernel(const __grid_constant__ GnParams p) {
  extern __shared__ float ch_tot[];     // [2][C] (dynamic: a small footprint lets these CTAs share an SM with a resident GEMM CTA)
  __shared__ float g_stat[64];
  const int C = p.C0 + p.C1, VC = C / 8, cpg = C / 32;
  const int bid = p.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int sample = bid / p.nchunk, chunk = bid - sample * p.nchunk;
  pdl_wait();
  pdl_launch_dependents();
  // prologue: every CTA rebuilds its sample's group statistics from the per-chunk channel sums (a few KB from L2;
  // chunks and channels are summed in a fixed order, so all CTAs of a sample -- and every run -- get identical bits)
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int which = i >= C ? 1 : 0, c = i - which * C;   // a group may straddle the x0 | x1 boundary: per channel
    const bool in0 = c < p.C0;
    const int Cs = in0 ? p.C0 : p.C1, cl = in0 ? c : c - p.C0, nch = in0 ? p.nch0 : p.nch1;
    const float* base = (in0 ? p.part0 : p.part1) + (size_t)sample * nch * 2 * Cs + which * Cs + cl;
    float a = 0.f;
    for (int k = 0; k < nch; ++k) a += base[(size_t)k * 2 * Cs];
    ch_tot[i] = a;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double sum = 0.0, sq = 0.0;
    for (int c = 0; c < cpg; ++c) {
      sum += (double)ch_tot[threadIdx.x * cpg + c];
      sq += (double)ch_tot[C + threadIdx.x * cpg + c];
    }
    const double mean = sum * p.inv_n;
    double var = sq * p.inv_n - mean * mean;      // flax: E[x^2] - E[x]^2, clipped at 0 (fp64 FMAs only: no division subroutine)
    if (var < 0.0) var = 0.0;
    g_stat[threadIdx.x * 2] = (float)mean;
    g_stat[threadIdx.x * 2 + 1] = rstd_f32((float)var + p.eps);
  }
  __syncthreads();
  const int cv = threadIdx.x % VC, r = threadIdx.x / VC, rows_per_pass = blockDim.x / VC;
  const int c0 = cv * 8;
  const bool from0 = c0 < p.C0;
  const int smul = SPLIT ? 2 : 1;
  const int src_ld = smul * (from0 ? p.C0 : p.C1), src_lo = from0 ? p.C0 : p.C1;
  const __nv_bfloat16* src = from0 ? p.x0 + (size_t)sample * p.HW * src_ld + c0
                                   : p.x1 + (size_t)sample * p.HW * src_ld + (c0 - p.C0);
  const int dst_ld = smul * C;
  float sc[8], sh[8];                       // y = x * sc + sh
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const float rs = g_stat[(c / cpg) * 2 + 1] * p.gamma[c];
    sc[e] = rs;
    sh[e] = p.beta[c] - g_stat[(c / cpg) * 2] * rs;
  }
  __nv_bfloat16* dst = p.out + (size_t)sample * p.HW * dst_ld + c0;
  const int px0 = chunk * p.px_per_chunk, px1 = min(p.HW, px0 + p.px_per_chunk);
  int px = px0 + r;
  for (; px + 3 * rows_per_pass < px1; px += 4 * rows_per_pass) {
    float fv[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) load8<SPLIT>(src + (size_t)(px + j * rows_per_pass) * src_ld, src_lo, fv[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = fmaf(fv[j][e], sc[e], sh[e]);
        fv[j][e] = p.apply_swish ? act_swish<SPLIT>(y) : y;
      }
      store8<SPLIT>(dst + (size_t)(px + j * rows_per_pass) * dst_ld, C, fv[j]);
    }
  }
  for (; px < px1; px += rows_per_pass) {
    float f[8];
    load8<SPLIT>(src + (size_t)px * src_ld, src_lo, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float y = fmaf(f[e], sc[e], sh[e]);
      f[e] = p.apply_swish ? act_swish<SPLIT>(y) : y;
    }
    store8<SPLIT>(dst + (size_t)px * dst_ld, C, f);
  }
}

// Small activations (H*W*C <= 32768 elements per sample: the 8x8 and 4x4 levels): one CTA per sample, the whole sample
// register-resident -> ONE kernel, one read and one write (the streaming form needs a statistics launch + an apply
// launch and re-reads the tensor; at 4-33 MB per tensor that was 22-41 us per GroupNorm, launch-latency bound).
// Requires C/8 to divide 256 and C/32 to be a multiple of 8 (C = 256, 512), so a thread's 8 channels sit in one group.
template <int NV, bool SPLIT>
__global__ void __launch_bounds__(256) gn_small_kernel(const __grid_constant__ GnParams p) {
  __shared__ float part[256 * 2];
  __shared__ float g_stat[64];
  const int C = p.C0 + p.C1, VC = C / 8, cpg = C / 32, vpg = cpg / 8;   // vectors per group
  const int sample = blockIdx.x, tid = threadIdx.x;
  const int cv = tid % VC, r = tid / VC, rows_per_pass = 256 / VC;
  const int c0 = cv * 8;
  const bool from0 = c0 < p.C0;
  const int smul = SPLIT ? 2 : 1;
  const int src_ld = smul * (from0 ? p.C0 : p.C1), src_lo = from0 ? p.C0 : p.C1;
  const __nv_bfloat16* src = from0 ? p.x0 + (size_t)sample * p.HW * src_ld + c0 : p.x1 + (size_t)sample * p.HW * src_ld + (c0 - p.C0);
  pdl_wait();
  pdl_launch_dependents();
  uint4 v[NV], vl[SPLIT ? NV : 1];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = *reinterpret_cast<const uint4*>(src + (size_t)(r + j * rows_per_pass) * src_ld);
    if (SPLIT) vl[j] = *reinterpret_cast<const uint4*>(src + (size_t)(r + j * rows_per_pass) * src_ld + src_lo);
  }
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float f[8];
    unpack8(v[j], f);
    if (SPLIT) {
      float g[8];
      unpack8(vl[j], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += g[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += f[e]; q = fmaf(f[e], f[e], q); }
  }
  part[tid * 2] = s;
  part[tid * 2 + 1] = q;
  __syncthreads();
  if (tid < 32) {        // group tid: vectors [tid*vpg, (tid+1)*vpg) of every pixel row, fixed order
    double sum = 0.0, sq = 0.0;
    for (int rr = 0; rr < rows_per_pass; ++rr)
      for (int k = 0; k < vpg; ++k) {
        const int t = rr * VC + tid * vpg + k;
        sum += (double)part[t * 2];
        sq += (double)part[t * 2 + 1];
      }
    const double mean = sum * p.inv_n;
    double var = sq * p.inv_n - mean * mean;
    if (var < 0.0) var = 0.0;
    g_stat[tid * 2] = (float)mean;
    g_stat[tid * 2 + 1] = rstd_f32((float)var + p.eps);
  }
  __syncthreads();
  const int grp = c0 / cpg;
  const float mean = g_stat[grp * 2], rstd = g_stat[grp * 2 + 1];
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float rs = rstd * p.gamma[c0 + e];
    sc[e] = rs;
    sh[e] = p.beta[c0 + e] - mean * rs;
  }
  __nv_bfloat16* dst = p.out + (size_t)sample * p.HW * smul * C + c0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float f[8];
    unpack8(v[j], f);
    if (SPLIT) {
      float g[8];
      unpack8(vl[j], g);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += g[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float y = fmaf(f[e], sc[e], sh[e]);
      f[e] = p.apply_swish ? act_swish<SPLIT>(y) : y;
    }
    store8<SPLIT>(dst + (size_t)(r + j * rows_per_pass) * smul * C, C, f);
  }
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

// FP32-faithful attention probabilities: P[r, :] = softmax over the row's diagonal block of `block` columns of scale * X[r, :]
// (rows of a batch entry of `rows_per_entry` rows; off-block entries 0), fp32 in, hi|lo bf16 pair out [rows][2*cols]; one warp per row.
__global__ void __launch_bounds__(256) softmax_rows_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                                 long rows, int cols, float scale, int block, int rows_per_entry) {
  const long row = (long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * cols;
  const int rl = (int)(row % rows_per_entry);
  const int lo = (rl / block) * block, hi = lo + block;
  float mx = -INFINITY;
  for (int j = lo + lane; j < hi; j += 32) mx = fmaxf(mx, xr[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lo + lane; j < hi; j += 32) sum += expf(xr[j] * scale - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  __nv_bfloat16* orow = out + row * 2 * cols;
  for (int j = lane; j < cols; j += 32) {
    const float pj = (j >= lo && j < hi) ? expf(xr[j] * scale - mx) * inv : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(pj);
    orow[j] = h;
    orow[cols + j] = __float2bfloat16_rn(pj - __bfloat162float(h));
  }
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
template <int Cin>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ x, int B, int H, int W,
                                                      const float* __restrict__ w, const float* __restrict__ bias,
                                                      int Cout, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float cin_smem[];   // [9*Cin][Cout] weights, then [Cout] bias
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) cin_smem[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) cin_smem[K * Cout + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int groups = Cout / 32;
  const size_t npix = (size_t)B * H * W;
  const size_t total = ((npix + 31) / 32) * 32 * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // a warp = 32 consecutive pixels of ONE 32-channel group: weight reads are smem broadcasts (the previous
    // pixel-major mapping had 4 channel groups per quarter-warp at a 128-byte stride = 4-way bank conflicts)
    const size_t wi = i >> 5;
    const int cg0 = (int)(wi % groups) * 32;
    size_t pix = (wi / groups) * 32 + (i & 31);
    if (pix >= npix) continue;
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
#pragma unroll
        for (int c = 0; c < Cin; ++c)
          in[(kh * 3 + kw) * 4 + c] = okp ? x[(((size_t)b * H + hi) * W + wi) * Cin + c] : 0.f;
      }
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = cin_smem[K * Cout + cg0 + j];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
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

// First conv on the tensor cores: gather the 3x3xCin (<= 27) fp32 neighbourhood of every pixel into one 64-wide bf16 K-block,
// split as hi = bf16(v), lo = bf16(v - hi) so the fp32 input keeps ~16 mantissa bits:  row = [hi(9*Cin) | lo(9*Cin) | 0...].
// The conv is then a 1-tap implicit GEMM with K = 64 against [w | w | 0] (sd_conv_gemm: bias, GroupNorm statistics for free).
// Tile form (128 % W == 0, whole image rows per 128-pixel tile): the tile's input rows plus a zero halo are staged in shared
// memory, then eight lanes share one pixel -- lane l8 forms entries [8 l8, 8 l8 + 8) of the row and stores them as one 16-byte
// piece, so a warp store instruction writes four whole 128-byte rows.  Which (tap, channel, hi / lo) an entry is depends only on
// the lane and is decoded once.  (The first version built the row in a per-thread local array and stored 16 bytes per lane into 32
// different rows: 50.7 us for 67 MB at batch 512; eight lanes per pixel reading global memory directly: 65.5 us.)
__global__ void __launch_bounds__(256) im2col_in_tile_kernel(const float* __restrict__ x, int B, int H, int W, int Cin,
                                                             __nv_bfloat16* __restrict__ out, int row_elems) {
  extern __shared__ float im_tile[];            // [(R + 2)][(W + 2)][Cin], R = 128 / W image rows
  const int R = 128 / W, tiles_per_img = H / R, Wp = W + 2;
  const int ntiles = B * tiles_per_img;
  const int K = 9 * Cin;
  const int l8 = threadIdx.x & 7;
  int off[8];
  unsigned lo_mask = 0, ok_mask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    int e = l8 * 8 + j;
    const bool lo = e >= K;
    if (lo) e -= K;
    const bool ok = e < K;
    const int tap = ok ? e / Cin : 0;
    off[j] = ((tap / 3) * Wp + tap % 3) * Cin + (ok ? e - tap * Cin : 0);     // halo-shifted: (dh + 1, dw + 1)
    if (lo) lo_mask |= 1u << j;
    if (ok) ok_mask |= 1u << j;
  }
  const int fill = (R + 2) * Wp * Cin;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img, r0 = (tile - b * tiles_per_img) * R;
    __syncthreads();                            // the previous tile has been read
    for (int i = threadIdx.x; i < fill; i += blockDim.x) {
      const int c = i % Cin, q = i / Cin;
      const int ww = q % Wp - 1, hh = r0 + q / Wp - 1;
      im_tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + (((size_t)b * H + hh) * W + ww) * Cin + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int pl = pass * 32 + (threadIdx.x >> 3);                  // pixel within the tile
      const int r = pl / W, wo = pl - r * W;
      const float* src = im_tile + (r * Wp + wo) * Cin;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = ((ok_mask >> j) & 1u) ? src[off[j]] : 0.f;
        const float h = __bfloat162float(__float2bfloat16_rn(v));
        f[j] = ((lo_mask >> j) & 1u) ? v - h : v;          // rounded to bf16 by the pack below: hi = bf16(v), lo = bf16(v - hi)
      }
      const __nv_bfloat162 a0 = __floats2bfloat162_rn(f[0], f[1]), a1 = __floats2bfloat162_rn(f[2], f[3]);
      const __nv_bfloat162 a2 = __floats2bfloat162_rn(f[4], f[5]), a3 = __floats2bfloat162_rn(f[6], f[7]);
      uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)tile * 128 + pl) * row_elems);
      dst[l8] = make_uint4(*reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                           *reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&a3));
      // SD_GEMM_SPLIT3 layout: the 64-wide block above is the "hi" half of a [hi(64) | lo(64)] pair whose lo half is zero
      for (int i = 8 + l8; i < row_elems / 8; i += 8) dst[i] = make_uint4(0, 0, 0, 0);
    }
  }
}

// Any geometry: one thread per pixel.
__global__ void __launch_bounds__(256) im2col_in_kernel(const float* __restrict__ x, int B, int H, int W, int Cin,
                                                        __nv_bfloat16* __restrict__ out, int row_elems) {
  const size_t npix = (size_t)B * H * W;
  const int K = 9 * Cin;
  for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * blockDim.x) {
    size_t q = pix;
    const int wo = q % W; q /= W;
    const int ho = q % H;
    const int b = (int)(q / H);
    __nv_bfloat16 row[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) row[i] = __float2bfloat16_rn(0.f);
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const int hi = ho + kh - 1, wi = wo + kw - 1;
        if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
        const float* src = x + (((size_t)b * H + hi) * W + wi) * Cin;
        for (int c = 0; c < Cin; ++c) {
          const float v = src[c];
          const __nv_bfloat16 h16 = __float2bfloat16_rn(v);
          row[(kh * 3 + kw) * Cin + c] = h16;
          row[K + (kh * 3 + kw) * Cin + c] = __float2bfloat16_rn(v - __bfloat162float(h16));
        }
      }
    uint4* dst = reinterpret_cast<uint4*>(out + pix * row_elems);
    const uint4* r4 = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = r4[i];
    for (int i = 8; i < row_elems / 8; ++i) dst[i] = make_uint4(0, 0, 0, 0);
  }
}

// temb = Dense1(swish(Dense0(sinusoidal(t))))  for n_t distinct times -> fp32 scratch [n_t, 4nf]
__global__ void __launch_bounds__(512) temb_dense_kernel(const float* __restrict__ t_dev, int t_stride,
                                                         const float* __restrict__ sched, const int* __restrict__ step_counter,
                                                         int nf, const float* __restrict__ w0, const float* __restrict__ b0,
                                                         const float* __restrict__ w1, const float* __restrict__ b1,
                                                         float* __restrict__ temb, int exact) {
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
    h1[j] = exact ? swish_exact_f(acc) : swish_f(acc);
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
                                                       int B, int nh, __nv_bfloat16* __restrict__ out, int split) {
  const size_t total = (size_t)B * nh;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = i / nh, j = i % nh;
    float v = temb[(size_t)b * temb_stride + j];
    if (class_emb) v += class_emb[(size_t)labels[b] * nh + j];
    if (split) {        // exact swish, rows [hi(nh) | lo(nh)]
      const float a = swish_exact_f(v);
      const __nv_bfloat16 h = __float2bfloat16_rn(a);
      out[(size_t)b * 2 * nh + j] = h;
      out[(size_t)b * 2 * nh + nh + j] = __float2bfloat16_rn(a - __bfloat162float(h));
    } else {
      out[i] = __float2bfloat16_rn(swish_f(v));
    }
  }
}

// out[0..n) = table[row * n .. ), row read from a device counter: lets a captured CUDA graph pick the current timestep's
// precomputed row (time-embedding biases) without any host value baked in
__global__ void __launch_bounds__(256) gather_row_kernel(const float4* __restrict__ table, int n4, const int* __restrict__ counter,
                                                         int max_row, float4* __restrict__ out) {
  int row = counter ? *counter : 0;
  row = row < 0 ? 0 : (row > max_row ? max_row : row);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) out[i] = table[(size_t)row * n4 + i];
}

static unsigned grid_for(size_t total, int threads) {
  const size_t want = (total + threads - 1) / threads;
  const size_t cap = 148 * 16;
  return (unsigned)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace sdb

extern "C" {

int sd_groupnorm_swish(const void* x0, int C0, const void* x1, int C1, int B, int HW, const float* gamma,
                       const float* beta, float eps, int apply_swish, const float* stats0, int nchunk0,
                       const float* stats1, int nchunk1, float* scratch, size_t scratch_floats, void* out,
                       void* stream) {
  return sd_groupnorm_swish_ex(x0, C0, x1, C1, B, HW, gamma, beta, eps, apply_swish, stats0, nchunk0, stats1, nchunk1, scratch,
                               scratch_floats, out, 0u, stream);
}

int sd_groupnorm_swish_ex(const void* x0, int C0, const void* x1, int C1, int B, int HW, const float* gamma,
                          const float* beta, float eps, int apply_swish, const float* stats0, int nchunk0,
                          const float* stats1, int nchunk1, float* scratch, size_t scratch_floats, void* out,
                          unsigned flags, void* stream) {
  using namespace sdb;
  const bool split = (flags & SD_GEMM_SPLIT3) != 0;
  if (!x0 || !gamma || !beta || !out || !scratch || (C1 > 0 && !x1)) return fail(kErrInvalidArg, "sd_groupnorm_swish: null pointer");
  if (C1 < 0) C1 = 0;
  const int C = C0 + C1;
  if (C0 < 8 || C0 % 8 || C1 % 8 || C % 32 || C > 1024) return fail(kErrInvalidArg, "sd_groupnorm_swish: channels must be multiples of 8, total a multiple of 32, <= 1024");
  if (B < 0 || HW < 1) return fail(kErrInvalidArg, "sd_groupnorm_swish: bad shape");
  if ((stats0 && nchunk0 < 1) || (stats1 && nchunk1 < 1)) return fail(kErrInvalidArg, "sd_groupnorm_swish: stats need nchunk >= 1");
  if (B == 0) return SD_OK;
  static const int cta_target = [] { const char* e = getenv("SDB_GN_CTAS"); return e ? atoi(e) : 148 * 4; }();   // tuning knob
  // SDB_GN_CARVEOUT=1: same shared-memory carveout as the tcgen05 GEMM (max shared), so that these memory-bound CTAs may share
  // an SM with the other stream's resident GEMM CTA (the GEMM ring leaves ~12 KB free for that).  Measured on B200: the step
  // time does not change (14.04 vs 14.02 ms) while the GroupNorm kernels alone get 11 % slower without L1, so the default is off.
  static const bool carve_once = [] {
    const char* e = getenv("SDB_GN_CARVEOUT");
    if (!e || atoi(e) == 0) return false;
    cudaFuncSetAttribute(gn_apply_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_stats_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_small_kernel<1, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_small_kernel<2, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_small_kernel<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_small_kernel<8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gn_small_kernel<16, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)carve_once;
  cudaStream_t st = (cudaStream_t)stream;
  auto threads_for = [](int Cc, int& k) {          // threads = (Cc/8) * k pixel rows per pass, whole warps, <= 256
    const int VC = Cc / 8;
    k = 256 / VC;
    if (k < 1) k = 1;
    while (k > 1 && (VC * k) % 32) --k;
    return VC * k;
  };
  auto chunks_for = [&](int k, int& px_per_chunk) {
    int nchunk = (cta_target + B - 1) / B;
    // 16x16 (and smaller) images at batches that fill the chip on their own: one CTA per sample.  Measured at batch 512
    // (tools/layer_bench.py with SDB_GN_CTAS, profiles/r01d_notes.md): 24.8 vs 28.2 us for 16^2 x 256, while the 32^2 layers
    // prefer two chunks per sample (48.1 vs 50.8 us); more, smaller CTAs are slower everywhere (1184 .. 9472: +6 .. +40 %).
    if (HW <= 256 && B >= 296) nchunk = 1;
    const int max_chunks = (HW + 4 * k - 1) / (4 * k);
    if (nchunk > max_chunks) nchunk = max_chunks;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 64) nchunk = 64;
    px_per_chunk = (HW + nchunk - 1) / nchunk;
    return (HW + px_per_chunk - 1) / px_per_chunk;
  };
  GnParams p{};
  // SDB_GN_REVERSE=1: the apply pass walks the tensor from its end.  The producing GEMM wrote the end last, so that part
  // is still in the 126 MB L2; and this pass then writes the START last, which is what the consuming GEMM reads first.
  static const int gn_reverse = [] { const char* e = getenv("SDB_GN_REVERSE"); return e && *e ? atoi(e) : 0; }();
  p.reverse = gn_reverse;
  p.x0 = (const __nv_bfloat16*)x0; p.x1 = (const __nv_bfloat16*)x1;
  p.C0 = C0; p.C1 = C1; p.B = B; p.HW = HW;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.apply_swish = apply_swish;
  p.inv_n = 1.0 / ((double)HW * (double)(C / 32));
  p.out = (__nv_bfloat16*)out;
  if (!stats0 && !stats1 && (C == 256 || C == 512) && ((size_t)HW * C / 8) % 256 == 0) {
    const size_t nv = (size_t)HW * C / 8 / 256;
    void (*kern)(const GnParams) = nullptr;
    if (nv == 1) kern = split ? gn_small_kernel<1, true> : gn_small_kernel<1, false>;
    else if (nv == 2) kern = split ? gn_small_kernel<2, true> : gn_small_kernel<2, false>;
    else if (nv == 4) kern = split ? gn_small_kernel<4, true> : gn_small_kernel<4, false>;
    else if (nv == 8) kern = split ? gn_small_kernel<8, true> : gn_small_kernel<8, false>;
    else if (nv == 16) kern = split ? gn_small_kernel<16, true> : gn_small_kernel<16, false>;
    if (kern) {
      return check_cuda(launch_pdl(kern, dim3((unsigned)B), dim3(256), 0, st, p), "sd_groupnorm_swish (small) launch");
    }
  }
  size_t used = 0;
  // statistics pass only for the sources whose producer did not already emit per-tile channel sums
  for (int srcI = 0; srcI < 2; ++srcI) {
    const int Cs = srcI == 0 ? C0 : C1;
    if (Cs == 0) { if (srcI == 1) { p.part1 = scratch; p.nch1 = 0; } continue; }
    const float* given = srcI == 0 ? stats0 : stats1;
    if (given) {
      if (srcI == 0) { p.part0 = given; p.nch0 = nchunk0; } else { p.part1 = given; p.nch1 = nchunk1; }
      continue;
    }
    int k;
    const int T = threads_for(Cs, k);
    if (T % 32 || T > 256) return fail(kErrUnsupported, "sd_groupnorm_swish: unsupported channel count");
    GnStatsParams sp{srcI == 0 ? p.x0 : p.x1, Cs, B, HW, 0, 0, scratch + used};
    sp.nchunk = chunks_for(k, sp.px_per_chunk);
    const size_t need = (size_t)B * sp.nchunk * 2 * Cs;
    if (used + need > scratch_floats) return fail(kErrInvalidArg, "sd_groupnorm_swish: scratch too small");
    used += need;
    if (srcI == 0) { p.part0 = sp.partial; p.nch0 = sp.nchunk; } else { p.part1 = sp.partial; p.nch1 = sp.nchunk; }
    cudaError_t err = launch_pdl(split ? gn_stats_kernel<true> : gn_stats_kernel<false>, dim3((unsigned)(B * sp.nchunk)), dim3(T),
                                 sizeof(float) * (size_t)k * 2 * Cs, st, sp);
    if (err != cudaSuccess) return check_cuda(err, "sd_groupnorm_swish (stats) launch");
  }
  int k;
  const int T = threads_for(C, k);
  if (T % 32 || T > 256) return fail(kErrUnsupported, "sd_groupnorm_swish: unsupported channel count");
  p.nchunk = chunks_for(k, p.px_per_chunk);
  return check_cuda(launch_pdl(split ? gn_apply_kernel<true> : gn_apply_kernel<false>, dim3((unsigned)(B * p.nchunk)), dim3(T),
                               sizeof(float) * 2 * (size_t)C, st, p),
                    "sd_groupnorm_swish (apply) launch");
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

int sd_softmax_rows_split(const float* x, void* out, long rows, int cols, float scale, int block, int rows_per_entry, void* stream) {
  using namespace sdb;
  if (!x || !out || rows < 0 || cols < 1 || block < 1 || rows_per_entry < 1 || (cols % block) != 0 || (rows_per_entry % block) != 0 ||
      rows_per_entry > cols)
    return fail(kErrInvalidArg, "sd_softmax_rows_split: block must divide cols and rows_per_entry (<= cols)");
  if (rows == 0) return SD_OK;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  softmax_rows_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out, rows, cols, scale, block, rows_per_entry);
  return check_cuda(cudaGetLastError(), "sd_softmax_rows_split launch");
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
  static PerDeviceOnce attr_once;
  attr_once.run([] {
    cudaFuncSetAttribute(conv_in_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(conv_in_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(conv_in_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    return cudaFuncSetAttribute(conv_in_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  });
  const size_t total = (((size_t)B * H * W + 31) / 32) * 32 * (Cout / 32);
  const unsigned grid = grid_for(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  switch (Cin) {
    case 1: conv_in_kernel<1><<<grid, 256, smem, st>>>(x, B, H, W, w_hwio, bias, Cout, o); break;
    case 2: conv_in_kernel<2><<<grid, 256, smem, st>>>(x, B, H, W, w_hwio, bias, Cout, o); break;
    case 3: conv_in_kernel<3><<<grid, 256, smem, st>>>(x, B, H, W, w_hwio, bias, Cout, o); break;
    default: conv_in_kernel<4><<<grid, 256, smem, st>>>(x, B, H, W, w_hwio, bias, Cout, o); break;
  }
  return check_cuda(cudaGetLastError(), "sd_conv_in launch");
}

int sd_gather_row(const float* table, int rows, int row_floats, const int* counter, float* out, void* stream) {
  using namespace sdb;
  if (!table || !out || rows < 1 || row_floats < 4 || (row_floats % 4) || ((uintptr_t)table % 16) || ((uintptr_t)out % 16))
    return fail(kErrInvalidArg, "sd_gather_row: row length must be a positive multiple of 4 floats, pointers 16-byte aligned");
  const int n4 = row_floats / 4;
  gather_row_kernel<<<(n4 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)table, n4, counter, rows - 1, (float4*)out);
  return check_cuda(cudaGetLastError(), "sd_gather_row launch");
}

int sd_im2col_in(const float* x, int B, int H, int W, int Cin, void* out, void* stream) {
  return sd_im2col_in_ex(x, B, H, W, Cin, out, 0u, stream);
}

int sd_im2col_in_ex(const float* x, int B, int H, int W, int Cin, void* out, unsigned flags, void* stream) {
  using namespace sdb;
  if (!x || !out || B < 0 || H < 1 || W < 1 || Cin < 1 || 18 * Cin > 64)
    return fail(kErrInvalidArg, "sd_im2col_in: 1 <= Cin <= 3 required (hi/lo split of 9*Cin values must fit one 64-wide K-block)");
  if (B == 0) return SD_OK;
  const size_t npix = (size_t)B * H * W;
  const int row_elems = (flags & SD_GEMM_SPLIT3) ? 128 : 64;
  if (W <= 128 && (128 % W) == 0 && (H % (128 / W)) == 0) {
    const int ntiles = B * (H / (128 / W));
    const size_t smem = (size_t)(128 / W + 2) * (W + 2) * Cin * sizeof(float);
    const int cap = 148 * 8;
    im2col_in_tile_kernel<<<ntiles < cap ? ntiles : cap, 256, smem, (cudaStream_t)stream>>>(x, B, H, W, Cin, (__nv_bfloat16*)out, row_elems);
  } else {
    im2col_in_kernel<<<grid_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, Cin, (__nv_bfloat16*)out, row_elems);
  }
  return check_cuda(cudaGetLastError(), "sd_im2col_in launch");
}

int sd_time_embedding(const float* t_dev, int t_stride, const float* sched, const int* step_counter, int B, int nf,
                      const float* w0, const float* b0, const float* w1, const float* b1, const float* class_emb,
                      const int* labels, float* temb_scratch, void* act_temb_out, void* stream) {
  return sd_time_embedding_ex(t_dev, t_stride, sched, step_counter, B, nf, w0, b0, w1, b1, class_emb, labels, temb_scratch,
                              act_temb_out, 0u, stream);
}

int sd_time_embedding_ex(const float* t_dev, int t_stride, const float* sched, const int* step_counter, int B, int nf,
                         const float* w0, const float* b0, const float* w1, const float* b1, const float* class_emb,
                         const int* labels, float* temb_scratch, void* act_temb_out, unsigned flags, void* stream) {
  using namespace sdb;
  const int split = (flags & SD_GEMM_SPLIT3) ? 1 : 0;
  if ((!t_dev && !sched) || !w0 || !b0 || !w1 || !b1 || !temb_scratch || !act_temb_out)
    return fail(kErrInvalidArg, "sd_time_embedding: null pointer");
  if (class_emb && !labels) return fail(kErrInvalidArg, "sd_time_embedding: class embedding needs labels");
  if (nf < 4 || nf % 2) return fail(kErrInvalidArg, "sd_time_embedding: nf must be even and >= 4");
  if (B == 0) return SD_OK;
  const bool shared_t = sched != nullptr || t_stride == 0;
  const int n_t = shared_t ? 1 : B;
  cudaStream_t st = (cudaStream_t)stream;
  temb_dense_kernel<<<n_t, 512, sizeof(float) * 5 * nf, st>>>(t_dev, t_stride, sched, step_counter, nf, w0, b0, w1, b1, temb_scratch, split);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return check_cuda(err, "sd_time_embedding (dense) launch");
  const size_t total = (size_t)B * 4 * nf;
  temb_act_kernel<<<grid_for(total, 256), 256, 0, st>>>(temb_scratch, shared_t ? 0 : 4 * nf, class_emb, labels, B, 4 * nf,
                                                        (__nv_bfloat16*)act_temb_out, split);
  return check_cuda(cudaGetLastError(), "sd_time_embedding (act) launch");
}

}  // extern "C"
