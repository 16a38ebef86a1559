// Implicit-GEMM convolution / NIN / Dense on the 5th-gen tensor cores (sm_100a).
//
// Replaces the nn.Conv 3x3 (reference cifar/models/layers.py:95-107), NIN (:464-475)
// and nn.Dense (:556, cifar/models/ddpm.py:65-66) call sites of the CIFAR score-net.
//
//   out[m, n] = sum_k A[m, k] * Wt[n, k] + bias[n] + rowbias[img(m), n] + residual[m, n]
//
// m enumerates output pixels (b, h, w) of an NHWC tensor, k enumerates
// (segment, tap, channel).  A is never materialised: for every 64-channel
// K-block the producer warp issues ONE 4-D TMA load of the (channel, w, h, b)
// box shifted by the filter tap; out-of-bounds coordinates are zero-filled by
// the TMA unit, which is exactly SAME padding.  The box lands in shared memory
// as 128 rows x 128 B with the 128-byte swizzle, i.e. the canonical K-major
// UMMA operand layout, and is consumed by tcgen05.mma (M=128, N<=256, K=16)
// with the fp32 accumulator in tensor memory.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..9 = epilogue (tcgen05.ld -> bias /
// time-embedding / GroupNorm statistics -> global).  Persistent over output
// tiles, a 204 KB smem ring cut into 2..8 stages, two TMEM accumulators so the
// epilogue of tile i overlaps the MMAs of tile i+1.
//
// Tile shapes (chosen per launch in launch_gemm):
//   * N = 256 layers with >= 148 tiles: cta_group::2 pairs (kernel<PAIR = true>), each SM feeds its own pixel tile
//     and half of the weight tile;
//   * N = 128 layers and all 1x1 layers: two adjacent pixel tiles per CTA, OPERANDS SWAPPED -- the weights are the
//     M = 128 operand, the 256 pixels the N = 256 operand, so one 128-clk instruction does the work of two 64-clk
//     ones that would saturate the shared-memory port; pairs are swapped the same way (M = 256 channels);
//   * swapped accumulators are [channel][pixel]: bias is one register, GroupNorm sums are thread-local, a lane-pair
//     exchange packs channel pairs for 64-byte store runs;
//   * 3x3 stride-1 segments are fed as activation slabs shared by the three vertical taps;
//   * everything else (N < 128, flat / batched matrices, softmax epilogue of the attention probabilities used by the
//     JVP path) runs unswapped with thread = pixel row.
//
// Round 2:
//   * SD_GEMM_SPLIT3 (the FP32-faithful arm): operands are hi|lo bf16 pairs, every source contributes three K segments that
//     share weight columns (seg_bk0), the epilogues write / read pairs -- conv_gemm_impl, batched_gemm_impl;
//   * fused GroupNorm + swish epilogue for swapped tiles (sd_conv_gemm_gn): one TMEM pass, values parked as packed bf16
//     registers, fp32 group statistics, 32x32 images exchange channel sums inside a thread-block cluster of 4 through DSMEM;
//   * the four phases of the fused upsample + conv as one launch (up_all: phase = slowest digit of the tile index).
#include "common.cuh"
#include "tcgen05_util.cuh"
#include "../../include/superdiff_b200.h"
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>

#ifndef SDB_RING_KB
#define SDB_RING_KB 204    // 3 x 68 KB pair-slab stages; leaves ~12 KB of the SM for a co-resident GroupNorm CTA of the other stream
#endif

namespace sdb {

constexpr int BM = 128;            // rows (pixels) per tile == TMEM lanes
constexpr int BK = 64;             // bf16 elements per K-block == one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int RING_BYTES = SDB_RING_KB * 1024;     // operand ring, cut into 2..8 slots of the size a launch needs (48 KB K-blocks: 4)
constexpr int MAX_STAGES = 8;
constexpr int MAX_BN = 256;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES_MAX = MAX_BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int MAX_SEGS = 9;            // 3 sources x the three terms of the 3xbf16 split product (SD_GEMM_SPLIT3)
constexpr int MAX_SLABS = 3;           // 9-tap stride-1 segments that may be fed as activation slabs
constexpr int EPI_WARPS = 8;               // two warps per TMEM lane quarter, each takes alternate 32-column groups
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int GEMM_THREADS = 64 + EPI_THREADS;
constexpr int EBIAS_FLOATS = 12 * MAX_BN;          // bias + row-bias table (<= 8 images per tile); with stats_out: 1 row + [4 warps][2][256] column partials;
                                                   // fused GroupNorm on pixel-row tiles: + gamma / beta table [2][256] + warp partials [8][2][32]
constexpr size_t GEMM_SMEM = (size_t)RING_BYTES + EBIAS_FLOATS * 4 + 256 + 1024;  // + barriers + alignment slack

struct GemmParams {
  CUtensorMap a_map[MAX_SEGS];
  CUtensorMap b_map;
  int nseg;
  int seg_taps[MAX_SEGS];
  int seg_cblocks[MAX_SEGS];  // C / 64
  int seg_bk0[MAX_SEGS];      // K-block (column / 64 of the B operand) where the segment's weights start: several segments may
                              // share B columns (split precision: hi and lo activations against the same hi weights)
  int seg_slab[MAX_SEGS];     // -1, or the index of the segment's slab tensor map (3x3 stride-1 segment fed as slabs)
  int flat;                   // 1: matrix mode, A is [batch][M][K] (batch stride may be 0 = shared)
  int a_batched, b_batched;   // matrix mode: does the batch index select an A / B slice
  int m_tiles_per_batch, M_per_batch;
  const void* seg_src[MAX_SEGS];    // host-only: segment sources (slab tensor maps are built from the 9-tap stride-1 ones)
  int seg_C[MAX_SEGS], seg_ld[MAX_SEGS], slab_B;
  int slab;                   // 1: the 3x3 stride-1 segments (seg_slab >= 0) are fed as activation slabs shared by the three vertical taps
  int slab_bytes, a_region_bytes;   // bytes of one slab; bytes reserved for A at the start of a ring slot
  int slab_taps;              // vertical taps that share one slab: 3 (3x3 conv) or 2 (the 2x2-tap phases of the fused upsample + conv)
  CUtensorMap a_slab_map[MAX_SLABS];     // per slab segment: map with a (64, W, hb+2, 1) box
  int num_stages, stage_bytes; // operand ring geometry: stage_bytes = A bytes + B bytes of one K-block (1 KB multiple)
  int pair;                   // 1 (with cluster == 2, non-dual): the two CTAs form one tcgen05 cta_group::2 pair -- M = 256 (two m-tiles),
                              //    each CTA feeds its own A tile and HALF of the weight tile, so an SM receives 32 KB instead of 48 KB per K-block
  int cluster;                // CTAs per thread-block cluster (1, 2, 4): consecutive m-units, same n-tile
  int mcast;                  // CTAs that share one B tile through TMA multicast (== cluster), or 1: the cluster exists only so that
                              //    the CTAs holding one image can exchange GroupNorm sums through distributed shared memory
  const float* gn_gamma;      // fused GroupNorm + swish epilogue (swapped tiles only): out = swish(GN(acc + bias) * gamma + beta)
  const float* gn_beta;       //    with the 32-group statistics of the whole image formed in the epilogue
  float gn_eps;
  int gn_swish;
  __nv_bfloat16* gn_raw_out;  // optional second output of a fused launch: the raw (un-normalised) conv result, bf16, same layout / ld as
                              //    `out` -- for tensors that stay on the residual stream while their GroupNorm feeds the next layer
  int gx_units;               // > 1: the gx_units CTAs that hold one image (consecutive blockIdx, same tile round) exchange their GroupNorm
                              //    channel sums through global memory (gx_data / gx_cnt) instead of a thread-block cluster: clusters of 4
                              //    CTAs at one CTA per SM only fit 33 times on the 148 SMs (GPCs of 18-20 SMs), plain CTAs use all of them
  float* gx_data;             // [2 * grid / gx_units slots][gx_units][2][128] channel sums
  int* gx_cnt;                // [slots][2]: arrivals, readers done (self-resetting: the last reader zeroes both)
  int dual;                   // 1: each CTA tile is TWO adjacent 128-row m-tiles sharing one B tile (block_n <= 128)
  int swap;                   // 1 (dual, conv mode, N = 128): operands swapped inside the MMA -- D^T[128 channels x 256 pixels] =
                              //    W[128 x K] * X^T: ONE M=128, N=256 instruction per K step instead of two N=128 ones (see launch_gemm)
  long long out_batch_stride; // elements
  int h_box, tiles_per_img, imgs_per_tile;
  int M_total, HW, N_out, block_n, n_tiles, m_tiles, num_kb;
  const float* bias;
  const float* rowbias;
  int rowbias_ld;
  const __nv_bfloat16* residual;
  float* stats_out;           // optional [m_tiles][2][N_out]: per-tile column sum / sum of squares of the output
  int res_ld;
  void* out;
  int out_ld;
  unsigned flags;
  int stride2;                // 1: 3x3 stride-2 SAME conv (pad (0,1)): taps read input pixel (2*ho + kh, 2*wo + kw) through a TMA map
                              //    with element strides (1,2,2,1); the tile geometry (h_box, ...) is that of the OUTPUT image
  int up_phase;               // -1, or a*2+b: this launch computes output pixels (2i+a, 2j+b) of a fused nearest-x2 upsample + 3x3 conv
  int up_all;                 // 1: ALL four phases in this launch -- the phase is the slowest digit of the tile index and selects the
                              //    weight matrix (batch coordinate of the B map); up_phase is then only a ">= 0" marker
  int img_H, img_W;           // source image size (conv mode)
  int stats_tpi_total, stats_slot0;   // stats_out slot = img * stats_tpi_total + stats_slot0 + tile-in-image
  int imgs_in_tile;           // images covered by one 128-row tile (row-bias table rows), 1 when HW >= 128 or flat
  float softmax_scale;        // SD_EPI_SOFTMAX: out = softmax(scale * acc) over the row's block of softmax_block columns
  int softmax_block;
};

__device__ __forceinline__ float swishf(float v) { return __fdividef(v, 1.f + __expf(-v)); }
// one-MUFU form used by the GroupNorm kernels of the bf16 arm (scorenet_ops.cu::swish_f): v/2 * (1 + tanh(v/2))
__device__ __forceinline__ float swish_tanh_f(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// Epilogue for 16 consecutive output columns of one row.  `eb`: 16 floats of (bias + row bias) in shared memory or
// nullptr.  p.residual (generic API) is read straight from global memory; the score-net itself adds its residuals
// inside the MMA as an identity-weight K segment, so its epilogues never touch global memory for inputs.
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, const uint32_t (&acc)[16], size_t row_off, int n0,
                                                 const float* eb, float (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (eb) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(eb + j);
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  const bool full = (n0 + 16 <= p.N_out);
  const bool split = (p.flags & SD_GEMM_SPLIT3) != 0;     // residual / out rows are [hi(N) | lo(N)] bf16 pairs, value = hi + lo
  if (full) {
    if (p.residual) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        if (part == 1 && !split) break;
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + n0 + (part ? p.N_out : 0));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 u = rp[h];
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[h * 8 + 2 * j] += __uint_as_float(w[j] << 16);
            v[h * 8 + 2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
          }
        }
      }
    }
    if (p.flags & SD_EPI_SWISH) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = swishf(v[j]);
    }
    if (split && !(p.flags & SD_EPI_OUT_F32)) {
      uint32_t wh[8], wl[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * j] - __low2float(h2), v[2 * j + 1] - __high2float(h2));
        wh[j] = *reinterpret_cast<const uint32_t*>(&h2);
        wl[j] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      uint4* oh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n0);
      uint4* ol = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n0 + p.N_out);
      oh[0] = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      oh[1] = make_uint4(wh[4], wh[5], wh[6], wh[7]);
      ol[0] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      ol[1] = make_uint4(wl[4], wl[5], wl[6], wl[7]);
    } else if (p.flags & SD_EPI_OUT_F32) {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row_off + n0);
#pragma unroll
      for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n0);
      op[0] = make_uint4(w[0], w[1], w[2], w[3]);
      op[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  } else {
    // ragged N tail (e.g. the 3-channel output conv): scalar, masked
    for (int j = 0; j < 16; ++j) {
      const int n = n0 + j;
      if (n >= p.N_out) break;
      float x = v[j];
      if (p.residual) x += __bfloat162float(p.residual[row_off + n]) + (split ? __bfloat162float(p.residual[row_off + p.N_out + n]) : 0.f);
      if (p.flags & SD_EPI_SWISH) x = swishf(x);
      if (p.flags & SD_EPI_OUT_F32) reinterpret_cast<float*>(p.out)[row_off + n] = x;
      else {
        const __nv_bfloat16 hb = __float2bfloat16_rn(x);
        reinterpret_cast<__nv_bfloat16*>(p.out)[row_off + n] = hb;
        if (split) reinterpret_cast<__nv_bfloat16*>(p.out)[row_off + p.N_out + n] = __float2bfloat16_rn(x - __bfloat162float(hb));
      }
    }
  }
}

// Column sums over the 32 rows held by a warp: in = 16 values per lane (one row, 16 columns).  Reduce-scatter
// butterfly over (v, v*v): 31 shuffles instead of 160; on return lane L holds the warp total of
// column (L & 15) -- the plain sum for L < 16, the sum of squares for L >= 16.  Fixed order => deterministic.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float s[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) { s[j] = v[j]; s[16 + j] = v[j] * v[j]; }
#pragma unroll
  for (int step = 16, half = 16; step >= 1; step >>= 1, half >>= 1) {
    const bool upper = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? s[i] : s[i + half];
      const float keep = upper ? s[i + half] : s[i];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return s[0];
}

// output row r (0..127) of m-tile `m_tile` -> validity and element offset of the row start in out / residual
__device__ __forceinline__ bool row_offset(const GemmParams& p, int m_tile, int r, size_t& off, int up_phase) {
  if (p.flat) {
    const int batch = m_tile / p.m_tiles_per_batch;
    const int rl = (m_tile - batch * p.m_tiles_per_batch) * BM + r;
    off = (size_t)batch * (size_t)p.out_batch_stride + (size_t)rl * p.out_ld;
    return rl < p.M_per_batch && m_tile < p.m_tiles;
  }
  const int m = m_tile * BM + r;
  if (p.up_phase >= 0) {      // low-res pixel (img, i, j) -> output pixel (2i+a, 2j+b) of the 2H x 2W image
    const int img = m / p.HW, rem = m - img * p.HW;
    const int i = rem / p.img_W, j = rem - i * p.img_W;
    off = ((size_t)(img * 2 * p.img_H + 2 * i + (up_phase >> 1)) * (2 * p.img_W) + 2 * j + (up_phase & 1)) * (size_t)p.out_ld;
  } else {
    off = (size_t)m * p.out_ld;
  }
  return m < p.M_total;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }   // the epilogue warps only

// Totals over the SEG lanes of an aligned lane group of V per-lane values (butterfly reduce-scatter; plain butterflies once
// there are more lanes than values).  On return a lane holds max(1, V / SEG) totals in v[0..): those of the original values
// ((lane % SEG) * V) / SEG + i.  Fixed order: deterministic.
template <int V, int SEG>
__device__ __forceinline__ void seg_reduce_scatter(float (&v)[V], int lane) {
  int n = V;
#pragma unroll
  for (int o = SEG / 2; o >= 1; o >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        if (i < n) {
          const float send = upper ? v[i] : v[i + n];
          const float keep = upper ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    }
  }
}

// Fused GroupNorm + swish on a thread = pixel-row tile that holds whole images (8x8: two per tile, 4x4: eight): the statistics of
// one (image, 8-channel group) are the sums over the image's rows -- a lane segment of one warp (4x4) or two warps (8x8) -- of a
// thread's 8 adjacent columns.  NCH = 32-column chunks per thread (block_n / 64), HW = pixels per image.
template <int NCH, int HW, bool PAIR>
__device__ __forceinline__ void gn_rows_epilogue(const GemmParams& p, uint32_t taddr, int row, int q, int chalf, int lane, int warp_e, int et,
                                                 const float* eb_row, float* gtab, float* wred, bool row_ok, size_t row_off, int n_base,
                                                 uint64_t* tmem_empty_bar) {
  constexpr int SEG = HW < 32 ? HW : 32;
  for (int n = et; n < 2 * p.block_n; n += EPI_THREADS) {
    const int which = n >= p.block_n ? 1 : 0, nn = n - which * p.block_n;
    gtab[which * MAX_BN + nn] = (which ? p.gn_beta : p.gn_gamma)[n_base + nn];
  }
  uint32_t pk[NCH * 16];
  const int segid = SEG == 16 ? lane >> 4 : 0;
  const int idx0 = ((lane & (SEG - 1)) * 8) / SEG;           // which of a chunk's 8 totals this lane ends up holding
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = chalf * 32 + k * 64;
    float s[8];                              // (sum, sum of squares) of the chunk's four 8-column groups
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[16];                        // 16 columns at a time: the 64 parked registers leave little room (168 per thread)
      tmem_ld16(taddr + c + h * 16, r);
      tmem_wait_ld();
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float x0 = __uint_as_float(r[2 * jj]) + eb_row[c + h * 16 + 2 * jj];
        const float x1 = __uint_as_float(r[2 * jj + 1]) + eb_row[c + h * 16 + 2 * jj + 1];
        const int g = h * 2 + jj / 4;                            // 8 adjacent columns = one group
        s[2 * g] += x0 + x1;
        s[2 * g + 1] = fmaf(x0, x0, fmaf(x1, x1, s[2 * g + 1]));
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
        pk[k * 16 + h * 8 + jj] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    // totals over the rows of the image inside this warp (per chunk: 8 live values instead of 32)
    seg_reduce_scatter<8, SEG>(s, lane);
    wred[(warp_e * 2 + segid) * 32 + k * 8 + idx0] = s[0];
  }
  tcgen05_fence_before();                  // accumulator drained: hand it back to the MMA warp
  __syncwarp();
  if (lane == 0) {
    if constexpr (PAIR) mbar_arrive_rank(tmem_empty_bar, 0);
    else mbar_arrive(tmem_empty_bar);
  }
  epi_bar();
  // this row's image: its lane segment (HW <= 32) or the two warps of its row-quarter pair (HW = 64), same column half
  const float* t0 = wred + (warp_e * 2 + (SEG == 16 ? lane >> 4 : 0)) * 32;
  const float* t1 = t0;
  if (HW == 64) {
    const int pw = chalf * 4 + (((q ^ 1) + 2) & 3);              // partner warp: same chalf, row quarter q ^ 1
    t1 = wred + (pw * 2) * 32;
  }
  const float inv_n = 1.f / (float)(HW * 8);
  __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + n_base;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = chalf * 32 + k * 64;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float mean[2], rstd[2];              // the two groups of these 16 columns (statistics formed here: 32 fewer live registers)
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int g = k * 4 + h * 2 + gg;
        const float sm = HW == 64 ? t0[2 * g] + t1[2 * g] : t0[2 * g];
        const float sq = HW == 64 ? t0[2 * g + 1] + t1[2 * g + 1] : t0[2 * g + 1];
        mean[gg] = sm * inv_n;
        rstd[gg] = rsqrtf(fmaxf(fmaf(sq, inv_n, -mean[gg] * mean[gg]), 0.f) + p.gn_eps);      // flax: E[x^2] - E[x]^2, clipped at 0
      }
      uint32_t w[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int gg = jj / 4, cc = c + h * 16 + 2 * jj;
        const uint32_t u = pk[k * 16 + h * 8 + jj];
        const float sc0 = rstd[gg] * gtab[cc], sc1 = rstd[gg] * gtab[cc + 1];
        float y0 = fmaf(__uint_as_float(u << 16) - mean[gg], sc0, gtab[MAX_BN + cc]);
        float y1 = fmaf(__uint_as_float(u & 0xFFFF0000u) - mean[gg], sc1, gtab[MAX_BN + cc + 1]);
        if (p.gn_swish) { y0 = swish_tanh_f(y0); y1 = swish_tanh_f(y1); }
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
        w[jj] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      if (row_ok) {
        uint4* op = reinterpret_cast<uint4*>(orow + c + h * 16);
        op[0] = make_uint4(w[0], w[1], w[2], w[3]);
        op[1] = make_uint4(w[4], w[5], w[6], w[7]);
        if (p.gn_raw_out != nullptr) {       // second output: the raw tensor, straight from the parked bf16 pairs
          uint4* rp = reinterpret_cast<uint4*>(p.gn_raw_out + row_off + n_base + c + h * 16);
          rp[0] = make_uint4(pk[k * 16 + h * 8 + 0], pk[k * 16 + h * 8 + 1], pk[k * 16 + h * 8 + 2], pk[k * 16 + h * 8 + 3]);
          rp[1] = make_uint4(pk[k * 16 + h * 8 + 4], pk[k * 16 + h * 8 + 5], pk[k * 16 + h * 8 + 6], pk[k * 16 + h * 8 + 7]);
        }
      }
    }
  }
  epi_bar();                               // bias / gamma tables and the warp partials are rewritten by the next tile
}

// PAIR = true is a separate instantiation: a kernel that contains cta_group::2 instructions can only be launched as
// a cluster of 2 ("cluster misconfiguration" otherwise), so the single-CTA kernel must not contain them.
template <bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* ebias = reinterpret_cast<float*>(smem + (size_t)RING_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)RING_BYTES + EBIAS_FLOATS * 4);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* gn_bar = tmem_empty + 2;          // [2]: cluster exchange of GroupNorm sums, one per tile parity
  uint32_t* tmem_ptr_sh = reinterpret_cast<uint32_t*>(gn_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_units = p.dual ? (p.m_tiles + 1) / 2 : p.m_tiles;
  // tiles are handed out per cluster: group g = (m_group, n_tile); CTA `crank` of the cluster takes m-unit m_group * cluster + crank
  // (units past the end are processed as all-zero tiles so that every CTA keeps feeding its slice of the shared B tile)
  const int csize = p.cluster;
  const int crank = csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / csize, num_clusters = gridDim.x / csize;
  const int tiles_pp = ((m_units + csize - 1) / csize) * p.n_tiles;          // tile groups (per upsample phase)
  const int total_tiles = tiles_pp * (p.up_all ? 4 : 1);
  const int msize = p.mcast;                                                  // multicast group (1 = every CTA loads its own B tile)
  const uint16_t cmask = (uint16_t)((1u << msize) - 1u);
  const int nsub = p.dual ? 2 : 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s)
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.a_map[s]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.b_map) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], PAIR ? 1 : msize); }
      for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
      for (int s = 0; s < 2; ++s) mbar_init(&gn_bar[s], (uint32_t)csize);      // one arrival per CTA of the cluster
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // whole TMEM (512 columns): two 256-column fp32 accumulators; 1 CTA / SM by construction (smem)
    if constexpr (PAIR) {   // both CTAs of the pair issue the 2-SM allocation from the same warp, same smem destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_sh)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_sh)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();      // peers' mbarriers are initialised before any multicast / remote commit targets them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sh;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the
  // previous kernel's tail; its outputs (our A operand / bias tables) are visible after this point
  pdl_wait();
  pdl_launch_dependents();

  // ---- K loop as a sequence of "steps", one smem ring slot each -------------------------------------------------
  //  plain step: one 64-channel K-block of one filter tap: A tile(s) + one B tile                      (1 MMA group)
  //  slab step (p.slab, segment 0 = 3x3 stride-1 conv): for one (channel block, kw) the CTA loads the activation rows
  //    h0-1 .. h0+hb ONCE (hb+2 image rows, shifted by kw-1 along w; out-of-image rows/columns are zero-filled) plus
  //    the three weight tiles kh = 0,1,2; the three vertical taps are the same rows read at smem offsets of
  //    kh * (row pitch) -- whole 1024-byte swizzle atoms because a pixel row is W * 128 B                (3 MMA groups)
  //    => the activation bytes an SM must receive drop from 9 to 3*(hb+2)/hb tiles per channel block; the kernel is
  //    bound by the ~48 B/clk an SM can take in from L2 (measured), not by the MMA rate.
  const int STAGES = p.num_stages;               // ring depth / slot size chosen per launch
  const int STAGE_BYTES = p.stage_bytes;
  const uint32_t b_off = (uint32_t)p.a_region_bytes;          // B tile(s) follow the A region inside a slot
  const uint32_t b_cta_bytes = (uint32_t)(PAIR ? p.block_n / 2 : p.block_n) * BK * 2;   // B bytes landing in THIS CTA per weight tile
  const uint32_t b_tile_bytes = b_cta_bytes;                  // pitch of the (up to three) B tiles inside a slot
  const uint32_t row_pitch = (uint32_t)p.img_W * BK * 2;      // one image row of a 64-channel block in smem

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      auto load_b = [&](uint8_t* dst, uint64_t* bar, int kcol, int n_tile, int bz) {
        if constexpr (PAIR) {
          tma_load_3d_pair(&p.b_map, dst, bar, kcol, n_tile * p.block_n + crank * (p.block_n / 2), bz);
        } else if (msize == 1) {
          tma_load_3d(&p.b_map, dst, bar, kcol, n_tile * p.block_n, bz);
        } else {      // this CTA fetches its 1/msize slice of the B tile for the whole cluster
          const int rows_per = p.block_n / msize;
          tma_load_3d_mcast(&p.b_map, dst + (size_t)crank * rows_per * (BK * 2), bar, kcol, n_tile * p.block_n + crank * rows_per, bz, cmask);
        }
      };
      auto load_a = [&](const CUtensorMap* map, uint8_t* dst, uint64_t* bar, int c0, int cw, int ch, int cimg) {
        if constexpr (PAIR) tma_load_4d_pair(map, dst, bar, c0, cw, ch, cimg);
        else tma_load_4d(map, dst, bar, c0, cw, ch, cimg);
      };
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int ph = p.up_all ? tile / tiles_pp : p.up_phase;
        const int tpp = p.up_all ? tile - ph * tiles_pp : tile;
        const int m_group = tpp / p.n_tiles, n_tile = tpp - m_group * p.n_tiles, m_unit = m_group * csize + crank;
        int c1[2], c2[2], c3[2], bz = p.up_all ? ph : 0;
        for (int sub = 0; sub < nsub; ++sub) {
          const int m_tile = m_unit * nsub + sub;     // may be == m_tiles for an odd tail: TMA zero-fills, epilogue masks
          if (p.flat) {
            const int batch = m_tile / p.m_tiles_per_batch;
            c1[sub] = (m_tile - batch * p.m_tiles_per_batch) * BM; c2[sub] = 0;
            c3[sub] = p.a_batched ? batch : 0;
            if (sub == 0) bz = p.b_batched ? batch : 0;
          } else {
            c1[sub] = 0;
            c2[sub] = (m_tile % p.tiles_per_img) * p.h_box;
            c3[sub] = (m_tile / p.tiles_per_img) * p.imgs_per_tile;
          }
        }
        for (int seg = 0; seg < p.nseg; ++seg) {
          const int cblocks = p.seg_cblocks[seg], taps = p.seg_taps[seg];
          const int kb_base = p.seg_bk0[seg];           // K-block index (weight column / 64) where the segment starts
          if (p.slab && p.seg_slab[seg] >= 0) {
            const CUtensorMap* smap = &p.a_slab_map[p.seg_slab[seg]];
            const int nt = p.slab_taps;             // 3: 3x3 conv; 2: the 2x2 taps of upsample phase ph
            const uint32_t tx = (PAIR ? 2u : 1u) * ((uint32_t)p.slab_bytes + (uint32_t)nt * b_cta_bytes);
            // first slab row / column shift of tap column kw: (c2 - 1, kw - 1) for the 3x3 conv; the phase's taps read source rows
            // (ph >> 1) - 1 + {0, 1} and columns (ph & 1) - 1 + {0, 1} of the low-resolution image (same as the per-tap path below)
            const int r0 = nt == 3 ? c2[0] - 1 : c2[0] + (ph >> 1) - 1;
            const int w0 = nt == 3 ? -1 : (ph & 1) - 1;
            for (int cb = 0; cb < cblocks; ++cb)
              for (int kw = 0; kw < nt; ++kw) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
                if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[stage], tx);
                load_a(smap, sa, &full_bar[stage], cb * BK, w0 + kw, r0, c3[0]);
                for (int kh = 0; kh < nt; ++kh)
                  load_b(sa + b_off + kh * b_tile_bytes, &full_bar[stage], (kb_base + (kh * nt + kw) * cblocks + cb) * BK, n_tile, bz);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
          } else {
            // (tap row, tap column, channel block) advance as counters, no div/mod: this single thread issues every TMA
            // of the CTA, and ncu showed it busy while the MMA warp waited for data with the div/mod chain of the first version
            const int tside = taps == 9 ? 3 : (taps == 4 ? 2 : 1);
            const uint32_t tx = (PAIR ? 2u : 1u) * ((uint32_t)nsub * A_BYTES + b_cta_bytes);
            int kb = kb_base;
            for (int th = 0; th < tside; ++th)
              for (int tw = 0; tw < tside; ++tw) {
                int dh = 0, dw = 0;
                if (tside == 3) { dh = th - 1; dw = tw - 1; }
                else if (tside == 2) { dh = (ph >> 1) - 1 + th; dw = (ph & 1) - 1 + tw; }
                for (int cb = 0; cb < cblocks; ++cb, ++kb) {
                  mbar_wait(&empty_bar[stage], phase ^ 1);
                  uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
                  if (!PAIR || crank == 0) mbar_expect_tx(&full_bar[stage], tx);
                  for (int sub = 0; sub < nsub; ++sub) {
                    if (p.stride2)   // input coordinates of output row c2 / column 0 for tap (kh, kw): (2*c2 + kh, kw); index H / W is OOB -> 0
                      load_a(&p.a_map[seg], sa + sub * A_BYTES, &full_bar[stage], cb * BK, tw, 2 * c2[sub] + th, c3[sub]);
                    else
                      load_a(&p.a_map[seg], sa + sub * A_BYTES, &full_bar[stage], cb * BK, c1[sub] + dw, c2[sub] + dh, c3[sub]);
                  }
                  load_b(sa + b_off, &full_bar[stage], kb * BK, n_tile, bz);
                  if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
              }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && !(PAIR && crank != 0)) {
      // ===================== MMA issuer (pair mode: the leader CTA issues for both SMs) =====================
      // instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major both, N>>3 @17, M>>4 @24
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)((PAIR ? 2 * BM : BM) >> 4) << 24);
      const uint32_t idesc_swap = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * BM >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * MAX_BN;
        uint32_t accumulate = 0;
        for (int seg = 0; seg < p.nseg; ++seg) {
          const bool slab = p.slab && p.seg_slab[seg] >= 0;
          const int nsteps = slab ? p.seg_cblocks[seg] * p.slab_taps : p.seg_taps[seg] * p.seg_cblocks[seg];
          const int groups = slab ? p.slab_taps : 1;
          for (int st = 0; st < nsteps; ++st) {
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
            for (int g = 0; g < groups; ++g) {
              const uint64_t b_desc = umma_desc_sw128(sa + b_off + (uint32_t)g * b_tile_bytes);
              if (p.swap) {
                // weights are the M = 128 operand, the two adjacent pixel tiles (contiguous in the slot: 256 rows) the N = 256 one
                const uint64_t x_desc = umma_desc_sw128(slab ? sa + (uint32_t)g * row_pitch : sa);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  // pair: M = 256 channels (each CTA feeds the 128 weight rows it already loads as "its half of B"),
                  // N = 256 pixels (each CTA feeds its own m-tile) -- the same bytes in shared memory, roles exchanged
                  if constexpr (PAIR) umma_bf16_pair(d_tmem, b_desc + (uint64_t)(k * 2), x_desc + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
                  else umma_bf16(d_tmem, b_desc + (uint64_t)(k * 2), x_desc + (uint64_t)(k * 2), idesc_swap, accumulate | (uint32_t)k);
                }
                accumulate = 1;
                continue;
              }
              if (!PAIR && nsub == 2 && p.block_n <= 64) {
                // narrow N (the 3-channel output conv): a 128 x 16 x 16 MMA is all latency, and the four K steps of one
                // sub-tile are a dependent chain on the same TMEM columns (ncu: one UTCHMMA per ~120 clocks).  Alternate the
                // two sub-tiles' independent chains instead of running them back to back; per-accumulator order is unchanged.
                const uint64_t a0 = umma_desc_sw128(slab ? sa + (uint32_t)g * row_pitch : sa);
                const uint64_t a1 = umma_desc_sw128(slab ? sa + (uint32_t)(g + p.h_box) * row_pitch : sa + (uint32_t)A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                  umma_bf16(d_tmem, a0 + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
                  umma_bf16(d_tmem + 128u, a1 + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
                }
                accumulate = 1;
                continue;
              }
              for (int sub = 0; sub < nsub; ++sub) {
                // slab: vertical tap g of sub-tile `sub` = the slab rows starting (g + sub * h_box) image rows in
                const uint32_t a_addr = slab ? sa + (uint32_t)(g + sub * p.h_box) * row_pitch : sa + (uint32_t)sub * A_BYTES;
                const uint64_t a_desc = umma_desc_sw128(a_addr);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {   // +32 B per UMMA_K inside the swizzle atom
                  if constexpr (PAIR) umma_bf16_pair(d_tmem + (uint32_t)sub * 128u, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
                  else umma_bf16(d_tmem + (uint32_t)sub * 128u, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, accumulate | (uint32_t)k);
                }
              }
              accumulate = 1;
            }
            // frees the smem slot once these MMAs retire (in every CTA whose TMA writes land here)
            if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
            else if (msize == 1) umma_commit(&empty_bar[stage]);
            else umma_commit_mcast(&empty_bar[stage], cmask);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        if constexpr (PAIR) umma_commit_pair(&tmem_full[acc]);   // each CTA's epilogue drains its own 128 TMEM lanes
        else umma_commit(&tmem_full[acc]);                        // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;             // which alternate 32-column groups this warp handles
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t gn_it = 0;                            // tiles this CTA has normalised (parity selects the exchange buffer / barrier)
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int ph = p.up_all ? tile / tiles_pp : p.up_phase;
      const int tpp = p.up_all ? tile - ph * tiles_pp : tile;
      const int m_group = tpp / p.n_tiles, n_tile = tpp - m_group * p.n_tiles, m_unit = m_group * csize + crank;
      const int stats_slot0 = p.up_all ? ph * p.tiles_per_img : p.stats_slot0;
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const int row = q * 32 + lane;
      const int n_base = n_tile * p.block_n;
      if (p.swap) {
        // ---- swapped tile: TMEM lane = output channel, column = pixel (256 pixels = the unit's two m-tiles, one image).
        // bias / time-embedding bias are one register per thread, GroupNorm channel sums are thread-local running sums
        // (no shuffles); a lane-pair exchange packs adjacent channels so every store instruction writes two 64-byte
        // runs of the NHWC row (even lanes pixel P, odd lanes pixel P + 1).
        const int et = threadIdx.x - 64;
        const int ch_base = n_base + (PAIR ? crank * BM : 0);       // pair: this CTA's accumulator holds channels [crank*128, +128)
        const int ch = ch_base + row;
        const bool ch_ok = ch < p.N_out;
        const int m_tile0 = PAIR ? m_group * 2 : m_unit * 2;        // pair: columns 0..127 = CTA 0's m-tile, 128..255 = CTA 1's
        const size_t pix0 = (size_t)m_tile0 * BM;
        float bv = 0.f;
        if (ch_ok) {
          if (p.bias) bv = p.bias[ch];
          if (p.rowbias) bv += p.rowbias[(size_t)min((int)(pix0 / p.HW), (p.M_total - 1) / p.HW) * p.rowbias_ld + ch];
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * MAX_BN;
        const bool odd = (lane & 1) != 0;
        __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out) + (ch & ~1);
        const bool pair_ok = (ch | 1) < p.N_out;
        const bool do_stats = p.stats_out != nullptr;
        const bool split = (p.flags & SD_GEMM_SPLIT3) != 0;     // out rows are [hi(N) | lo(N)]: the lo half is stored N_out channels further
        if (p.gn_gamma != nullptr) {
          // ---- fused GroupNorm(32 groups) + swish: the conv output feeds ONLY act(normalize(h)) (cifar/models/layers.py:552-558), so
          // the raw tensor is never written.  ONE pass over the accumulator: + bias, per-channel sums (thread = channel: no
          // shuffles), and the 128 values of this thread are parked as 64 packed bf16 pairs in registers -- the same rounding the
          // separate GroupNorm pass sees when it re-reads the raw bf16 tensor -- so the accumulator is released to the MMA warp
          // BEFORE the statistics are complete.  (A first version re-read TMEM for the second pass: tcgen05.ld competes with the
          // running MMAs for TMEM bandwidth and the launches became epilogue-bound, 185 -> 206 us at 16x16 / K = 4608.)
          // The sums of an image's other tiles (32x32: four units) would come from the cluster peers through distributed shared
          // memory (SDB_GN_FUSE=2, see launch_gemm); at 16x16 the unit IS the image.
          float* gsum = ebias + MAX_BN;            // [chalf][which][128]
          float* ctot = gsum + 4 * BM;             // [which][128]: channel totals over the image
          float* xsum = ctot + 2 * BM;             // [parity][which][128]: this CTA's channel sums, read by its cluster peers
          const bool xchg = !PAIR && csize > 1;
          // several images per unit (8x8: HW = 64, a 256-pixel unit holds four): the 64-column chunk `it` IS image `it` -- this
          // thread holds 32 of its pixels, the partner warp (other chalf) the other 32 -- so bias, sums and statistics are per chunk
          const bool mi = p.imgs_per_tile > 1;
          float bvi[4] = {bv, bv, bv, bv};
          if (mi && ch_ok && p.rowbias) {
            const int img0 = (int)(pix0 / p.HW), img_last = (p.M_total - 1) / p.HW;
            const float b0 = p.bias ? p.bias[ch] : 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it) bvi[it] = b0 + p.rowbias[(size_t)min(img0 + it, img_last) * p.rowbias_ld + ch];
          }
          uint32_t pk[64];
          float ps[4], pq[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int c = chalf * 32 + it * 64;
            uint32_t r[2][16];
            tmem_ld16(taddr + c, r[0]);
            tmem_ld16(taddr + c + 16, r[1]);
            tmem_wait_ld();
            float ssum = 0.f, ssq = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float x0 = __uint_as_float(r[h][2 * jj]) + bvi[it], x1 = __uint_as_float(r[h][2 * jj + 1]) + bvi[it];
                ssum += x0 + x1;
                ssq = fmaf(x0, x0, fmaf(x1, x1, ssq));
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
                pk[it * 16 + h * 8 + jj] = *reinterpret_cast<const uint32_t*>(&h2);
              }
            ps[it] = ssum; pq[it] = ssq;
          }
          tcgen05_fence_before();                  // accumulator drained: hand it back to the MMA warp now
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_rank(&tmem_empty[acc], 0);
            else mbar_arrive(&tmem_empty[acc]);
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
          float rs4[4], sh4[4];
          const int gw = et / BM, gn_n = et - gw * BM;                   // threads 0..255: (which, channel)
          float unit_tot = 0.f;                                          // this unit's channel total (raw stats; single-image units)
          if (mi) {
            float* gmi = ebias;                    // [image][chalf][which][128]
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              gmi[((it * 2 + chalf) * 2 + 0) * BM + row] = ps[it];
              gmi[((it * 2 + chalf) * 2 + 1) * BM + row] = pq[it];
            }
            const float gam_m = ch_ok ? p.gn_gamma[ch] : 0.f, bet_m = ch_ok ? p.gn_beta[ch] : 0.f;    // issued ahead of the barrier
            epi_bar();
            const int cpg_m = p.N_out >> 5, g0_m = (row / cpg_m) * cpg_m;
            const float ginv_m = 1.f / ((float)p.HW * (float)cpg_m);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              float gs = 0.f, gq = 0.f;
              for (int cc = 0; cc < cpg_m; ++cc) {
                gs += gmi[((it * 2 + 0) * 2 + 0) * BM + g0_m + cc] + gmi[((it * 2 + 1) * 2 + 0) * BM + g0_m + cc];
                gq += gmi[((it * 2 + 0) * 2 + 1) * BM + g0_m + cc] + gmi[((it * 2 + 1) * 2 + 1) * BM + g0_m + cc];
              }
              const float gmean = gs * ginv_m;
              const float gvar = fmaxf(fmaf(gq, ginv_m, -gmean * gmean), 0.f);
              rs4[it] = rsqrtf(gvar + p.gn_eps) * gam_m;
              sh4[it] = bet_m - gmean * rs4[it];
            }
            epi_bar();                             // every thread has read the sums before the next unit overwrites them
          } else {
          gsum[(chalf * 2 + 0) * BM + row] = (ps[0] + ps[1]) + (ps[2] + ps[3]);
          gsum[(chalf * 2 + 1) * BM + row] = (pq[0] + pq[1]) + (pq[2] + pq[3]);
          epi_bar();
          const int gbuf = (int)(gn_it & 1u);
          unit_tot = gsum[(0 * 2 + gw) * BM + gn_n] + gsum[(1 * 2 + gw) * BM + gn_n];
          int* gx_done = nullptr;
          if (xchg) {
            xsum[(gbuf * 2 + gw) * BM + gn_n] = gsum[(0 * 2 + gw) * BM + gn_n] + gsum[(1 * 2 + gw) * BM + gn_n];
            epi_bar();                                                    // every channel sum of this CTA is written ...
            if (et == 0)
              for (int r = 0; r < csize; ++r) mbar_arrive_rank(&gn_bar[gbuf], (uint32_t)r);     // ... and released to the cluster
            mbar_wait_cluster(&gn_bar[gbuf], (gn_it >> 1) & 1u);
            float t = 0.f;
            for (int r = 0; r < csize; ++r) t += ld_dsmem_f32(&xsum[(gbuf * 2 + gw) * BM + gn_n], (uint32_t)r);    // fixed order
            ctot[gw * BM + gn_n] = t;
          } else if (p.gx_units > 1) {
            // exchange through global memory: the image's gx_units CTAs are consecutive blocks in the same tile round (the grid is
            // a multiple of gx_units and co-resident: cooperative launch); a group alternates between two slots, so a slot is
            // reused only after its last reader has zeroed the counters (that reader arrives at the next round after the reset)
            const int U = p.gx_units, u = (int)blockIdx.x % U;
            const int slot = ((int)blockIdx.x / U) * 2 + gbuf;
            float* gd = p.gx_data + (size_t)slot * U * 2 * BM;
            int* cnt = p.gx_cnt + slot * 2;
            gd[u * 2 * BM + et] = unit_tot;                                 // et == gw * 128 + gn_n
            epi_bar();
            if (et == 0) {
              __threadfence();                                             // the CTA's sums (ordered by the barrier) before the arrival
              atomicAdd(cnt, 1);
              const long long t0 = clock64();
              while (*reinterpret_cast<volatile int*>(cnt) < U) {
                __nanosleep(64);
                if (clock64() - t0 > 4000000000LL) __trap();              // peers not resident: a launch failure, never a hung GPU
              }
              __threadfence();
            }
            epi_bar();
            float t = 0.f;
            for (int r = 0; r < U; ++r) t += __ldcg(gd + r * 2 * BM + et);  // fixed order; L2 (the peers' writes are not in this SM's L1)
            ctot[gw * BM + gn_n] = t;
            gx_done = cnt + 1;
          } else {
            ctot[gw * BM + gn_n] = gsum[(0 * 2 + gw) * BM + gn_n] + gsum[(1 * 2 + gw) * BM + gn_n];
          }
          ++gn_it;
          const float gam = ch_ok ? p.gn_gamma[ch] : 0.f, bet = ch_ok ? p.gn_beta[ch] : 0.f;    // issued ahead of the barrier
          epi_bar();
          if (gx_done != nullptr && et == 0) {                             // every thread of this CTA has read the slot
            if (atomicAdd(gx_done, 1) == p.gx_units - 1) {
              gx_done[-1] = 0; gx_done[0] = 0;
              __threadfence();
            }
          }
          // group statistics in fp32, every thread for its own channel's group (a handful of shared-memory reads and FMAs).
          // fp64 here -- first per thread, then one thread per group behind a barrier -- put the fp64 division / rsqrt
          // subroutines on the epilogue's critical path: ncu showed 40-55 % of all stall samples at that barrier and the tensor
          // pipe fell from 93 % to 77 % at 16x16 / K = 4608.  flax's GroupNorm computes E[x^2] - E[x]^2 in fp32 as well.
          const int cpg = p.N_out >> 5;                                   // channels per group (32 groups); divides 128
          const int g0 = (row / cpg) * cpg;
          float gs = 0.f, gq = 0.f;
          for (int cc = 0; cc < cpg; ++cc) { gs += ctot[g0 + cc]; gq += ctot[BM + g0 + cc]; }
          const float ginv = 1.f / ((float)p.HW * (float)cpg);
          const float gmean = gs * ginv;
          const float gvar = fmaxf(fmaf(gq, ginv, -gmean * gmean), 0.f);  // flax: E[x^2] - E[x]^2, clipped at 0
          const float rs = rsqrtf(gvar + p.gn_eps) * gam;
          const float sh = bet - gmean * rs;
#pragma unroll
          for (int it = 0; it < 4; ++it) { rs4[it] = rs; sh4[it] = sh; }
          }      // !mi
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int c = chalf * 32 + it * 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[16];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const uint32_t w = pk[it * 16 + h * 8 + jj];
                const float y0 = fmaf(__uint_as_float(w << 16), rs4[it], sh4[it]), y1 = fmaf(__uint_as_float(w & 0xFFFF0000u), rs4[it], sh4[it]);
                v[2 * jj] = p.gn_swish ? swish_tanh_f(y0) : y0;
                v[2 * jj + 1] = p.gn_swish ? swish_tanh_f(y1) : y1;
              }
              __nv_bfloat16* orow = obase + (pix0 + (size_t)(c + h * 16 + (odd ? 1 : 0))) * (size_t)p.out_ld;
              const bool ok = pair_ok && pix0 + (size_t)(c + h * 16 + 15) < (size_t)p.M_total;
              const size_t ld2 = 2 * (size_t)p.out_ld;
              // exchange first (every lane takes part), then one branch around the eight stores (see the plain epilogue below)
              uint32_t wn[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[2 * j] : v[2 * j + 1], 1);
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(odd ? recv : v[2 * j], odd ? v[2 * j + 1] : recv);
                wn[j] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              if (ok) {
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint32_t*>(orow + (size_t)j * ld2) = wn[j];
              }
              if (p.gn_raw_out != nullptr) {
                // second output: the raw tensor (already bf16 in the parked registers).  A parked word is (pixel j, pixel j+1) of
                // this thread's channel; a stored word is (channel even, channel odd) of one pixel: swap halves with the lane partner.
                __nv_bfloat16* rrow = p.gn_raw_out + (ch & ~1) + (pix0 + (size_t)(c + h * 16 + (odd ? 1 : 0))) * (size_t)p.out_ld;
                uint32_t wr[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const uint32_t mine = pk[it * 16 + h * 8 + jj];
                  const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, 1);
                  wr[jj] = odd ? __byte_perm(mine, other, 0x3276) : __byte_perm(mine, other, 0x5410);
                }
                if (ok) {
#pragma unroll
                  for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint32_t*>(rrow + (size_t)jj * ld2) = wr[jj];
                }
              }
            }
          }
          if (p.gn_raw_out != nullptr && p.stats_out != nullptr) {
            // channel sums of the raw tensor for later GroupNorms over it (skip connections): the unit's totals go to its first
            // 128-pixel tile's slot, zeros to the second's (consumers add the slots of an image)
            const int m_tile = m_tile0;
            if (m_tile < p.m_tiles && ch_base + gn_n < p.N_out) {
              const float t = unit_tot;
              const size_t slot = (size_t)(m_tile / p.tiles_per_img) * p.stats_tpi_total + stats_slot0 + (m_tile % p.tiles_per_img);
              p.stats_out[(slot * 2 + gw) * p.N_out + ch_base + gn_n] = t;
              if (m_tile + 1 < p.m_tiles) p.stats_out[((slot + 1) * 2 + gw) * p.N_out + ch_base + gn_n] = 0.f;
            }
          }
          continue;
        }
        float* wstat = ebias + MAX_BN;                          // [chalf][sub][which][128]
        float ssum = 0.f, ssq = 0.f;
        if (!(p.flags & 0x100u)) {      // 0x100: timing probe (tools/gemm_probe.py) -- release the accumulator without draining it
          // four 32-column groups per warp (columns chalf*32 + it*64); groups 0,1 belong to the unit's first m-tile, 2,3 to
          // the second.  Rolled on purpose: unrolled, the epilogue was ~4000 instructions (64 KB of SASS) and ncu showed 17 %
          // of all stall samples as instruction-cache misses (stall_no_inst).
#pragma unroll 1
          for (int it = 0; it < 4; ++it) {
            const int c = chalf * 32 + it * 64;
            uint32_t r[2][16];
            tmem_ld16(taddr + c, r[0]);
            tmem_ld16(taddr + c + 16, r[1]);
            tmem_wait_ld();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[h][j]) + bv;
              if (do_stats) {                    // uniform: 1x1 projections and shortcut-free layers carry no GroupNorm sums
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  ssum += v[j];
                  ssq = fmaf(v[j], v[j], ssq);
                }
              }
              // first pixel of this lane (even lanes take pixel P, odd lanes P + 1) and the address step between consecutive pixels.
              // Upsample phase (a, b) of the fused upsample + conv: low-resolution pixel (img, i, j) lands on (2i + a, 2j + b) of
              // the 2H x 2W output, so a 16-pixel group (one low-resolution row or a part of one: img_W % 16 == 0) is a run of
              // output pixels with stride 2.
              size_t pstep = (size_t)p.out_ld;
              __nv_bfloat16* orow;
              if (p.up_phase >= 0) {
                size_t off;
                row_offset(p, 0, (int)(pix0 + (size_t)(c + h * 16 + (odd ? 1 : 0))), off, ph);       // m_tile 0 + row = the pixel index
                orow = obase + off;
                pstep = 2 * (size_t)p.out_ld;
              } else {
                orow = obase + (pix0 + (size_t)(c + h * 16 + (odd ? 1 : 0))) * (size_t)p.out_ld;
              }
              const bool ok = pair_ok && pix0 + (size_t)(c + h * 16 + 15) < (size_t)p.M_total;    // tiles are whole (HW % 256 == 0)
              // The exchange first (every lane takes part), then ONE branch around the eight stores: with the test inside the
              // loop the compiler emitted a branch + reconvergence pair per store (ncu: 13.5 instructions per element, 2.6 of
              // them control flow, in launches whose four K-blocks cannot hide this epilogue).
              float e0[8], e1[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float recv = __shfl_xor_sync(0xffffffffu, odd ? v[2 * j] : v[2 * j + 1], 1);
                e0[j] = odd ? recv : v[2 * j];                                          // channels (ch & ~1, ch | 1) of one pixel
                e1[j] = odd ? v[2 * j + 1] : recv;
              }
              if (ok) {
                const size_t ld2 = 2 * pstep;
                if (!split) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)j * ld2) = __floats2bfloat162_rn(e0[j], e1[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0[j], e1[j]);
                    *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)j * ld2) = h2;
                    *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)j * ld2 + p.N_out) =
                        __floats2bfloat162_rn(e0[j] - __low2float(h2), e1[j] - __high2float(h2));
                  }
                }
              }
            }
            if (do_stats && (it & 1)) {
              const int sub = it >> 1;
              wstat[((chalf * 2 + sub) * 2 + 0) * BM + row] = ssum;
              wstat[((chalf * 2 + sub) * 2 + 1) * BM + row] = ssq;
              ssum = 0.f; ssq = 0.f;
            }
          }
        }
        if (do_stats) {
          epi_bar();
          for (int i = et; i < 4 * BM; i += EPI_THREADS) {
            const int sub = i / (2 * BM), which = (i / BM) & 1, n = i & (BM - 1);
            const int m_tile = m_tile0 + sub;
            if (m_tile < p.m_tiles && ch_base + n < p.N_out) {
              const float t = wstat[((0 * 2 + sub) * 2 + which) * BM + n] + wstat[((1 * 2 + sub) * 2 + which) * BM + n];
              const size_t slot = (size_t)(m_tile / p.tiles_per_img) * p.stats_tpi_total + stats_slot0 + (m_tile % p.tiles_per_img);
              p.stats_out[(slot * 2 + which) * p.N_out + ch_base + n] = t;
            }
          }
          epi_bar();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_rank(&tmem_empty[acc], 0);
          else mbar_arrive(&tmem_empty[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
        continue;
      }
      bool released = false;                       // the fused-GroupNorm path hands the accumulator back itself (early)
      for (int sub = 0; sub < nsub; ++sub) {
        const int m_tile = m_unit * nsub + sub;
        size_t row_off;
        const bool row_ok = row_offset(p, m_tile, row, row_off, ph);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * MAX_BN + (uint32_t)sub * 128u;
        if (p.flags & SD_EPI_SOFTMAX) {
          // the whole score row (block_n == N <= 256 columns) sits in one TMEM lane: softmax over the row's diagonal block
          // [lo, hi) (several small images share one 128-row tile), zeros elsewhere.  The two warps of a lane quarter take
          // alternate 32-column groups and exchange the row max / row sum through shared memory (two named barriers).
          const int rl = (p.flat ? (m_tile % p.m_tiles_per_batch) : 0) * BM + row;
          const int lo = (rl / p.softmax_block) * p.softmax_block, hi = lo + p.softmax_block;
          const float sl2 = p.softmax_scale * 1.4426950408889634f;       // exp(s*x) = exp2(s*log2(e)*x): one FFMA + MUFU.EX2
          float* xch = ebias;                                             // [128 rows][2 halves]
          float mx = -INFINITY;
          for (int c = chalf * 32; c < p.block_n; c += 64) {
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr + c, r0);
            tmem_ld16(taddr + c + 16, r1);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c + j >= lo && c + j < hi) mx = fmaxf(mx, __uint_as_float(r0[j]));
              if (c + 16 + j >= lo && c + 16 + j < hi) mx = fmaxf(mx, __uint_as_float(r1[j]));
            }
          }
          xch[row * 2 + chalf] = mx;
          epi_bar();
          mx = fmaxf(xch[row * 2], xch[row * 2 + 1]);
          epi_bar();
          const float mxs = mx * sl2;
          float sum = 0.f;
          for (int c = chalf * 32; c < p.block_n; c += 64) {
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr + c, r0);
            tmem_ld16(taddr + c + 16, r1);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (c + j >= lo && c + j < hi) sum += exp2f(fmaf(__uint_as_float(r0[j]), sl2, -mxs));
              if (c + 16 + j >= lo && c + 16 + j < hi) sum += exp2f(fmaf(__uint_as_float(r1[j]), sl2, -mxs));
            }
          }
          xch[row * 2 + chalf] = sum;
          epi_bar();
          const float inv = 1.f / (xch[row * 2] + xch[row * 2 + 1]);
          for (int c = chalf * 32; c < p.block_n; c += 64) {
            uint32_t r0[16], r1[16];
            tmem_ld16(taddr + c, r0);
            tmem_ld16(taddr + c + 16, r1);
            tmem_wait_ld();
            uint32_t w[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c0 = c + 2 * j, c1 = c + 16 + 2 * j;
              const float e0 = (c0 >= lo && c0 < hi) ? exp2f(fmaf(__uint_as_float(r0[2 * j]), sl2, -mxs)) * inv : 0.f;
              const float e1 = (c0 + 1 >= lo && c0 + 1 < hi) ? exp2f(fmaf(__uint_as_float(r0[2 * j + 1]), sl2, -mxs)) * inv : 0.f;
              const float e2 = (c1 >= lo && c1 < hi) ? exp2f(fmaf(__uint_as_float(r1[2 * j]), sl2, -mxs)) * inv : 0.f;
              const float e3 = (c1 + 1 >= lo && c1 + 1 < hi) ? exp2f(fmaf(__uint_as_float(r1[2 * j + 1]), sl2, -mxs)) * inv : 0.f;
              const __nv_bfloat162 ha = __floats2bfloat162_rn(e0, e1), hb = __floats2bfloat162_rn(e2, e3);
              w[j] = *reinterpret_cast<const uint32_t*>(&ha);
              w[8 + j] = *reinterpret_cast<const uint32_t*>(&hb);
            }
            if (row_ok) {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + c);
              op[0] = make_uint4(w[0], w[1], w[2], w[3]);
              op[1] = make_uint4(w[4], w[5], w[6], w[7]);
              op[2] = make_uint4(w[8], w[9], w[10], w[11]);
              op[3] = make_uint4(w[12], w[13], w[14], w[15]);
            }
          }
          epi_bar();                                  // the exchange area is reused by the next tile
          continue;
        }
        // ---- phase 0: cooperative, latency-tolerant loads.  The first version read bias / row bias / residual from
        // global memory per 16-column chunk; with one epilogue warp per scheduler the L2 latency was fully exposed
        // (ncu: 35-45 % of the kernel's stall samples) and a 128x256 tile cost ~17k cycles, more than its MMAs for
        // K <= 2304.  Now: (bias + row bias) go into a small smem table (one named barrier) and phase 1 only touches
        // TMEM + smem; residual adds of the score-net ride the MMA as an identity-weight K segment fed by TMA.
        const int et = threadIdx.x - 64;                                   // 0..EPI_THREADS-1 among the epilogue warps
        const bool use_tab = p.bias != nullptr || p.rowbias != nullptr;
        float* eb = ebias;
        const int vcols = min(p.block_n, p.N_out - n_base);
        if (use_tab) {
          const int img0 = p.flat ? m_tile / p.m_tiles_per_batch : (m_tile * BM) / p.HW;
          const int img_last = p.flat ? img0 : (p.M_total - 1) / p.HW;
          for (int k = 0; k < p.imgs_in_tile; ++k) {
            const int im = min(img0 + k, img_last);
            for (int n = et; n < p.block_n; n += EPI_THREADS) {
              float v = 0.f;
              if (n < vcols) {
                if (p.bias) v = p.bias[n_base + n];
                if (p.rowbias) v += p.rowbias[(size_t)im * p.rowbias_ld + n_base + n];
              }
              eb[k * MAX_BN + n] = v;
            }
          }
        }
        if (use_tab) epi_bar();
        // ---- phase 1: this thread's row
        const float* eb_row = eb + ((p.imgs_in_tile > 1) ? (row / p.HW) * MAX_BN : 0);
        if (p.gn_gamma != nullptr) {
          // fused GroupNorm + swish, thread = pixel-row form (host: one sub-tile, N = 256, whole 8x8 / 4x4 images per tile)
          float* gtab = ebias + 8 * MAX_BN;
          float* wred = gtab + 2 * MAX_BN;
          const int warp_e = warp - 2;
          // pairs run 256-column tiles (or 128 below 148 tiles), single CTAs reach this path with 128-column tiles (see launch_gemm)
          if constexpr (PAIR) {
            if (p.block_n == 128) {              // halved-N pairs (4x4 level below 148 tiles)
              if (p.HW == 64) gn_rows_epilogue<2, 64, true>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
              else gn_rows_epilogue<2, 16, true>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
            } else if (p.HW == 64) gn_rows_epilogue<4, 64, true>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
            else gn_rows_epilogue<4, 16, true>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
          } else {
            if (p.HW == 64) gn_rows_epilogue<2, 64, false>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
            else gn_rows_epilogue<2, 16, false>(p, taddr, row, q, chalf, lane, warp_e, et, eb_row, gtab, wred, row_ok, row_off, n_base, &tmem_empty[acc]);
          }
          released = true;
          continue;
        }
        const bool do_stats = p.stats_out != nullptr;         // host guarantees one image per tile and N_out % 16 == 0
        float* wstat = ebias + MAX_BN;                          // [4 warps][2][MAX_BN]
        for (int c = chalf * 32; c < p.block_n; c += 64) {
          uint32_t r0[16], r1[16];
          tmem_ld16(taddr + c, r0);
          const bool second = (c + 16 < p.block_n);
          if (second) tmem_ld16(taddr + c + 16, r1);
          tmem_wait_ld();
          float v0[16], v1[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { v0[j] = 0.f; v1[j] = 0.f; }
          if (row_ok) {
            if (c < vcols) epilogue_store16(p, r0, row_off, n_base + c, use_tab ? eb_row + c : nullptr, v0);
            if (second && c + 16 < vcols) epilogue_store16(p, r1, row_off, n_base + c + 16, use_tab ? eb_row + c + 16 : nullptr, v1);
          }
          if (do_stats) {
            if (c < vcols) wstat[(q * 2 + (lane >> 4)) * MAX_BN + c + (lane & 15)] = warp_colsum16(v0, lane);
            if (second && c + 16 < vcols) wstat[(q * 2 + (lane >> 4)) * MAX_BN + c + 16 + (lane & 15)] = warp_colsum16(v1, lane);
          }
        }
        if (do_stats) {
          epi_bar();
          if (m_tile < p.m_tiles)
            for (int i = et; i < 2 * vcols; i += EPI_THREADS) {
              const int which = i >= vcols ? 1 : 0, n = i - which * vcols;
              const float t = wstat[(0 * 2 + which) * MAX_BN + n] + wstat[(1 * 2 + which) * MAX_BN + n] +
                              wstat[(2 * 2 + which) * MAX_BN + n] + wstat[(3 * 2 + which) * MAX_BN + n];
              const size_t slot = (size_t)(m_tile / p.tiles_per_img) * p.stats_tpi_total + stats_slot0 + (m_tile % p.tiles_per_img);
              p.stats_out[(slot * 2 + which) * p.N_out + n_base + n] = t;
            }
        }
        if (use_tab || do_stats) epi_bar();   // table / partials are rewritten by the next (sub-)tile
      }
      if (!released) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_rank(&tmem_empty[acc], 0);     // the leader's MMA warp waits for both CTAs' epilogues
          else mbar_arrive(&tmem_empty[acc]);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();      // no CTA exits while a peer may still multicast into its smem or signal its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------


// SDB_GN_FUSE, overridable at run time (sd_set_gn_fuse: bench.py measures the same launches with and without the fused epilogue)
static std::atomic<int> g_gn_fuse{-1};
static int gn_fuse_level() {
  int v = g_gn_fuse.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SDB_GN_FUSE");
    v = e ? atoi(e) : 2;
    g_gn_fuse.store(v, std::memory_order_relaxed);
  }
  return v;
}

// Exchange area of the global-memory GroupNorm exchange (GemmParams::gx_units): one region per stream that issues such launches
// (launches on one stream are serialised, so a region has one user at a time), 2 slots per group of CTAs.
constexpr int GX_REGIONS = 8, GX_MAX_GROUPS = 128, GX_MAX_UNITS = 4;
__device__ float g_gx_data[GX_REGIONS * GX_MAX_GROUPS * 2 * GX_MAX_UNITS * 2 * BM];
__device__ int g_gx_cnt[GX_REGIONS * GX_MAX_GROUPS * 2 * 2];

// -> region index of `st`, or -1 when more than GX_REGIONS distinct streams have asked (the launch then uses clusters)
static int gx_region_of(cudaStream_t st) {
  static std::mutex mu;
  static cudaStream_t seen[GX_REGIONS];
  static int nseen = 0;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < nseen; ++i)
    if (seen[i] == st) return i;
  if (nseen == GX_REGIONS) return -1;
  seen[nseen] = st;
  return nseen++;
}

static int launch_gemm(GemmParams& p, int N, long K, const void* Wt, int ldb, long long strideB, int nbatchB,
                       const float* bias, const float* rowbias, int rowbias_ld, const void* residual, unsigned flags,
                       void* out, int out_ld, cudaStream_t st, const char* who, float* stats_out = nullptr) {
  if (flags & SD_GEMM_SPLIT3) {
    if (flags & SD_EPI_SOFTMAX) return fail(kErrInvalidArg, std::string(who) + ": no softmax epilogue in split precision (use sd_softmax_rows_split)");
    if (!(flags & SD_EPI_OUT_F32) && ((N % 16) != 0 || out_ld < 2 * N))
      return fail(kErrInvalidArg, std::string(who) + ": split output needs N % 16 == 0 and out_ld >= 2N");
  }
  const int n_pad = (N + 15) / 16 * 16;
  p.block_n = n_pad < MAX_BN ? n_pad : MAX_BN;
  // few tiles (low-resolution layers): halve the N tile so more of the 148 SMs get work
  const int phases = p.up_all ? 4 : 1;       // independent tile sets in one launch
  if (!(flags & SD_EPI_SOFTMAX) && p.block_n > 128 && phases * p.m_tiles * ((n_pad + p.block_n - 1) / p.block_n) < num_sms()) p.block_n = 128;
  // 1x1 layers (NIN, attention projections; K = 256..512) are bound by the epilogue and the output write, not by the MMAs:
  // run them as N = 128 column blocks so they take the operand-swapped path below (64-byte store runs, thread-local GN sums)
  static const int want_swap = [] { const char* e = getenv("SDB_GEMM_SWAP"); return e ? atoi(e) : 1; }();   // tuning knob
  bool all_1tap = !p.flat;
  for (int sgi = 0; sgi < p.nseg; ++sgi) all_1tap = all_1tap && p.seg_taps[sgi] == 1;
  // upsample phases: the strided-pixel store of the swapped epilogue needs a 16-pixel group inside one low-resolution row
  static const int want_swap_up = [] { const char* e = getenv("SDB_GEMM_SWAP_UP"); return e ? atoi(e) : 1; }();   // tuning knob
  const bool up_ok = p.up_phase < 0 || (want_swap_up && (p.img_W % 16) == 0 && N == MAX_BN);
  static const int want_swap_s2 = [] { const char* e = getenv("SDB_GEMM_SWAP_S2"); return e ? atoi(e) : 1; }();   // tuning knob: stride-2 convs swapped too
  // several images per 128-pixel tile (8x8, 4x4): a 256-pixel unit spans images, which the swapped epilogue allows when nothing in
  // it is per image -- no time-embedding row bias, no per-tile channel sums, no fused GroupNorm (those launches keep the
  // thread = pixel-row form, whose fused epilogue holds whole images per tile)
  static const int want_swap_multi = [] { const char* e = getenv("SDB_GEMM_SWAP_MULTI"); return e ? atoi(e) : 1; }();   // tuning knob
  // ... or, at 8x8 (a 256-pixel unit = four whole images, one per 64-column chunk of the accumulator), the fused GroupNorm epilogue
  // with per-chunk bias and statistics (the conditions repeat the fuse test further down: a swapped multi-image launch with a
  // GroupNorm or a row bias must end up fused)
  const bool mi_fuse_ok = want_swap_multi >= 1 && p.gn_gamma != nullptr && gn_fuse_level() >= 1 && !(flags & SD_GEMM_SPLIT3) && p.num_kb >= 16 &&
                          p.HW == 64 && !stats_out && p.up_phase < 0 && !p.stride2 && (N % 32) == 0 && (BM % (N / 32)) == 0 && (N % 128) == 0 &&
                          getenv("SDB_GN_MULTI_SWAP_OFF") == nullptr;
  const bool unit_ok = p.imgs_per_tile == 1 ? (p.tiles_per_img % 2) == 0
                                            : ((want_swap_multi && !rowbias && !stats_out && p.gn_gamma == nullptr && p.up_phase < 0) || mi_fuse_ok);
  const bool swap_epi_ok = want_swap && !p.flat && up_ok && (!p.stride2 || want_swap_s2) && (N % 128) == 0 && unit_ok &&
                           !(flags & (SD_EPI_OUT_F32 | SD_EPI_SOFTMAX | SD_EPI_SWISH)) && !residual &&
                           (out_ld % 2) == 0 && ((uintptr_t)out % 4) == 0;
  if (swap_epi_ok && all_1tap && N > 128) p.block_n = 128;
  p.n_tiles = (n_pad + p.block_n - 1) / p.block_n;
  // narrow N: pair two m-tiles per CTA tile so the B tile is fetched once per 256 rows (same smem traffic per
  // MMA cycle as the 128x256 tile, which runs near the tensor peak)
  p.dual = (!(flags & SD_EPI_SOFTMAX) && p.block_n <= 128 && phases * p.m_tiles * p.n_tiles >= 2 * num_sms() &&
            (!p.flat || !p.b_batched || (p.m_tiles_per_batch % 2) == 0)) ? 1 : 0;
  // thread-block clusters (optional): consecutive m-units share the B tile (weights) through TMA multicast.  Built to
  // cut the L2->SM traffic per MMA (the kernel moves ~44-54 B/clk/SM of TMA traffic at 57-73 % tensor activity, the
  // practical L2->SM ceiling), but measured on B200 it does not: cluster 2 is 2 % slower and cluster 4 45 % slower than
  // independent CTAs (multicast at cluster size <= 4 does not reduce L2 reads, and lock-stepped CTAs lose slack),
  // so the default is 1.  Parity-tested (tests/test_scorenet_ops_gpu.py with SDB_GEMM_CLUSTER=2).
  static const int want_cluster = [] { const char* e = getenv("SDB_GEMM_CLUSTER"); return e ? atoi(e) : 1; }();   // tuning knob, default off (see below)
  {
    const int m_units_h = p.dual ? (p.m_tiles + 1) / 2 : p.m_tiles;
    int c = want_cluster;
    while (c > 1 && (p.block_n % (8 * c) != 0 || m_units_h * p.n_tiles < 2 * num_sms() || p.b_batched)) c >>= 1;
    p.cluster = c < 1 ? 1 : c;
    // cta_group::2 pairs for wide tiles: halves the weight bytes each SM has to receive per K-block
    static const int want_pair = [] { const char* e = getenv("SDB_GEMM_PAIR"); return e ? atoi(e) : 1; }();   // tuning knob
    p.pair = 0;
    // pairs also for the 8x8 level (256 m-tiles at batch 512): same 1.73 waves as single 128 x 256 tiles, but each SM receives
    // 32 KB instead of 48 KB per K-block of a layer that is bound by the L2 -> SM feed
    static const int pair_min_tiles = [] { const char* e = getenv("SDB_GEMM_PAIR_MIN_TILES"); return e ? atoi(e) : 148; }();   // tuning knob
    if (want_pair && !p.dual && p.block_n == MAX_BN && !p.b_batched && !(flags & SD_EPI_SOFTMAX) &&
        phases * p.m_tiles * p.n_tiles >= pair_min_tiles) {
      p.pair = 1;
      p.cluster = 2;
    }
    // low-resolution layers whose N = 256 was halved above to spread < 148 tiles over the SMs (4x4 level at batch 512: 64 m-tiles
    // x 2): as pairs of M = 256 pixels x N = 128 channels each SM still owns one 128 x 128 accumulator but receives its pixel tile
    // and 64 weight rows (24 KB) instead of 128 (32 KB) per K-block -- these launches run at 2.5x their MMA time on the L2 -> SM feed
    static const int want_pair128 = [] { const char* e = getenv("SDB_GEMM_PAIR128"); return e ? atoi(e) : 1; }();   // tuning knob
    if (want_pair && want_pair128 && !p.pair && !p.dual && p.cluster == 1 && p.block_n == 128 && n_pad == MAX_BN && !p.flat &&
        !p.b_batched && !(flags & SD_EPI_SOFTMAX) && p.m_tiles >= 2) {
      p.pair = 1;
      p.cluster = 2;
    }
  }
  // N = 128 conv layers (all 32x32 layers of the score-net): as two 128 x 128 x 16 instructions per K step (dual) the tensor
  // pipe ran at 45-60 % -- every 64-clk instruction re-reads its 4 KB A tile AND the 4 KB weight tile from shared memory
  // (128 B/clk, the whole SM port, shared with the TMA writes).  Swapped, the weights are the M = 128 operand and the 256
  // pixels the N = 256 operand: one 128-clk instruction per K step, 12 KB of operand reads (96 B/clk) -- the shape the
  // N = 256 layers already run at 85-95 % of peak with.  The accumulator is then [channel][pixel]; see the epilogue.
  p.swap = (swap_epi_ok && p.dual && !p.pair && p.cluster == 1 && p.block_n == 128) ? 1 : 0;
  // pairs (N = 256 tiles): the same exchange of roles -- M = 256 channels over the two CTAs, N = 256 pixels -- gives the pair kernel
  // the [channel][pixel] epilogue too (no shuffle butterfly for the GroupNorm sums, 64-byte store runs)
  static const int want_pair_swap = [] { const char* e = getenv("SDB_GEMM_PAIR_SWAP"); return e ? atoi(e) : 1; }();   // tuning knob
  if (want_pair_swap && swap_epi_ok && p.pair && p.block_n == MAX_BN && (N % MAX_BN) == 0) p.swap = 1;
  p.mcast = p.pair ? 2 : p.cluster;          // B-tile slices per CTA (pair halves or multicast slices)
  // Fused GroupNorm + swish epilogue (p.gn_gamma set by the caller): only in the [channel][pixel] epilogue of swapped tiles, and only
  // when the CTAs that hold one image can exchange their channel sums -- the unit is the image (16x16: pair-swapped, or dual-swapped
  // with N = 128), or the image's 2 / 4 units form a thread-block cluster (32x32 dual-swapped: 4 units) that exists for the
  // distributed-shared-memory exchange only (no multicast: measured slower, see above).  Otherwise the launch runs unfused and
  // emits stats_out for sd_groupnorm_swish as before; *p_fused tells the caller which happened.
  {
    const int want_gn_fuse = gn_fuse_level();   // tuning knob (SDB_GN_FUSE / sd_set_gn_fuse): 0 off, 1 unit == image only, 2 + clusters
    // the fused epilogue is ~3x the plain one per tile: only worth it where the tile's MMAs hide it (K >= 1024; the first conv, one
    // 64-wide K block, went from 72 to 157 us fused against a 54 us GroupNorm pass)
    const bool long_k = p.num_kb >= 16;
    bool fuse = long_k && want_gn_fuse && p.gn_gamma != nullptr && !(flags & SD_GEMM_SPLIT3) && p.swap && (N % 32) == 0 && (BM % (N / 32)) == 0 && p.cluster == (p.pair ? 2 : 1);
    if (fuse && p.imgs_per_tile > 1) {
      fuse = mi_fuse_ok;                       // 8x8: four whole images per unit, statistics per 64-column chunk (no exchange)
    } else if (fuse) {
      if (p.pair) {
        fuse = p.tiles_per_img == 2;
      } else {
        // images spanning several units (32x32: four) need a thread-block cluster for the exchange.  Clusters of 4 CTAs with 204 KB
        // of shared memory each only fit where a GPC has 4 free SMs: 33 clusters (132 of 148 SMs) are co-resident on B200, and the
        // persistent grid is sized to that (see the launch below; with 37 clusters the last four ran as a second wave, 119 -> 275 us).
        // Kernel time is then break-even with conv + separate GroupNorm (178 vs 119 + 48 us at K = 1152, 344 vs 311 + 48 at K = 3456),
        // but 268 MB of DRAM traffic per GroupNorm disappear, and the power-capped timestep gets 2 % faster (profiles/r02_notes.md).
        const int units_per_img = p.tiles_per_img / 2;
        fuse = (units_per_img == 1 || (want_gn_fuse >= 2 && (units_per_img == 2 || units_per_img == 4))) && (num_sms() % units_per_img) == 0;
        if (fuse && units_per_img > 1) {
          // default: plain CTAs on every SM exchanging through global memory (cooperative launch); SDB_GN_GX=0 or no free region:
          // the cluster + distributed-shared-memory form on the 33 co-resident clusters of 4
          static const int want_gx = [] { const char* e = getenv("SDB_GN_GX"); return e ? atoi(e) : 1; }();   // tuning knob
          // __device__ variables have one instance per device: addresses looked up per device (cached)
          static float* gx_data_of[64] = {};
          static int* gx_cnt_of[64] = {};
          static std::mutex gx_mu;
          float* gx_data = nullptr;
          int* gx_cnt = nullptr;
          {
            int dev = 0;
            cudaGetDevice(&dev);
            std::lock_guard<std::mutex> lock(gx_mu);
            if (dev >= 0 && dev < 64) {
              if (!gx_data_of[dev]) {
                void *d = nullptr, *c = nullptr;
                if (cudaGetSymbolAddress(&d, g_gx_data) == cudaSuccess && cudaGetSymbolAddress(&c, g_gx_cnt) == cudaSuccess) {
                  gx_data_of[dev] = reinterpret_cast<float*>(d);
                  gx_cnt_of[dev] = reinterpret_cast<int*>(c);
                } else {
                  cudaGetLastError();
                }
              }
              gx_data = gx_data_of[dev];
              gx_cnt = gx_cnt_of[dev];
            }
          }
          const int region = (want_gx && gx_data && p.n_tiles == 1 && num_sms() / units_per_img <= GX_MAX_GROUPS) ? gx_region_of(st) : -1;
          if (region >= 0) {
            p.gx_units = units_per_img;
            p.gx_data = gx_data + (size_t)region * GX_MAX_GROUPS * 2 * GX_MAX_UNITS * 2 * BM;
            p.gx_cnt = gx_cnt + (size_t)region * GX_MAX_GROUPS * 2 * 2;
          } else {
            p.cluster = units_per_img; p.mcast = 1;
          }
        }
      }
    }
    // thread = pixel-row tiles that hold whole images (8x8: two per tile, 4x4: eight; N = 256 so that a group is 8 adjacent columns)
    if (!fuse && long_k && want_gn_fuse && p.gn_gamma != nullptr && !(flags & (SD_GEMM_SPLIT3 | SD_EPI_OUT_F32 | SD_EPI_SOFTMAX | SD_EPI_SWISH)) &&
        !residual && !p.swap && !p.dual && !p.flat && p.up_phase < 0 && !p.stride2 && N == MAX_BN &&
        (p.block_n == 128 || (p.pair && p.block_n == MAX_BN)) && (p.HW == 16 || p.HW == 64) && p.imgs_per_tile == BM / p.HW &&
        (out_ld % 8) == 0 && p.cluster == (p.pair ? 2 : 1))
      fuse = true;
    if (!fuse) {
      if (p.gn_raw_out) out = p.gn_raw_out;      // unfused: the raw result goes where the caller expects the raw tensor
      p.gn_gamma = nullptr;
      p.gn_raw_out = nullptr;
    }
  }
  {
    // ring slot = A region + the B rows this CTA receives; slab mode (see the kernel) packs three vertical taps per slot
    const int nsub_h = p.dual ? 2 : 1;
    const int b_cta = (p.pair ? p.block_n / 2 : p.block_n) * BK * 2;
    static const int want_slab = [] { const char* e = getenv("SDB_GEMM_SLAB"); return e ? atoi(e) : 1; }();   // tuning knob
    p.slab = 0;
    p.a_region_bytes = nsub_h * A_BYTES;
    int groups = 1;
    bool any_cand = false;
    for (int sgi = 0; sgi < p.nseg; ++sgi) any_cand = any_cand || p.seg_slab[sgi] >= 0;
    bool slab_on = false;
    if (want_slab && any_cand && (p.dual || p.pair) && (p.mcast == 1 || p.pair) && p.imgs_per_tile == 1 &&
        (nsub_h == 1 || (p.tiles_per_img % 2) == 0)) {
      const int hb = nsub_h * p.h_box;
      const int nt = p.slab_taps;                // vertical taps per slab: hb + nt - 1 image rows
      const int slab_bytes = (hb + nt - 1) * p.img_W * BK * 2;
      int a_region = slab_bytes > nsub_h * A_BYTES ? slab_bytes : nsub_h * A_BYTES;
      a_region = (a_region + 1023) / 1024 * 1024;
      if (2 * (a_region + nt * b_cta) <= RING_BYTES && hb + nt - 1 <= 256) {
        for (int sgi = 0; sgi < p.nseg; ++sgi) {
          if (p.seg_slab[sgi] < 0) continue;
          const cuuint64_t ld = (cuuint64_t)p.seg_ld[sgi];
          cuuint64_t dims[4] = {(cuuint64_t)p.seg_C[sgi], (cuuint64_t)p.img_W, (cuuint64_t)p.img_H, (cuuint64_t)p.slab_B};
          cuuint64_t strides[3] = {ld * 2, ld * 2 * p.img_W, ld * 2 * p.img_W * p.img_H};
          cuuint32_t box[4] = {BK, (cuuint32_t)p.img_W, (cuuint32_t)(hb + nt - 1), 1};
          int rc = encode_map(&p.a_slab_map[p.seg_slab[sgi]], p.seg_src[sgi], 4, dims, strides, box);
          if (rc != SD_OK) return rc;
        }
        slab_on = true;
        p.slab = 1;
        p.slab_bytes = slab_bytes;
        p.a_region_bytes = a_region;
        groups = nt;
      }
    }
    if (!slab_on)
      for (int sgi = 0; sgi < MAX_SEGS; ++sgi) p.seg_slab[sgi] = -1;
    int sb = p.a_region_bytes + groups * b_cta;
    sb = (sb + 1023) / 1024 * 1024;
    int ns = RING_BYTES / sb;
    if (ns > MAX_STAGES) ns = MAX_STAGES;
    static const int max_stages_env = [] { const char* e = getenv("SDB_GEMM_STAGES"); return e ? atoi(e) : MAX_STAGES; }();   // tuning knob
    if (ns > max_stages_env) ns = max_stages_env;
    if (ns < 2) ns = 2;
    p.stage_bytes = sb;
    p.num_stages = ns;
  }
  if (((uintptr_t)Wt % 16) != 0 || (ldb % 8) != 0) return fail(kErrInvalidArg, std::string(who) + ": B operand must be 16-byte aligned with ld % 8 == 0");
  {
    // rows beyond N inside the last box are zero-filled by TMA (OOB) and masked at the store
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)nbatchB};
    cuuint64_t strides[2] = {(cuuint64_t)ldb * 2, (cuuint64_t)(nbatchB > 1 ? strideB : (long long)N * ldb) * 2};
    cuuint32_t box[3] = {BK, (cuuint32_t)(p.block_n / p.mcast), 1};   // multicast slice or pair half
    int rc = encode_map(&p.b_map, Wt, 3, dims, strides, box);
    if (rc != SD_OK) return rc;
  }
  p.N_out = N;
  p.bias = bias;
  p.rowbias = rowbias;
  p.rowbias_ld = rowbias_ld;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.res_ld = out_ld;
  p.out = out;
  p.out_ld = out_ld;
  p.flags = flags;
  p.imgs_in_tile = (!p.flat && p.HW < BM) ? BM / p.HW : 1;
  p.stats_out = (p.gn_gamma && !p.gn_raw_out) ? nullptr : stats_out;   // a fused GroupNorm epilogue needs no channel sums downstream unless the raw tensor is kept too
  if (stats_out && ((p.flat ? (p.M_per_batch % BM) != 0 : (p.HW % BM) != 0) || (N % 16) != 0 || (flags & SD_EPI_SOFTMAX)))
    return fail(kErrInvalidArg, std::string(who) + ": stats_out needs whole 128-row tiles per image / batch entry and N a multiple of 16");
  const int out_align = (flags & SD_EPI_OUT_F32) ? 4 : 8;
  const bool vec_ok = (out_ld % out_align == 0) && ((uintptr_t)out % 16 == 0) && (!residual || ((uintptr_t)residual % 16 == 0 && out_ld % 8 == 0)) &&
                      (!bias || (uintptr_t)bias % 16 == 0) && (!rowbias || ((uintptr_t)rowbias % 16 == 0 && rowbias_ld % 4 == 0)) &&
                      (p.out_batch_stride % out_align == 0);
  if (!vec_ok && N >= 16) return fail(kErrInvalidArg, std::string(who) + ": out/residual/bias must be 16-byte aligned (ld multiple of 8 bf16 / 4 fp32)");

  static PerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM);
    return e;
  });
  if (attr_err != cudaSuccess) return check_cuda(attr_err, who);
  const int m_units_h = p.dual ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int groups = ((m_units_h + p.cluster - 1) / p.cluster) * p.n_tiles * phases;
  const int max_clusters = num_sms() / p.cluster;
  int grid = (groups < max_clusters ? groups : max_clusters) * p.cluster;
  if (p.gx_units > 1) grid = grid / p.gx_units * p.gx_units;     // whole images per tile round (the tile count is a multiple of gx_units)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GEMM_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[3];
  int nattr = 0;
  if (p.gx_units > 1) {        // the CTAs of an image wait for each other: the whole (persistent, <= 1 CTA per SM) grid must be co-resident
    attr[nattr].id = cudaLaunchAttributeCooperative;
    attr[nattr].val.cooperative = 1;
    ++nattr;
  }
  if (p.cluster > 1) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = p.cluster;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  if (pdl_enabled()) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  if (nattr) { cfg.attrs = attr; cfg.numAttrs = nattr; }
  if (p.cluster > 2 && !p.pair) {
    // a persistent grid must be co-resident: clusters of 4 CTAs (204 KB of shared memory each = one CTA per SM) only fit where a GPC
    // has 4 free SMs, so fewer than num_sms / 4 clusters may be active at once; ask the runtime and shrink the grid to that
    static int max_active4 = -1;
    if (max_active4 < 0) {
      int n = 0;
      cudaLaunchConfig_t q = cfg;
      q.gridDim = dim3((unsigned)(num_sms() / p.cluster * p.cluster));
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<false>, &q) != cudaSuccess || n < 1) { n = num_sms() / p.cluster; cudaGetLastError(); }
      max_active4 = n;
      if (getenv("SDB_GEMM_VERBOSE")) fprintf(stderr, "[gemm] max active clusters of %d: %d (of %d)\n", p.cluster, n, num_sms() / p.cluster);
    }
    const int want = (int)cfg.gridDim.x / p.cluster;
    if (want > max_active4) cfg.gridDim = dim3((unsigned)(max_active4 * p.cluster));
  }
  if (p.pair) return check_cuda(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<true>, p), who);
  return check_cuda(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<false>, p), who);
}

}  // namespace sdb

struct GnFuse { const float* gamma; const float* beta; float eps; int swish; int* fused; void* raw_out; };

static int conv_gemm_impl(const sd_gemm_src* srcs, int num_srcs, int B, int H, int W, const void* Wt, int N,
                          const float* bias, const float* rowbias, int rowbias_ld, const void* residual,
                          unsigned flags, void* out, int out_ld, float* stats_out, void* stream, int up_phase,
                          int stats_tpi_total, int stats_slot0, const char* who, int stride2 = 0, const GnFuse* gn = nullptr) {
  using namespace sdb;
  if (gn && gn->fused) *gn->fused = 0;
  if (!srcs || num_srcs < 1 || num_srcs > 3) return fail(kErrInvalidArg, std::string(who) + ": 1..3 sources required");
  if (!Wt || !out || B < 0 || H < 1 || W < 1 || N < 1) return fail(kErrInvalidArg, std::string(who) + ": bad argument");
  if (B == 0) return SD_OK;
  GemmParams p{};
  if (W > BM || (BM % W) != 0) return fail(kErrUnsupported, std::string(who) + ": W must divide 128");
  if (H * W < 16) return fail(kErrUnsupported, std::string(who) + ": images smaller than 16 pixels are not supported (row-bias table holds 8 images per tile)");
  const int h_box = (H * W >= BM) ? BM / W : H;
  if ((H % h_box) != 0 || (BM % (W * h_box)) != 0)
    return fail(kErrUnsupported, std::string(who) + ": H*W must divide or be a multiple of 128 in whole rows");
  p.h_box = h_box;
  p.tiles_per_img = (H * W >= BM) ? (H * W) / BM : 1;
  p.imgs_per_tile = (H * W >= BM) ? 1 : BM / (H * W);
  p.flat = 0;
  p.up_all = up_phase == 4 ? 1 : 0;          // 4: all phases of the fused upsample + conv in one launch
  p.up_phase = p.up_all ? 0 : up_phase;
  p.stride2 = stride2;
  p.img_H = H;
  p.img_W = W;
  p.stats_tpi_total = stats_tpi_total > 0 ? stats_tpi_total : p.tiles_per_img;
  p.stats_slot0 = stats_slot0;
  // Segments.  Plain: one per source, weight columns in source order.  SD_GEMM_SPLIT3: every source is a hi|lo pair
  // [.., 2C] and contributes three segments -- (hi, W_hi), (lo, W_hi), (hi, W_lo) -- against weights [N, 2*K_half] = [hi | lo].
  const bool split = (flags & SD_GEMM_SPLIT3) != 0;
  long K_half = 0;
  for (int s = 0; s < num_srcs; ++s) K_half += (long)srcs[s].taps * srcs[s].C;
  const long K = split ? 2 * K_half : K_half;
  for (int s = 0; s < MAX_SEGS; ++s) p.seg_slab[s] = -1;
  int nseg = 0, nslab = 0;
  long k_off = 0;
  for (int s = 0; s < num_srcs; ++s) {
    const sd_gemm_src& src = srcs[s];
    if (!src.ptr || src.C < BK || (src.C % BK) != 0) return fail(kErrInvalidArg, std::string(who) + ": source channels must be a multiple of 64");
    if (!(src.taps == 1 || src.taps == 9 || (src.taps == 4 && up_phase >= 0))) return fail(kErrInvalidArg, std::string(who) + ": taps must be 1 or 9");
    if (((uintptr_t)src.ptr % 16) != 0) return fail(kErrInvalidArg, std::string(who) + ": source must be 16-byte aligned");
    const int ld = src.ld > 0 ? src.ld : (split ? 2 * src.C : src.C);
    if (ld < (split ? 2 : 1) * src.C || (ld % 8) != 0) return fail(kErrInvalidArg, std::string(who) + ": source ld must cover the channels and be a multiple of 8");
    const int nterm = split ? 3 : 1;
    for (int term = 0; term < nterm; ++term) {
      if (nseg >= MAX_SEGS) return fail(kErrInvalidArg, std::string(who) + ": too many K segments");
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(src.ptr) + (term == 1 ? src.C : 0);     // lo half
      p.seg_taps[nseg] = src.taps;
      p.seg_cblocks[nseg] = src.C / BK;
      p.seg_bk0[nseg] = (int)(((term == 2 ? K_half : 0) + k_off) / BK);
      p.seg_src[nseg] = base; p.seg_C[nseg] = src.C; p.seg_ld[nseg] = ld;
      // slab candidates: 3x3 stride-1 segments; the 2x2-tap phases of the fused upsample + conv (two vertical taps per slab:
      // 9 instead of 16 image rows per (channel block, tap column) at 16x16 -- that launch is bound by the L2 -> SM feed)
      static const int want_slab_up = [] { const char* e = getenv("SDB_GEMM_SLAB_UP"); return e ? atoi(e) : 1; }();   // tuning knob
      if (((src.taps == 9 && !stride2 && up_phase < 0) || (src.taps == 4 && up_phase >= 0 && want_slab_up)) && nslab < MAX_SLABS)
        p.seg_slab[nseg] = nslab++;
      int rc;
      const cuuint64_t l2 = (cuuint64_t)ld * 2;
      if (stride2) {
        // the source is the (2H x 2W) input; to load N elements at element stride 2 the box extent is 2N
        cuuint64_t dims[4] = {(cuuint64_t)src.C, (cuuint64_t)(2 * W), (cuuint64_t)(2 * H), (cuuint64_t)B};
        cuuint64_t strides[3] = {l2, l2 * (2 * W), l2 * (2 * W) * (2 * H)};
        cuuint32_t box[4] = {BK, (cuuint32_t)(2 * W), (cuuint32_t)(2 * p.h_box), (cuuint32_t)p.imgs_per_tile};
        cuuint32_t estr[4] = {1, 2, 2, 1};
        rc = encode_map(&p.a_map[nseg], base, 4, dims, strides, box, estr);
      } else {
        cuuint64_t dims[4] = {(cuuint64_t)src.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {l2, l2 * W, l2 * W * H};
        cuuint32_t box[4] = {BK, (cuuint32_t)W, (cuuint32_t)p.h_box, (cuuint32_t)p.imgs_per_tile};
        rc = encode_map(&p.a_map[nseg], base, 4, dims, strides, box);
      }
      if (rc != SD_OK) return rc;
      ++nseg;
    }
    k_off += (long)src.taps * src.C;
  }
  p.nseg = nseg;
  p.num_kb = (int)(K / BK);
  p.slab_B = B;
  p.slab_taps = up_phase >= 0 ? 2 : 3;
  if (gn && gn->gamma && gn->beta) {
    p.gn_gamma = gn->gamma; p.gn_beta = gn->beta; p.gn_eps = gn->eps; p.gn_swish = gn->swish;
    p.gn_raw_out = reinterpret_cast<__nv_bfloat16*>(gn->raw_out);
  }
  p.M_total = B * H * W;
  p.HW = H * W;
  p.m_tiles = (p.M_total + BM - 1) / BM;
  p.m_tiles_per_batch = 1;
  const int rc = launch_gemm(p, N, K, Wt, (int)K, p.up_all ? (long long)N * K : 0, p.up_all ? 4 : 1, bias, rowbias, rowbias_ld,
                             residual, flags, out, out_ld, (cudaStream_t)stream, who, stats_out);
  if (rc == SD_OK && gn && gn->fused) *gn->fused = p.gn_gamma != nullptr ? 1 : 0;
  return rc;
}

extern "C" int sd_conv_gemm(const sd_gemm_src* srcs, int num_srcs, int B, int H, int W, const void* Wt, int N,
                            const float* bias, const float* rowbias, int rowbias_ld, const void* residual,
                            unsigned flags, void* out, int out_ld, float* stats_out, void* stream) {
  return conv_gemm_impl(srcs, num_srcs, B, H, W, Wt, N, bias, rowbias, rowbias_ld, residual, flags, out, out_ld,
                        stats_out, stream, -1, 0, 0, "sd_conv_gemm");
}

extern "C" int sd_conv_gemm_gn(const sd_gemm_src* srcs, int num_srcs, int B, int H, int W, const void* Wt, int N,
                               const float* bias, const float* rowbias, int rowbias_ld, unsigned flags, void* out, int out_ld,
                               float* stats_out, const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_swish,
                               void* raw_out, int* fused_host, void* stream) {
  if (!gn_gamma || !gn_beta || !fused_host) return sdb::fail(sdb::kErrInvalidArg, "sd_conv_gemm_gn: null GroupNorm parameters / fused flag");
  if (raw_out && (((uintptr_t)raw_out % 16) != 0 || raw_out == out)) return sdb::fail(sdb::kErrInvalidArg, "sd_conv_gemm_gn: raw_out must be a distinct 16-byte aligned buffer");
  GnFuse gn{gn_gamma, gn_beta, gn_eps, gn_swish, fused_host, raw_out};
  return conv_gemm_impl(srcs, num_srcs, B, H, W, Wt, N, bias, rowbias, rowbias_ld, nullptr, flags, out, out_ld, stats_out, stream, -1,
                        0, 0, "sd_conv_gemm_gn", 0, &gn);
}

extern "C" int sd_set_gn_fuse(int level) {
  if (level < 0 || level > 2) return sdb::fail(sdb::kErrInvalidArg, "sd_set_gn_fuse: level must be 0, 1 or 2");
  const int prev = sdb::gn_fuse_level();
  sdb::g_gn_fuse.store(level, std::memory_order_relaxed);
  return prev;
}

extern "C" int sd_conv_gemm_s2(const void* x, int B, int H_in, int W_in, int C, const void* Wt, int N, const float* bias,
                               unsigned flags, void* out, float* stats_out, void* stream) {
  using namespace sdb;
  if (!x || (H_in % 2) || (W_in % 2)) return fail(kErrInvalidArg, "sd_conv_gemm_s2: even input size required");
  sd_gemm_src src{x, C, 9};
  const int out_ld = (flags & SD_GEMM_SPLIT3) && !(flags & SD_EPI_OUT_F32) ? 2 * N : N;
  return conv_gemm_impl(&src, 1, B, H_in / 2, W_in / 2, Wt, N, bias, nullptr, 0, nullptr, flags, out, out_ld, stats_out, stream,
                        -1, 0, 0, "sd_conv_gemm_s2", 1);
}

extern "C" int sd_upconv_gemm(const void* x, int B, int H, int W, int C, const void* Wt4, int N, const float* bias,
                              unsigned flags, void* out, float* stats_out, void* stream) {
  using namespace sdb;
  if (!x || !Wt4 || !out) return fail(kErrInvalidArg, "sd_upconv_gemm: null pointer");
  if (flags & (SD_EPI_OUT_F32 | SD_EPI_SOFTMAX)) return fail(kErrInvalidArg, "sd_upconv_gemm: bf16 output only");
  if (stats_out && ((H * W) % BM) != 0) return fail(kErrInvalidArg, "sd_upconv_gemm: stats_out needs H*W % 128 == 0");
  const int tpi = (H * W >= BM) ? (H * W) / BM : 1;
  sd_gemm_src src{x, C, 4};
  const int kmul = (flags & SD_GEMM_SPLIT3) ? 2 : 1;          // split: per phase [N, 2 * 4C] = [hi | lo]; out rows [hi(N) | lo(N)]
  static const int one_launch = [] { const char* e = getenv("SDB_UPCONV_ONE_LAUNCH"); return e ? atoi(e) : 1; }();   // tuning knob
  if (one_launch)      // the four phases as one persistent launch: one ramp / tail instead of four (K = 4C per phase is short)
    return conv_gemm_impl(&src, 1, B, H, W, Wt4, N, bias, nullptr, 0, nullptr, flags, out, kmul * N, stats_out, stream, 4,
                          4 * tpi, 0, "sd_upconv_gemm");
  for (int ph = 0; ph < 4; ++ph) {
    const __nv_bfloat16* w = reinterpret_cast<const __nv_bfloat16*>(Wt4) + (size_t)ph * N * 4 * C * kmul;
    int rc = conv_gemm_impl(&src, 1, B, H, W, w, N, bias, nullptr, 0, nullptr, flags, out, kmul * N, stats_out, stream, ph,
                            4 * tpi, ph * tpi, "sd_upconv_gemm");
    if (rc != SD_OK) return rc;
  }
  return SD_OK;
}

static int batched_gemm_impl(const void* A, int lda, long long strideA, const void* Bt, int ldb, long long strideB,
                            int batch, int M, int N, int K, const float* bias, const void* residual,
                            unsigned flags, void* out, int ldc, long long strideC, void* stream,
                            float softmax_scale, int softmax_block, float* stats_out = nullptr) {
  using namespace sdb;
  if (!A || !Bt || !out || batch < 0 || M < 1 || N < 1 || K < BK || (K % BK) != 0)
    return fail(kErrInvalidArg, "sd_batched_gemm: bad argument (K must be a multiple of 64)");
  if (batch == 0) return SD_OK;
  if (((uintptr_t)A % 16) != 0 || (lda % 8) != 0 || (strideA % 8) != 0 || (strideB % 8) != 0)
    return fail(kErrInvalidArg, "sd_batched_gemm: A must be 16-byte aligned with lda, strides multiples of 8");
  GemmParams p{};
  p.softmax_scale = softmax_scale;
  p.softmax_block = softmax_block;
  if (flags & SD_EPI_SOFTMAX) {
    if (N > MAX_BN || (N % 32) != 0 || softmax_block < 1 || (N % softmax_block) != 0 || (BM % softmax_block != 0 && softmax_block % BM != 0) ||
        bias || residual || (flags & (SD_EPI_OUT_F32 | SD_EPI_SWISH)))
      return fail(kErrInvalidArg, "sd_attention_probs: N must be a multiple of 16, <= 256, and a multiple of the block");
  }
  p.flat = 1;
  p.a_batched = strideA != 0;
  p.b_batched = strideB != 0;
  p.M_per_batch = M;
  p.m_tiles_per_batch = (M + BM - 1) / BM;
  p.m_tiles = p.m_tiles_per_batch * batch;
  p.tiles_per_img = p.m_tiles_per_batch;          // stats_out slot = batch * tiles + tile-in-batch
  p.stats_tpi_total = p.m_tiles_per_batch;
  p.stats_slot0 = 0;
  p.imgs_per_tile = 1;
  p.up_phase = -1;
  p.out_batch_stride = strideC;
  const bool split = (flags & SD_GEMM_SPLIT3) != 0;      // A [.., 2K] and Bt [.., 2K] are hi|lo pairs: hi*hi + lo*hi + hi*lo
  if (split && (lda < 2 * K || ldb < 2 * K)) return fail(kErrInvalidArg, "sd_batched_gemm: split operands need lda, ldb >= 2K");
  p.nseg = split ? 3 : 1;
  for (int s = 0; s < MAX_SEGS; ++s) p.seg_slab[s] = -1;
  p.num_kb = (split ? 3 : 1) * (K / BK);
  p.M_total = M * batch;
  p.HW = 1;
  const int nbA = p.a_batched ? batch : 1;
  for (int s = 0; s < p.nseg; ++s) {
    p.seg_taps[s] = 1;
    p.seg_cblocks[s] = K / BK;
    p.seg_bk0[s] = s == 2 ? K / BK : 0;
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)M, 1, (cuuint64_t)nbA};
    const cuuint64_t bstride = (cuuint64_t)(p.a_batched ? strideA : (long long)M * lda) * 2;
    cuuint64_t strides[3] = {(cuuint64_t)lda * 2, bstride, bstride};
    cuuint32_t box[4] = {BK, BM, 1, 1};
    int rc = encode_map(&p.a_map[s], reinterpret_cast<const __nv_bfloat16*>(A) + (s == 1 ? K : 0), 4, dims, strides, box);
    if (rc != SD_OK) return rc;
  }
  const int KB = split ? 2 * K : K;
  return launch_gemm(p, N, KB, Bt, ldb, strideB, p.b_batched ? batch : 1, bias, nullptr, 0, residual, flags, out, ldc,
                     (cudaStream_t)stream, "sd_batched_gemm", stats_out);
}

extern "C" int sd_batched_gemm_stats(const void* A, int lda, long long strideA, const void* Bt, int ldb, long long strideB,
                                     int batch, int M, int N, int K, const float* bias, const void* residual,
                                     unsigned flags, void* out, int ldc, long long strideC, float* stats_out, void* stream) {
  if (flags & SD_EPI_SOFTMAX) return sdb::fail(sdb::kErrInvalidArg, "sd_batched_gemm_stats: no softmax epilogue");
  return batched_gemm_impl(A, lda, strideA, Bt, ldb, strideB, batch, M, N, K, bias, residual, flags, out, ldc, strideC,
                           stream, 1.f, 1, stats_out);
}

extern "C" int sd_batched_gemm(const void* A, int lda, long long strideA, const void* Bt, int ldb, long long strideB,
                               int batch, int M, int N, int K, const float* bias, const void* residual,
                               unsigned flags, void* out, int ldc, long long strideC, void* stream) {
  if (flags & SD_EPI_SOFTMAX) return sdb::fail(sdb::kErrInvalidArg, "sd_batched_gemm: use sd_attention_probs for the softmax epilogue");
  return batched_gemm_impl(A, lda, strideA, Bt, ldb, strideB, batch, M, N, K, bias, residual, flags, out, ldc, strideC,
                           stream, 1.f, 1);
}

extern "C" int sd_attention_probs(const void* Q, int ldq, long long strideQ, const void* Kt, int ldk, long long strideK,
                                  int batch, int S, int C, float scale, int block, void* P, void* stream) {
  return batched_gemm_impl(Q, ldq, strideQ, Kt, ldk, strideK, batch, S, S, C, nullptr, nullptr, SD_EPI_SOFTMAX, P, S,
                           (long long)S * S, stream, scale, block);
}
