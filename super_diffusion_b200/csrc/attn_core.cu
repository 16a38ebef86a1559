// Fused attention core on the 5th-gen tensor cores (sm_100a):   out = softmax_blocks(scale * Q K^T) V + bias + residual
//
// Replaces the  w = softmax(q k^T * C^-1/2);  h = w v;  x + NIN(h)  tail of AttnBlock (reference
// cifar/models/layers.py:505-511) after the bind-time folding of models/ddpm.py (Q = q' = h Wq Wk^T + Wk bq, K = h,
// V^T = (h Wv Wo)^T, bias = bv Wo + bo, residual = x).  The probability matrix never leaves the SM:
//
//   phase 1   S[128 x Sp]  = Q tile . K^T            tcgen05.mma, K loop over the C channels, accumulator in TMEM cols [0, Sp)
//   softmax   8 epilogue warps read S from TMEM (row max, then exp + row sum; the 1 / sum is applied to the output row), write exp as bf16 straight into shared memory in
//             the 128-byte-swizzled K-major layout a UMMA A operand needs (4 blocks of [128 rows x 64 keys])
//   phase 2   O[128 x C]   = P . (V^T)^T             A operand = P from shared memory, B = V^T tiles by TMA, TMEM cols [256, 256+C)
//   epilogue  + bias + residual -> bf16 out, optional per-tile channel sums for the following GroupNorm
//
// (This header describes attn_core_kernel, the round-1 form that still serves C != 256; the score-net's own shape, C = 256, takes
// attn_core_v2_kernel further down: the same phases software-pipelined over two tiles.)
// One CTA per SM, persistent over (batch entry, 128-row query tile); the MMA issuer starts the next tile's phase 1 as soon as
// this tile's phase 2 is issued (S is free by then), so it overlaps the epilogue; warp 0 = TMA producer (runs ahead through a 3-stage ring,
// so the V^T tiles of phase 2 and the next tile's Q / K arrive during the softmax), warp 1 = MMA issuer, warps 2..9 = softmax +
// epilogue.  Before: two launches (probabilities 70 us + P V 91 us at batch 512, S = 256) with P written to and re-read from HBM.
#include "common.cuh"
#include "tcgen05_util.cuh"
#include "../../include/superdiff_b200.h"
#include <cstdlib>

namespace sdb {

constexpr int AC_BM = 128, AC_BK = 64, AC_STAGES = 3;
constexpr int AC_STAGE_BYTES = 48 * 1024;           // Q tile 16 KB + K tile <= 32 KB   |   V^T tile <= 32 KB
constexpr int AC_P_BYTES = 64 * 1024;               // P: 128 rows x <= 256 keys bf16, 4 swizzled [128 x 64] blocks
constexpr int AC_EPI_WARPS = 8, AC_EPI_THREADS = 256, AC_THREADS = 64 + AC_EPI_THREADS;
constexpr int AC_AUX_FLOATS = 256 /* bias */ + 256 /* row exchange */ + 8 * 256 /* column partials */;
constexpr size_t AC_SMEM = (size_t)AC_STAGES * AC_STAGE_BYTES + AC_P_BYTES + AC_AUX_FLOATS * 4 + 256 + 1024;

struct AttnCoreParams {
  CUtensorMap q_map, k_map, v_map;      // Q: (C, Sp, nb) box (64, 128, 1); K: (C, Sp, nb) box (64, Sp, 1); V^T: (Sp, C, nb) box (64, C, 1)
  CUtensorMap r_map, i_map;             // v2: residual (C, Sp, nb) box (64, 128, 1); identity tiles (64, 256) box (64, 128)
  int nb, Sp, C, block, tiles;
  float scale;
  const float* bias;                    // [C] or null
  const __nv_bfloat16* residual;        // [nb][Sp][C] or null
  __nv_bfloat16* out;                   // [nb][Sp][C]
  float* stats_out;                     // [nb][Sp/128][2][C] or null
};

__device__ __forceinline__ void ac_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(AC_EPI_THREADS) : "memory"); }

// column sums over the 32 rows of a warp (see gemm_tcgen05.cu::warp_colsum16): lane L ends with column (L & 15), sum for L < 16,
// sum of squares for L >= 16
__device__ __forceinline__ float ac_colsum16(const float (&v)[16], int lane) {
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

__global__ void __launch_bounds__(AC_THREADS, 1) attn_core_kernel(const __grid_constant__ AttnCoreParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* pbuf = smem + (size_t)AC_STAGES * AC_STAGE_BYTES;
  float* aux = reinterpret_cast<float*>(pbuf + AC_P_BYTES);
  float* bias_sh = aux;                 // [256]
  float* xch = aux + 256;               // [128 rows][2 column halves]
  float* wstat = aux + 512;             // [8][256]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux + AC_AUX_FLOATS);
  uint64_t* empty_bar = full_bar + AC_STAGES;
  uint64_t* s_full = empty_bar + AC_STAGES;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = p_ready + 1;
  uint64_t* tile_done = o_full + 1;
  uint32_t* tmem_ptr_sh = reinterpret_cast<uint32_t*>(tile_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.Sp / AC_BM;
  const int kb1 = p.C / AC_BK, kb2 = p.Sp / AC_BK;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.q_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.k_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.v_map) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < AC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(s_full, 1); mbar_init(p_ready, AC_EPI_WARPS); mbar_init(o_full, 1); mbar_init(tile_done, AC_EPI_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) bias_sh[i] = (p.bias && i < p.C) ? p.bias[i] : 0.f;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sh;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int b = tile / m_tiles, mt = tile - b * m_tiles;
        for (int kb = 0; kb < kb1; ++kb) {          // phase 1 operands: Q tile + all keys, one 64-channel block
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * AC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)(AC_BM + p.Sp) * AC_BK * 2);
          tma_load_3d(&p.q_map, st, &full_bar[stage], kb * AC_BK, mt * AC_BM, b);
          tma_load_3d(&p.k_map, st + AC_BM * AC_BK * 2, &full_bar[stage], kb * AC_BK, 0, b);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        for (int kb = 0; kb < kb2; ++kb) {          // phase 2 operand: V^T, one 64-key block of all C channels
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * AC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)p.C * AC_BK * 2);
          tma_load_3d(&p.v_map, st, &full_bar[stage], kb * AC_BK, 0, b);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Sp >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
        // No wait for the previous tile's epilogue here: S was released when its softmax arrived on p_ready (which this thread
        // waited for before issuing that tile's phase 2), and O / P are only touched again after the NEXT p_ready, which the
        // epilogue warps reach after they have drained O.  So the next tile's S = Q K^T runs under the current epilogue.
        tcgen05_fence_after();
        for (int kb = 0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * AC_STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(sa), b_desc = umma_desc_sw128(sa + AC_BM * AC_BK * 2);
#pragma unroll
          for (int k = 0; k < AC_BK / 16; ++k)
            umma_bf16(tmem_S, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc1, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full);
        mbar_wait(p_ready, it & 1);                  // P is in shared memory (generic-proxy writes fenced by the writers)
        tcgen05_fence_after();
        for (int kb = 0; kb < kb2; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t a_desc = umma_desc_sw128(smem_u32(pbuf) + (uint32_t)kb * (AC_BM * AC_BK * 2));
          const uint64_t b_desc = umma_desc_sw128(smem_u32(smem + (size_t)stage * AC_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < AC_BK / 16; ++k)
            umma_bf16(tmem_O, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc2, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;             // alternate 32-column groups
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    const float sl2 = p.scale * 1.4426950408889634f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
      const int b = tile / m_tiles, mt = tile - b * m_tiles;
      const int rl = mt * AC_BM + row;                                   // row within the batch entry
      const int lo = (rl / p.block) * p.block, hi = lo + p.block;         // this row's softmax block of key columns
      mbar_wait(s_full, it & 1);
      tcgen05_fence_after();
      const uint32_t tS = tmem_S + ((uint32_t)(q * 32) << 16);
      float mx = -INFINITY;
      for (int c = chalf * 32; c < p.Sp; c += 64) {
        uint32_t r0[16], r1[16];
        tmem_ld16(tS + c, r0);
        tmem_ld16(tS + c + 16, r1);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c + j >= lo && c + j < hi) mx = fmaxf(mx, __uint_as_float(r0[j]));
          if (c + 16 + j >= lo && c + 16 + j < hi) mx = fmaxf(mx, __uint_as_float(r1[j]));
        }
      }
      xch[row * 2 + chalf] = mx;
      ac_epi_bar();
      mx = fmaxf(xch[row * 2], xch[row * 2 + 1]);
      ac_epi_bar();
      const float mxs = mx * sl2;
      // second (last) pass over S: e = exp(scale (s - max)) once per element; the UNNORMALISED e goes to shared memory as the
      // bf16 A operand of phase 2 and 1 / sum is applied to the output row in the epilogue (saves a third TMEM pass and a
      // second exp per element)
      float sum = 0.f;
      for (int c = chalf * 32; c < p.Sp; c += 64) {
        uint32_t r[2][16];
        tmem_ld16(tS + c, r[0]);
        tmem_ld16(tS + c + 16, r[1]);
        tmem_wait_ld();
        uint8_t* blk = pbuf + (size_t)(c >> 6) * (AC_BM * AC_BK * 2) + (size_t)row * 128;     // [128 x 64]-key block, this row
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c0 = c + h * 16 + 2 * j;
            const float e0 = (c0 >= lo && c0 < hi) ? exp2f(fmaf(__uint_as_float(r[h][2 * j]), sl2, -mxs)) : 0.f;
            const float e1 = (c0 + 1 >= lo && c0 + 1 < hi) ? exp2f(fmaf(__uint_as_float(r[h][2 * j + 1]), sl2, -mxs)) : 0.f;
            sum += e0 + e1;
            const __nv_bfloat162 hh = __floats2bfloat162_rn(e0, e1);
            w[j] = *reinterpret_cast<const uint32_t*>(&hh);
          }
          // two 16-byte chunks (8 keys each) of the row's 128-byte line, XOR-swizzled with (row & 7) like TMA's SWIZZLE_128B
          const int ch0 = ((c & 63) >> 3) + h * 2;
          *reinterpret_cast<uint4*>(blk + (((ch0) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(blk + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      xch[row * 2 + chalf] = sum;
      ac_epi_bar();
      const float inv = 1.f / (xch[row * 2] + xch[row * 2 + 1]);

      // ---- epilogue of phase 2: O / sum + bias + residual -> bf16, optional channel sums.  The residual row is fetched into
      // registers now, so its L2 latency hides behind the phase-2 MMAs instead of stalling every 16-column chunk.
      const size_t row_off = ((size_t)b * p.Sp + rl) * (size_t)p.C;
      uint4 resv[4][2][2];
      if (p.residual) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = chalf * 32 + ci * 64;
          if (c < p.C) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c);
            resv[ci][0][0] = rp[0]; resv[ci][0][1] = rp[1]; resv[ci][1][0] = rp[2]; resv[ci][1][1] = rp[3];
          }
        }
      }
      mbar_wait(o_full, it & 1);
      tcgen05_fence_after();
      const uint32_t tO = tmem_O + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = chalf * 32 + ci * 64;
        if (c >= p.C) break;
        uint32_t r[2][16];
        tmem_ld16(tO + c, r[0]);
        tmem_ld16(tO + c + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int n0 = c + h * 16;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_sh + n0 + j);
            v[j] = fmaf(__uint_as_float(r[h][j]), inv, bb.x); v[j + 1] = fmaf(__uint_as_float(r[h][j + 1]), inv, bb.y);
            v[j + 2] = fmaf(__uint_as_float(r[h][j + 2]), inv, bb.z); v[j + 3] = fmaf(__uint_as_float(r[h][j + 3]), inv, bb.w);
          }
          if (p.residual) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint4 u = resv[ci][h][hh];
              const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                v[hh * 8 + 2 * j] += __uint_as_float(ww[j] << 16);
                v[hh * 8 + 2 * j + 1] += __uint_as_float(ww[j] & 0xFFFF0000u);
              }
            }
          }
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            w[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + row_off + n0);
          op[0] = make_uint4(w[0], w[1], w[2], w[3]);
          op[1] = make_uint4(w[4], w[5], w[6], w[7]);
          if (p.stats_out) wstat[(q * 2 + (lane >> 4)) * 256 + n0 + (lane & 15)] = ac_colsum16(v, lane);
        }
      }
      if (p.stats_out) {
        ac_epi_bar();
        for (int i = et; i < 2 * p.C; i += AC_EPI_THREADS) {
          const int which = i >= p.C ? 1 : 0, n = i - which * p.C;
          const float t = wstat[(0 * 2 + which) * 256 + n] + wstat[(1 * 2 + which) * 256 + n] +
                          wstat[(2 * 2 + which) * 256 + n] + wstat[(3 * 2 + which) * 256 + n];
          p.stats_out[(((size_t)b * m_tiles + mt) * 2 + which) * p.C + n] = t;
        }
        ac_epi_bar();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tile_done);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}


// =====================================================================================================================
// v2 (C == 256, block % 16 == 0): the same four phases, software-pipelined over two tiles so that neither the tensor pipe nor
// the epilogue warps wait for each other, and with every pass over tensor memory made once:
//
//   * ONE pass over S: a thread pulls its 128 (64 at Sp = 128) scores into registers with back-to-back tcgen05.ld and releases the
//     S accumulator at once (s_free) -- the issuer starts the NEXT tile's Q K^T while this tile's max / exp run on registers
//     (v1 read S twice from tensor memory, 64 B/clk, and kept it until the probabilities were written);
//   * P V is issued with the operand roles exchanged: O^T[channel][query] = V^T[channel][key] . P[query][key]^T, two M = 128
//     channel halves, N = 128 queries (the bytes in shared memory are the same).  An epilogue thread now owns ONE channel:
//     bias is a register, the GroupNorm channel sums are thread-local (v1: a 31-shuffle reduce-scatter per 16 columns -- 35 %
//     of all stall samples in profiles/r01d_ncu_attn_core_phases.txt), a lane-pair exchange packs channel pairs so each store
//     instruction touches two 64-byte runs of the NHWC rows;
//   * the probabilities are normalised BEFORE they are rounded to bf16 (row sums exchanged through shared memory), so the
//     residual can ride the P V accumulation as four extra K-blocks  O^T[ch] += I[ch][k] x[query][k]  (identity tile x the
//     residual rows, both through the TMA ring; exact: 1.0 * bf16 into fp32) -- with 4-byte residual loads in the drain, ncu
//     showed 2/3 (HBM) resp. 1/4 (after an L2 prefetch) of all epilogue stall samples waiting for those words;
//   * the epilogue warps drain O of tile i-1 AFTER the softmax of tile i: P V of tile i-1 runs under softmax i, Q K^T of tile
//     i+1 under the drain, and the TMA ring keeps streaming through all of it.
//
// Issue order on the tensor pipe: S0, S1, PV0, S2, PV1, ...   (the producer loads in the same order)
// Barriers: s_full / s_free (S accumulator), p_ready (P in shared memory) / o_full (also "P may be overwritten"), o_free.

// Two [128 x 64] bf16 tiles with ones at (row == 64 j + column), j = 0, 1: the A operand that adds 64 channels of the residual
// tile to one 128-channel half of O^T inside the P V accumulation (constant-initialised: no allocation, no init launch).
struct IdentTiles {
  uint16_t v[2 * 128 * 64];
  constexpr IdentTiles() : v{} {
    for (int j = 0; j < 2; ++j)
      for (int c = 0; c < 64; ++c) v[(j * 128 + j * 64 + c) * 64 + c] = 0x3F80;      // bf16 1.0
  }
};
__device__ const IdentTiles g_attn_ident{};

constexpr int AC2_AUX_FLOATS = 2 * 256 /* row max [parity][row][half] */ + 2 * 256 /* row sum */;
constexpr size_t AC2_SMEM = (size_t)AC_STAGES * AC_STAGE_BYTES + AC_P_BYTES + AC2_AUX_FLOATS * 4 + 256 + 1024;

template <int SP>
__global__ void __launch_bounds__(AC_THREADS, 1) attn_core_v2_kernel(const __grid_constant__ AttnCoreParams p) {
  constexpr int C = 256;
  constexpr int NCH = SP / 64;                 // 32-column chunks of S per thread (two threads per row take alternate chunks)
  constexpr int KB1 = C / AC_BK, KB2 = SP / AC_BK, M_TILES = SP / AC_BM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* pbuf = smem + (size_t)AC_STAGES * AC_STAGE_BYTES;
  float* aux = reinterpret_cast<float*>(pbuf + AC_P_BYTES);
  float* xmax = aux;                    // [2][128][2]
  float* xsum = aux + 512;              // [2][128][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux + AC2_AUX_FLOATS);
  uint64_t* empty_bar = full_bar + AC_STAGES;
  uint64_t* s_full = empty_bar + AC_STAGES;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_ready = s_free + 1;
  uint64_t* o_full = p_ready + 1;
  uint64_t* o_free = o_full + 1;
  uint32_t* tmem_ptr_sh = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;    // grid <= tiles: >= 1
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.q_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.k_map) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.v_map) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < AC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(s_full, 1); mbar_init(s_free, AC_EPI_WARPS); mbar_init(p_ready, AC_EPI_WARPS);
      mbar_init(o_full, 1); mbar_init(o_free, AC_EPI_WARPS);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_sh)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_sh;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 256;     // O^T: channel half h in columns [256 + 128 h, +128)

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer: QK(0), [QK(k+1), V(k)] ... =====================
      int stage = 0;
      uint32_t phase = 0;
      auto load_qk = [&](int tile) {
        const int b = tile / M_TILES, mt = tile - b * M_TILES;
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * AC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)(AC_BM + SP) * AC_BK * 2);
          tma_load_3d(&p.q_map, st, &full_bar[stage], kb * AC_BK, mt * AC_BM, b);
          tma_load_3d(&p.k_map, st + AC_BM * AC_BK * 2, &full_bar[stage], kb * AC_BK, 0, b);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
      };
      load_qk((int)blockIdx.x);
      for (int k = 0; k < n_local; ++k) {
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        if (k + 1 < n_local) load_qk(tile + (int)gridDim.x);
        const int b = tile / M_TILES;
        for (int kb = 0; kb < KB2; ++kb) {          // V^T: one 64-key block of all 256 channels
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + (size_t)stage * AC_STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], (uint32_t)C * AC_BK * 2);
          tma_load_3d(&p.v_map, st, &full_bar[stage], kb * AC_BK, 0, b);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.residual) {
          const int mt = tile - b * M_TILES;
          for (int kb = 0; kb < KB1; ++kb) {        // residual: identity tile (kb & 1) + 64 channels of the tile's 128 rows
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + (size_t)stage * AC_STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], (uint32_t)2 * AC_BM * AC_BK * 2);
            tma_load_2d(&p.i_map, st, &full_bar[stage], 0, (kb & 1) * AC_BM);
            tma_load_3d(&p.r_map, st + AC_BM * AC_BK * 2, &full_bar[stage], kb * AC_BK, mt * AC_BM, b);
            if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer: S0, [S(k+1), PV(k)] ... =====================
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SP >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AC_BM >> 3) << 17) | ((uint32_t)(AC_BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      auto issue_s = [&]() {
        for (int kb = 0; kb < KB1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * AC_STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(sa), b_desc = umma_desc_sw128(sa + AC_BM * AC_BK * 2);
#pragma unroll
          for (int k = 0; k < AC_BK / 16; ++k)
            umma_bf16(tmem_S, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc1, (kb | k) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full);
      };
      issue_s();
      for (int k = 0; k < n_local; ++k) {
        if (k + 1 < n_local) {
          mbar_wait(s_free, k & 1);                  // every epilogue warp holds its part of S(k) in registers
          tcgen05_fence_after();
          issue_s();
        }
        mbar_wait(p_ready, k & 1);                   // P(k) is in shared memory (generic-proxy writes fenced by the writers)
        if (k > 0) mbar_wait(o_free, (k - 1) & 1);   // O(k-1) has been drained
        tcgen05_fence_after();
        for (int kb = 0; kb < KB2; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sv = smem_u32(smem + (size_t)stage * AC_STAGE_BYTES);
          const uint64_t p_desc = umma_desc_sw128(smem_u32(pbuf) + (uint32_t)kb * (AC_BM * AC_BK * 2));
#pragma unroll
          for (int h = 0; h < 2; ++h) {              // channel half h: A = its 128 rows of the V^T tile, B = the 128 query rows of P
            const uint64_t v_desc = umma_desc_sw128(sv + (uint32_t)h * (AC_BM * AC_BK * 2));
#pragma unroll
            for (int k4 = 0; k4 < AC_BK / 16; ++k4)
              umma_bf16(tmem_O + (uint32_t)h * 128u, v_desc + (uint64_t)(k4 * 2), p_desc + (uint64_t)(k4 * 2), idesc2, (kb | k4) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.residual) {
          for (int kb = 0; kb < KB1; ++kb) {         // + residual channels [64 kb, +64) into channel half kb / 2
            mbar_wait(&full_bar[stage], phase);
            tcgen05_fence_after();
            const uint32_t sr = smem_u32(smem + (size_t)stage * AC_STAGE_BYTES);
            const uint64_t i_desc = umma_desc_sw128(sr), x_desc = umma_desc_sw128(sr + AC_BM * AC_BK * 2);
#pragma unroll
            for (int k4 = 0; k4 < AC_BK / 16; ++k4)
              umma_bf16(tmem_O + (uint32_t)(kb >> 1) * 128u, i_desc + (uint64_t)(k4 * 2), x_desc + (uint64_t)(k4 * 2), idesc2, 1u);
            umma_commit(&empty_bar[stage]);
            if (++stage == AC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(o_full);
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;             // softmax: alternate 32-column chunks of the row; drain: channel half
    const int row = q * 32 + lane;
    const float sl2 = p.scale * 1.4426950408889634f;
    const bool odd = (lane & 1) != 0;
    const int ch = chalf * 128 + row, chp = ch & ~1;                  // drain: this thread's channel / its even-odd pair
    const float bv0 = p.bias ? p.bias[chp] : 0.f, bv1 = p.bias ? p.bias[chp + 1] : 0.f;
    const uint32_t tS = tmem_S + ((uint32_t)(q * 32) << 16);
    const uint32_t tO = tmem_O + (uint32_t)chalf * 128u + ((uint32_t)(q * 32) << 16);

    // O^T of tile `tile` (already normalised, residual included) -> out rows: a lane pair exchanges so that the even lane holds
    // channels (chp, chp + 1) of query c, the odd lane those of query c + 1; + bias, bf16x2 stores, thread-local channel sums.
    auto drain = [&](int tile) {
      const int b = tile / M_TILES, mt = tile - b * M_TILES;
      const size_t base = ((size_t)b * SP + (size_t)mt * AC_BM) * (size_t)C + (size_t)chp + (odd ? C : 0);
      float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int c = ci * 32;
        uint32_t r[2][16];
        tmem_ld16(tO + c, r[0]);
        tmem_ld16(tO + c + 16, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {             // queries c + 2j, c + 2j + 1
          const float a0 = __uint_as_float(r[j >> 3][(2 * j) & 15]), a1 = __uint_as_float(r[j >> 3][(2 * j + 1) & 15]);
          const float recv = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
          const float e0 = (odd ? recv : a0) + bv0, e1 = (odd ? a1 : recv) + bv1;
          s0 += e0; q0 = fmaf(e0, e0, q0);
          s1 += e1; q1 = fmaf(e1, e1, q1);
          *reinterpret_cast<__nv_bfloat162*>(p.out + base + (size_t)(c + 2 * j) * C) = __floats2bfloat162_rn(e0, e1);
        }
      }
      if (p.stats_out) {                           // channel totals over the tile's 128 queries: mine + the lane partner's
        const float t0 = s0 + __shfl_xor_sync(0xffffffffu, s0, 1), t1 = s1 + __shfl_xor_sync(0xffffffffu, s1, 1);
        const float u0 = q0 + __shfl_xor_sync(0xffffffffu, q0, 1), u1 = q1 + __shfl_xor_sync(0xffffffffu, q1, 1);
        float* so = p.stats_out + ((size_t)b * M_TILES + mt) * 2 * C;
        so[ch] = odd ? t1 : t0;
        so[C + ch] = odd ? u1 : u0;
      }
    };

    // iteration k: softmax of tile k (k < n_local), then the drain of tile k - 1 (k > 0) -- one copy of each in the code
    int prev = -1;
    for (int k = 0; k <= n_local; ++k) {
      const bool has = k < n_local;
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int par = k & 1;
      if (has) {
        const int mt = tile % M_TILES;
        const int rl = mt * AC_BM + row;                                   // row within the batch entry
        const int lo = (rl / p.block) * p.block, hi = lo + p.block;         // this row's softmax block of key columns
        mbar_wait(s_full, par);
        tcgen05_fence_after();
        uint32_t s[NCH * 2][16];
#pragma unroll
        for (int it = 0; it < NCH; ++it) {
          tmem_ld16(tS + chalf * 32 + it * 64, s[2 * it]);
          tmem_ld16(tS + chalf * 32 + it * 64 + 16, s[2 * it + 1]);
        }
        tmem_wait_ld();
        tcgen05_fence_before();                      // S is in registers: the next tile's Q K^T may overwrite it
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float mx = -INFINITY;
#pragma unroll
        for (int g = 0; g < NCH * 2; ++g) {          // block % 16 == 0: a 16-column group is inside or outside the row's block as a whole
          const int c16 = chalf * 32 + (g >> 1) * 64 + (g & 1) * 16;
          const bool in = c16 >= lo && c16 < hi;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float v = in ? __uint_as_float(s[g][j]) : -INFINITY;
            s[g][j] = __float_as_uint(v);
            mx = fmaxf(mx, v);
          }
        }
        xmax[par * 256 + row * 2 + chalf] = mx;
        ac_epi_bar();
        mx = fmaxf(xmax[par * 256 + row * 2], xmax[par * 256 + row * 2 + 1]);     // finite: the row's own column is inside its block
        const float mxs = mx * sl2;
        // e = exp(scale (s - max)) once per element (kept in the S registers); the row sum is exchanged with the partner thread
        // and the probabilities are normalised before they are rounded to bf16
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < NCH * 2; ++g)
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e = exp2f(fmaf(__uint_as_float(s[g][j]), sl2, -mxs));
            sum += e;
            s[g][j] = __float_as_uint(e);
          }
        xsum[par * 256 + row * 2 + chalf] = sum;
        ac_epi_bar();
        const float inv = 1.f / (xsum[par * 256 + row * 2] + xsum[par * 256 + row * 2 + 1]);
        uint32_t pk[NCH * 16];
#pragma unroll
        for (int g = 0; g < NCH * 2; ++g)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(__uint_as_float(s[g][2 * j]) * inv, __uint_as_float(s[g][2 * j + 1]) * inv);
            pk[g * 8 + j] = *reinterpret_cast<const uint32_t*>(&hh);
          }
        if (k > 0) {
          mbar_wait(o_full, par ^ 1);                // P V of the previous tile has retired: P may be overwritten, O(k-1) is complete
          tcgen05_fence_after();
        }
#pragma unroll
        for (int g = 0; g < NCH * 2; ++g) {
          const int c = chalf * 32 + (g >> 1) * 64;                        // the chunk's first key
          uint8_t* blk = pbuf + (size_t)(c >> 6) * (AC_BM * AC_BK * 2) + (size_t)row * 128;     // [128 x 64]-key block, this row
          // two 16-byte pieces (8 keys each) of the row's 128-byte line, XOR-swizzled with (row & 7) like TMA's SWIZZLE_128B
          const int ch0 = ((c & 63) >> 3) + (g & 1) * 2;
          *reinterpret_cast<uint4*>(blk + (((ch0) ^ (row & 7)) << 4)) = make_uint4(pk[g * 8], pk[g * 8 + 1], pk[g * 8 + 2], pk[g * 8 + 3]);
          *reinterpret_cast<uint4*>(blk + (((ch0 + 1) ^ (row & 7)) << 4)) = make_uint4(pk[g * 8 + 4], pk[g * 8 + 5], pk[g * 8 + 6], pk[g * 8 + 7]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
      } else {
        mbar_wait(o_full, par ^ 1);                  // the last tile's P V
        tcgen05_fence_after();
      }
      if (k > 0) {
        drain(prev);
        if (has) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_free);
        }
      }
      prev = tile;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace sdb

extern "C" int sd_attention_core(const void* Q, int ldq, long long strideQ, const void* K, int ldk, long long strideK,
                                 const void* Vt, int ldv, long long strideV, int batch, int S, int C, float scale, int block,
                                 const float* bias, const void* residual, void* out, float* stats_out, void* stream) {
  using namespace sdb;
  if (!Q || !K || !Vt || !out || batch < 0) return fail(kErrInvalidArg, "sd_attention_core: null pointer");
  if (!(S == 128 || S == 256) || C < 64 || C > 256 || (C % 64) != 0)
    return fail(kErrUnsupported, "sd_attention_core: S must be 128 or 256 and C a multiple of 64 up to 256");
  if (block < 1 || (S % block) != 0 || (AC_BM % block != 0 && block % AC_BM != 0))
    return fail(kErrInvalidArg, "sd_attention_core: block must divide S and divide or be a multiple of 128");
  if (stats_out && block != S) return fail(kErrInvalidArg, "sd_attention_core: stats_out needs one image per batch entry (block == S)");
  if ((((uintptr_t)Q | (uintptr_t)K | (uintptr_t)Vt | (uintptr_t)out | (uintptr_t)(residual ? residual : out)) % 16) != 0 ||
      (ldq % 8) || (ldk % 8) || (ldv % 8) || (strideQ % 8) || (strideK % 8) || (strideV % 8))
    return fail(kErrInvalidArg, "sd_attention_core: operands must be 16-byte aligned with leading dimensions / strides multiples of 8");
  if (batch == 0) return SD_OK;
  AttnCoreParams p{};
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)batch};
    cuuint64_t sq[2] = {(cuuint64_t)ldq * 2, (cuuint64_t)strideQ * 2}, sk[2] = {(cuuint64_t)ldk * 2, (cuuint64_t)strideK * 2};
    cuuint32_t boxq[3] = {AC_BK, AC_BM, 1}, boxk[3] = {AC_BK, (cuuint32_t)S, 1};
    int rc = encode_map(&p.q_map, Q, 3, dims, sq, boxq);
    if (rc == SD_OK) rc = encode_map(&p.k_map, K, 3, dims, sk, boxk);
    cuuint64_t dimv[3] = {(cuuint64_t)S, (cuuint64_t)C, (cuuint64_t)batch};
    cuuint64_t sv[2] = {(cuuint64_t)ldv * 2, (cuuint64_t)strideV * 2};
    cuuint32_t boxv[3] = {AC_BK, (cuuint32_t)C, 1};
    if (rc == SD_OK) rc = encode_map(&p.v_map, Vt, 3, dimv, sv, boxv);
    if (rc != SD_OK) return rc;
  }
  p.nb = batch; p.Sp = S; p.C = C; p.block = block; p.tiles = batch * (S / AC_BM);
  p.scale = scale; p.bias = bias; p.residual = (const __nv_bfloat16*)residual; p.out = (__nv_bfloat16*)out; p.stats_out = stats_out;
  static PerDeviceOnce attr_once;
  const cudaError_t attr_err = attr_once.run([] {
    cudaError_t e = cudaFuncSetAttribute(attn_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AC_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_v2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AC2_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_v2_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AC2_SMEM);
    return e;
  });
  static const bool force_v1 = [] { const char* e = getenv("SDB_ATTN_V1"); return e != nullptr && e[0] == '1'; }();   // A/B switch: the round-1 kernel
  if (attr_err != cudaSuccess) return check_cuda(attr_err, "sd_attention_core");
  const int grid = p.tiles < num_sms() ? p.tiles : num_sms();
  if (C == 256 && (block % 16) == 0 && scale > 0.f && !force_v1) {
    if (residual) {
      // the residual rides the P V accumulation: its rows as a TMA operand + the constant identity tiles
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)S, (cuuint64_t)batch};
      cuuint64_t sr[2] = {(cuuint64_t)C * 2, (cuuint64_t)S * C * 2};
      cuuint32_t boxr[3] = {AC_BK, AC_BM, 1};
      int rc = encode_map(&p.r_map, residual, 3, dims, sr, boxr);
      // a __device__ variable has one instance per device: look the address up per device (cached)
      static void* ident_of[64] = {};
      static std::mutex imu;
      int dev = 0;
      cudaGetDevice(&dev);
      void* ident = nullptr;
      {
        std::lock_guard<std::mutex> lock(imu);
        if (dev < 0 || dev >= 64) return fail(kErrUnsupported, "sd_attention_core: device ordinal out of range");
        if (!ident_of[dev]) {
          const cudaError_t ierr = cudaGetSymbolAddress(&ident_of[dev], g_attn_ident);
          if (ierr != cudaSuccess) { ident_of[dev] = nullptr; return check_cuda(ierr, "sd_attention_core (identity tiles)"); }
        }
        ident = ident_of[dev];
      }
      cuuint64_t dimi[2] = {AC_BK, 2 * AC_BM};
      cuuint64_t si[1] = {AC_BK * 2};
      cuuint32_t boxi[2] = {AC_BK, AC_BM};
      if (rc == SD_OK) rc = encode_map(&p.i_map, ident, 2, dimi, si, boxi);
      if (rc != SD_OK) return rc;
    }
    // the score-net's own shape (cifar/models/layers.py:505-511 at 256 channels): two-tile software pipeline, see attn_core_v2_kernel
    if (S == 256) attn_core_v2_kernel<256><<<grid, AC_THREADS, AC2_SMEM, (cudaStream_t)stream>>>(p);
    else attn_core_v2_kernel<128><<<grid, AC_THREADS, AC2_SMEM, (cudaStream_t)stream>>>(p);
    return check_cuda(cudaGetLastError(), "sd_attention_core launch");
  }
  attn_core_kernel<<<grid, AC_THREADS, AC_SMEM, (cudaStream_t)stream>>>(p);
  return check_cuda(cudaGetLastError(), "sd_attention_core launch");
}
