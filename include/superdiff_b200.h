/* superdiff_b200 — C ABI of the B200-native SuperDiff sampling path.
 *
 * The reference (mo-rsa24/super-diffusion) is pure Python and has no FFI
 * layer; the seams this library plugs into are its Python closures
 * (SURVEY.md §8b).  Each entry point names the reference code whose per-step
 * body it replaces.  Conventions:
 *   - every pointer is a BORROWED DEVICE pointer unless the name ends in
 *     `_host`; the caller (PyTorch) owns all memory;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never synchronise the device, and may be captured into a CUDA graph;
 *   - return value: 0 on success, negative error code otherwise
 *     (SD_ERR_*); sd_last_error() returns a thread-local message;
 *   - stateless and re-entrant.  There is NO CPU fallback: without an
 *     sm_100 device the launch fails and the error code says so.
 */
#ifndef SUPERDIFF_B200_H_
#define SUPERDIFF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_INVALID_ARG (-1)
#define SD_ERR_UNSUPPORTED (-2)
#define SD_ERR_CUDA (-3)

#define SD_MAX_MODELS 8

/* how the mixing weights kappa_j (sum_j kappa_j = 1) are chosen */
enum sd_mode {
  SD_MODE_OR = 0,    /* softmax_j(T*(logq_j + logp_j)): cifar/dynamics.py:124 (T=1e6),
                        notebooks/superposition_edu.ipynb:813 (T=1), clip_eval.py:402 */
  SD_MODE_AND = 1,   /* equalise the density increments: superposition_edu.ipynb:899-905
                        (M=2 closed form), general M by the linear solve of SURVEY.md A.3 */
  SD_MODE_AVG = 2,   /* kappa_j = 1/M: cifar/dynamics.py:155-171 */
  SD_MODE_FIXED = 3  /* kappa read from `weights` (caller supplied) */
};

/* which log-density increment is accumulated into logq */
enum sd_dlogq_mode {
  SD_DLOGQ_CIFAR_MAXSUB = 0, /* cifar/dynamics.py:131-135: increment minus its max over models */
  SD_DLOGQ_ITO = 1,          /* superposition_edu.ipynb:777-780: R_i + ito_scale*dt*a */
  SD_DLOGQ_NONE = 2          /* logq left untouched (cifar/dynamics.py:171 returns zeros) */
};

/* Fused VP-SDE SuperDiff step (one launch per timestep).
 * Replaces the body of get_joint_stoch_vf.joint_vf (cifar/dynamics.py:123-136),
 * get_avg_vf.joint_vf (:155-171) and the toy notebook's per-step cells
 * (superposition_edu.ipynb:813-819, :899-905, :938-946), after the M score
 * evaluations:
 *     kappa   = mode(logq | Gram reductions)
 *     dx      = -dt*(a*x - 2*b*sum_j kappa_j s_j) + sqrt(2*sigma*b*dt)*noise
 *     x_out   = x + dx                      (x_out may alias x)
 *     logq_i += dlogq_mode(R_i),  R_i = [<dx,s_i> + dt*a*<x,s_i> - dt*b*|s_i|^2]/sigma
 * x, noise, scores[j], x_out: [B, D] fp32 contiguous; logq, weights: [B, M].
 * `sched` (optional, device): table of per-step (a, b, sigma, dt) quadruples;
 * when non-NULL the scalars are read from sched[4 * (*step_counter)] instead of
 * the by-value arguments so that a captured CUDA graph can be replayed for
 * every timestep (step_counter NULL = row 0).
 * `logp_bias` (optional): [M] additive bias inside the OR softmax.
 * `ito_scale`: constant multiplier of dt*a in SD_DLOGQ_ITO mode; the notebook's
 * value is ndim*D (= 4 for the 2-D toy, superposition_edu.ipynb:778). */
int sd_step_vpsde(const float* x, const float* noise, const float* const* scores_host /* M device ptrs, host array */,
                  int M, int B, int D,
                  float a_t, float b_t, float sigma_t, float dt,
                  const float* sched, const int* step_counter,
                  int mode, int dlogq_mode, float temperature, const float* logp_bias, float ito_scale,
                  float* logq, float* x_out, float* weights, void* stream);

/* Same as sd_step_vpsde with explicit launch shape (tuning / benchmarks):
 * threads per CTA (64..256, multiple of 32), float4 chunks per thread (1..4),
 * CTAs per sample (thread-block cluster size 1,2,4,8).  0 = heuristic.
 * SD_MODE_AND only: cluster = -1 selects the two-pass streaming kernel (one CTA per
 * sample, 1..2 chunks per thread per round, no limit on D) that the heuristic uses
 * when a sample fits neither registers nor shared memory; cluster = -2 selects the
 * shared-memory-resident kernel (bulk copies of noise + M scores into (M+1)*D*4 bytes
 * of shared memory; needs 16-byte aligned rows and D <= 16*threads), the heuristic's
 * choice for M >= 3 at D = 3072. */
int sd_step_vpsde_ex(const float* x, const float* noise, const float* const* scores_host,
                     int M, int B, int D,
                     float a_t, float b_t, float sigma_t, float dt,
                     const float* sched, const int* step_counter,
                     int mode, int dlogq_mode, float temperature, const float* logp_bias, float ito_scale,
                     float* logq, float* x_out, float* weights, void* stream,
                     int threads, int vec_per_thread, int cluster);

/* Deterministic (probability-flow ODE) SuperDiff step with a caller-supplied divergence term -- the body of
 * get_joint_vf.joint_vf after the jax.jvp calls (reference cifar/dynamics.py:87-96), get_avg_vf(stoch=False) (:165-167)
 * and get_vpsde's vector_field (:48-54):
 *     w = softmax(T*logq) | 1/M | fixed;   dx = -dt*(a*x - b*sum_i w_i s_i);   x_out = x + dx
 *     dlogq_i = dlogq_add[b][i] + sum_d s_i/sigma_eps * (dx + dt*(a*x - b*s_i));  CIFAR_MAXSUB subtracts the row max (:95)
 * with sigma_eps = t + 1e-3 (:85) and dlogq_add = dt*div_i = -dt*b*<jvp_i, eps_i> (:86,:93; see sd_rowdot).
 * No noise tensor is read: 4*B*D*(M+2) bytes.  sched rows are (a, b, sigma_eps, dt).  SD_MODE_AND is rejected. */
int sd_step_vpsde_ode(const float* x, const float* const* scores_host, int M, int B, int D,
                      float a_t, float b_t, float sigma_eps, float dt, const float* sched, const int* step_counter,
                      int mode, int dlogq_mode, float temperature, const float* logp_bias,
                      const float* dlogq_add /* [B][M] or NULL */, float* logq, float* x_out, float* weights, void* stream);

/* out[b*out_stride] = scale * <a[b,:], b[b,:]> over D fp32 elements (fp64 reduction): the Hutchinson contraction
 * (jvp_val*eps).sum((1,2,3)) of cifar/dynamics.py:86 written straight into column i of the [B][M] dlogq_add array. */
int sd_rowdot(const float* a, const float* b, int B, int D, float scale, float* out, int out_stride, void* stream);

/* Fused EDM-sigma SuperDiff step on Stable-Diffusion latents with
 * classifier-free guidance.  Replaces applications/images/clip_eval.py:395-413
 * (methods "and", "or") and :417-424 ("avg"):
 *     noise = sqrt(2|dsigma|sigma) * z
 *     kappa = AND closed form (:398-400) | softmax([T(ll_obj+logp), T ll_bg])[0] (:402) | kappa_fixed
 *     vf    = v_unc + g*((v_bg - v_unc) + kappa*(v_obj - v_bg));  latents_out = latents + 2*dsigma*vf + noise
 *     ll_k += and/avg: sum(-|dsigma|/sigma v_k^2 - dx v_k/sigma) (:409-410) | or: -sum(v_k(dx + dsigma v_k))/sigma (:412-413)
 * latents, z, v_obj, v_bg, v_unc, latents_out: [B, D] fp32; ll: [B, 2] in/out
 * (obj, bg); kappa_out: [B].  lift_term = sigma*lift/num_inference_steps (:399). */
int sd_step_edm_cfg(const float* latents, const float* z, const float* v_obj, const float* v_bg, const float* v_unc,
                    int B, int D, float sigma, float dsigma, float guidance, float lift_term,
                    int mode /* SD_MODE_AND | SD_MODE_OR | SD_MODE_AVG (kappa_fixed) */,
                    float temperature, float logp, float kappa_fixed,
                    float* ll, float* latents_out, float* kappa_out, void* stream);

/* Deterministic AND on SD latents (method "and_ode", applications/images/clip_eval.py:377-391):
 *     kappa = [sigma (dlog_o - dlog_b) + <v_o - v_b, v_o + v_b> + lift_term - <v_o - v_b, v_u + g (v_b - v_u)>] / (g |v_o - v_b|^2)
 *     vf = v_u + g ((v_b - v_u) + kappa (v_o - v_b));   latents_out = latents + dsigma * vf
 *     ll_k += dsigma * (dlog_k + <v_k, v_k - vf> / sigma)
 * dlog: [B][2] = (dlog_obj, dlog_bg) = -<eps, J eps> Hutchinson estimates of get_vel(..., get_div=True) (:97-103; see
 * sd_rowdot); lift_term = lift/dsigma*sigma/num_inference_steps (:383).  4*B*D*5 bytes. */
int sd_step_edm_ode(const float* latents, const float* v_obj, const float* v_bg, const float* v_unc, const float* dlog,
                    int B, int D, float sigma, float dsigma, float guidance, float lift_term,
                    float* ll /* [B][2] in/out */, float* latents_out, float* kappa_out /* [B] */, void* stream);

/* In-graph helper: *counter += delta (single thread).  Lets a captured graph
 * advance the row of `sched` it reads. */
int sd_counter_add(int* counter, int delta, void* stream);
/* Same, saturating: *counter = clamp(*counter + delta, 0, rows - 1).  The kernels that read `sched[*step_counter]`
 * (sd_step_vpsde*, sd_time_embedding, sd_scorenet_forward_sched) do not know the table length, so the row index must stay
 * inside the table: a sampler that advances its counter ONLY through this entry can be replayed past the last timestep
 * without reading out of bounds (it then repeats the last row). */
int sd_counter_add_sat(int* counter, int delta, int rows, void* stream);

/* ------------------------------------------------------------------------
 * Score-network ops (cifar/models/ddpm.py:47-101 and the layers it uses).
 * Activations are NHWC; GEMM operands are bf16, accumulation fp32 in TMEM.
 * ------------------------------------------------------------------------ */

/* One A-operand segment of an implicit GEMM: an NHWC bf16 tensor [B,H,W,C]
 * read through `taps` filter taps (9 = 3x3 SAME stride 1, 1 = 1x1 / NIN). */
typedef struct sd_gemm_src {
  const void* ptr; /* bf16 [B, H, W, C] ([B, H, W, 2C] hi|lo under SD_GEMM_SPLIT3) */
  int C;           /* channels (multiple of 64) */
  int taps;        /* 1 or 9 */
  int ld;          /* elements between consecutive pixels; 0 = dense (C, or 2C under SD_GEMM_SPLIT3) */
} sd_gemm_src;

#define SD_EPI_SWISH 1u     /* out = swish(out) after everything else */
#define SD_EPI_OUT_F32 2u   /* `out` is fp32 instead of bf16 */
#define SD_EPI_SOFTMAX 4u   /* internal: row softmax epilogue (sd_attention_probs) */
/* 3 x bf16 split precision (the FP32-faithful arm, SD_PRECISION_FP32_FAITHFUL): every ACTIVATION operand of the call --
 * the sources / A / Bt-as-activation, `residual`, and `out` unless SD_EPI_OUT_F32 -- is a hi|lo pair of bf16 tensors stored
 * side by side along the channel (K / N) axis: [.., 2C] = [hi(C) | lo(C)], value = hi + lo, hi = bf16(v), lo = bf16(v - hi)
 * (~16 mantissa bits).  Weights are [N, 2K] = [hi(K) | lo(K)] of the fp32 weights.  The product is evaluated as
 * hi*hi + lo*hi + hi*lo on the tensor cores with fp32 accumulation (the dropped lo*lo term is ~2^-18 relative), i.e. three
 * K segments per source that share the weight columns.  C / K / N arguments keep their LOGICAL meaning (channels per half);
 * leading dimensions are the physical ones (>= 2C). */
#define SD_GEMM_SPLIT3 8u

/* out[b,h,w,n] = sum_seg sum_tap sum_c src[b,h+dh,w+dw,c] * Wt[n, k(seg,tap,c)]
 *               + bias[n] + rowbias[b, n] + residual[b,h,w,n]
 * Replaces nn.Conv 3x3 / NIN / nn.Dense call sites: cifar/models/layers.py:95-107,
 * :464-475, :556, and the residual add of :565 / :511.
 * Wt: bf16 [N, K] row-major (K = sum_seg taps*C, segment-major, tap-major, then
 * channel).  bias: fp32 [N] or NULL.  rowbias: fp32 [B, rowbias_ld] or NULL
 * (the per-sample time-embedding projection, layers.py:556).  residual: bf16
 * [B,H,W,N] or NULL.  N multiple of 16, <= 256 per tile (larger N is tiled). */
int sd_conv_gemm(const sd_gemm_src* srcs_host, int num_srcs, int B, int H, int W,
                 const void* Wt, int N, const float* bias,
                 const float* rowbias, int rowbias_ld,
                 const void* residual, unsigned flags, void* out, int out_ld,
                 float* stats_out /* optional fp32 [B*H*W/128][2][N]: per 128-pixel tile, per output channel sum and
                                     sum of squares of `out` (feeds sd_groupnorm_swish, saving its statistics pass);
                                     needs H*W % 128 == 0 and N % 16 == 0 */,
                 void* stream);

/* sd_conv_gemm with the GroupNorm (+ swish) that consumes its output fused into the epilogue -- conv1 of a ResnetBlockDDPM
 * (cifar/models/layers.py:552-558: h = conv3x3(h); h += Dense(temb); h = act(normalize(h))), and conv2 / the first conv when the
 * next layer starts with a GroupNorm of that tensor (the next block's act(normalize(x)) at :552, an AttnBlock's normalize(x) at
 * :498, the final act(normalize(h)) of cifar/models/ddpm.py:98).  Where the tile shape allows it the 32-group statistics of the
 * whole image are formed inside the GEMM epilogue (channel sums are thread-local in the [channel][pixel] epilogue; the four CTAs of
 * one 32x32 image exchange theirs through global memory inside a cooperative launch -- or, with SDB_GN_GX=0 / on more than 8
 * streams, through distributed shared memory inside a thread-block cluster of 4; a unit of four 8x8 images keeps per-image sums;
 * pixel-row tiles that hold whole 8x8 / 4x4 images reduce over lane segments) and `out` receives [swish](GN(conv) * gamma + beta)
 * directly: *fused_host = 1.
 * raw_out (optional): a second output, the raw conv result (bf16, layout of `out`), for tensors that also stay on the residual
 * stream; `stats_out` then carries the raw tensor's channel sums (per 256-pixel unit: the second 128-pixel slot of a unit is 0).
 * Where the shape does not allow fusion (small batches, N != 256 at low resolution, split precision) the call behaves like
 * sd_conv_gemm and *fused_host = 0: the raw result (+ stats_out) lands in raw_out when it is given, else in `out`, and the caller
 * runs sd_groupnorm_swish itself. */
int sd_conv_gemm_gn(const sd_gemm_src* srcs_host, int num_srcs, int B, int H, int W, const void* Wt, int N,
                    const float* bias, const float* rowbias, int rowbias_ld, unsigned flags, void* out, int out_ld,
                    float* stats_out, const float* gn_gamma, const float* gn_beta, float gn_eps, int gn_swish,
                    void* raw_out, int* fused_host, void* stream);

/* How far sd_conv_gemm_gn may fuse: 0 = never (always the separate GroupNorm pass), 1 = only where one CTA / CTA pair holds the
 * image, 2 = also through thread-block clusters (default; environment SDB_GN_FUSE).  Process-wide; returns the previous level.
 * Callers that branch on *fused_host need nothing else; bench.py uses it to time the same launches both ways. */
int sd_set_gn_fuse(int level);

/* 3x3 stride-2 SAME conv of the Downsample block (cifar/models/layers.py:526-537; flax pads (0,1) on even sizes):
 * out[b,ho,wo,n] = sum_{kh,kw,c} x[b, 2ho+kh, 2wo+kw, c] * Wt[n, (kh*3+kw)*C + c] + bias[n], x = 0 outside.
 * The stride lives in the TMA descriptor (element strides 2 along w and h), so no im2col buffer is materialised.
 * x: bf16 [B,H_in,W_in,C]; out: bf16 [B,H_in/2,W_in/2,N]; stats_out as in sd_conv_gemm (output geometry). */
int sd_conv_gemm_s2(const void* x, int B, int H_in, int W_in, int C, const void* Wt, int N, const float* bias,
                    unsigned flags, void* out, float* stats_out, void* stream);

/* Nearest-neighbour x2 upsample followed by a 3x3 SAME conv (cifar/models/layers.py:514-523), fused and reduced:
 * output pixel (2i+a, 2j+b) only sees a 2x2 neighbourhood of the low-resolution input, so the layer is four
 * 2x2-tap implicit GEMMs (one per phase (a, b)) over x [B,H,W,C] with pre-summed weights -- 16/36 of the FLOPs of
 * convolving the upsampled tensor, and the upsampled tensor is never materialised.
 * Wt4: bf16 [4 phases][N][4 taps * C] (tap = r*2+c over source offsets (a-1+r, b-1+c)); out: bf16 [B,2H,2W,N];
 * stats_out (optional, H*W % 128 == 0): fp32 [B][4*H*W/128][2][N] channel sums for sd_groupnorm_swish. */
int sd_upconv_gemm(const void* x, int B, int H, int W, int C, const void* Wt4, int N, const float* bias,
                   unsigned flags, void* out, float* stats_out, void* stream);

/* Batched "NT" GEMM on the same tcgen05 kernel:
 *     out[b][m][n] = sum_k A[b][m][k] * Bt[b][n][k] + bias[n] + residual[b][m][n]
 * A: bf16 [batch][M][lda], Bt: bf16 [batch][N][ldb] (both K-contiguous); a batch
 * stride of 0 shares the operand across the batch.  Used for nn.Dense call sites
 * (batch = 1) and for the S = 256 attention products q k^T and p v of
 * cifar/models/layers.py:505-509.  K multiple of 64. */
int sd_batched_gemm(const void* A, int lda, long long strideA, const void* Bt, int ldb, long long strideB,
                    int batch, int M, int N, int K, const float* bias, const void* residual, unsigned flags,
                    void* out, int ldc, long long strideC, void* stream);

/* sd_batched_gemm that also emits per-128-row-tile channel sums of its output (as sd_conv_gemm's stats_out):
 * stats_out fp32 [batch][M/128][2][N]; needs M % 128 == 0 and N % 16 == 0.  Used for the fused attention tail
 * out = P (V Wo) + (b_v Wo + b_o) + x, whose output feeds the next GroupNorm (cifar/models/layers.py:509-511). */
int sd_batched_gemm_stats(const void* A, int lda, long long strideA, const void* Bt, int ldb, long long strideB,
                          int batch, int M, int N, int K, const float* bias, const void* residual, unsigned flags,
                          void* out, int ldc, long long strideC, float* stats_out, void* stream);

/* Fused attention core (cifar/models/layers.py:505-511 after the projections):
 *     out[b][i][:] = sum_j softmax_j(scale * <Q[b][i], K[b][j]>) V[b][j][:] + bias + residual[b][i][:]
 * with the softmax restricted to the diagonal block of `block` key columns row i belongs to (block < S packs S/block small
 * images into one batch entry).  Q, K: bf16 [batch][S][ld] (channel-contiguous), Vt: bf16 [batch][C][ldv] = V transposed
 * (key-contiguous), residual / out: bf16 [batch][S][C]; S in {128, 256}, C in {64, 128, 192, 256}.  The probability matrix
 * stays in shared memory (the softmax writes the UMMA operand of the second product).  C = 256 with block % 16 == 0 (the
 * score-net's shape) takes the two-tile software-pipelined kernel, in which the residual is accumulated on the tensor cores
 * (residual rows must be contiguous: [batch][S][C]).  stats_out (optional, block == S):
 * fp32 [batch][S/128][2][C] channel sums of out for sd_groupnorm_swish. */
int sd_attention_core(const void* Q, int ldq, long long strideQ, const void* K, int ldk, long long strideK,
                      const void* Vt, int ldv, long long strideV, int batch, int S, int C, float scale, int block,
                      const float* bias, const void* residual, void* out, float* stats_out, void* stream);

/* Attention probabilities in one launch (cifar/models/layers.py:505-507):
 *     P[b][i][:] = softmax_j(scale * <Q[b][i], K[b][j]>)   restricted to the diagonal block of `block` columns
 * that row i belongs to (block = S for ordinary attention; block < S packs S/block small images into one
 * batch entry so that low-resolution attention still fills 128-row tensor-core tiles; off-block entries are 0).
 * Q, K: bf16 [batch][S][ld] (K-contiguous, ld >= C), P: bf16 [batch][S][S]; S multiple of 16, <= 256.  The
 * score row never leaves tensor memory: max / sum / normalise run in the GEMM epilogue. */
int sd_attention_probs(const void* Q, int ldq, long long strideQ, const void* Kt, int ldk, long long strideK,
                       int batch, int S, int C, float scale, int block, void* P, void* stream);

/* Row softmax P[r,:] = softmax(scale * X[r,:]); X fp32 [rows, cols] -> P bf16
 * (jax.nn.softmax at cifar/models/layers.py:507 with the C^-1/2 scale of :505). */
int sd_softmax_rows(const float* x, void* out, long rows, int cols, float scale, void* stream);

/* FP32-faithful form: X fp32 [rows, cols] -> P as a hi|lo bf16 pair [rows, 2*cols] (SD_GEMM_SPLIT3 layout); the softmax runs
 * over the diagonal block of `block` columns that row (row % rows_per_entry) belongs to, other entries are 0 (packed
 * low-resolution images, as sd_attention_probs). */
int sd_softmax_rows_split(const float* x, void* out, long rows, int cols, float scale, int block, int rows_per_entry, void* stream);

/* GroupNorm(32 groups, eps) + optional swish over the channel-concatenation of
 * up to two NHWC bf16 tensors; writes bf16 [B,H,W,C0+C1].
 * Replaces act(normalize()(x)) at cifar/models/layers.py:552,557,498 and
 * cifar/models/ddpm.py:98 (flax nn.GroupNorm defaults, normalization.py:38-39). */
int sd_groupnorm_swish(const void* x0, int C0, const void* x1, int C1, int B, int HW,
                       const float* gamma, const float* beta, float eps, int apply_swish,
                       const float* stats0 /* optional [B][nchunk0][2][C0] channel sums from sd_conv_gemm's stats_out */, int nchunk0,
                       const float* stats1 /* same for x1, [B][nchunk1][2][C1] */, int nchunk1,
                       float* scratch /* >= 2*(4736+B)*(C0+C1) + 64*B floats: channel sums of sources without stats, group stats */,
                       size_t scratch_floats, void* out, void* stream);

/* Same with flags: SD_GEMM_SPLIT3 = sources and `out` are hi|lo bf16 pairs ([B,HW,2C], see the flag) and the swish is evaluated
 * exactly (v / (1 + exp(-v)) instead of the one-MUFU tanh form, whose 2^-11 error is below bf16 rounding but not below fp32). */
int sd_groupnorm_swish_ex(const void* x0, int C0, const void* x1, int C1, int B, int HW,
                          const float* gamma, const float* beta, float eps, int apply_swish,
                          const float* stats0, int nchunk0, const float* stats1, int nchunk1,
                          float* scratch, size_t scratch_floats, void* out, unsigned flags, void* stream);

/* Forward-mode derivative of sd_groupnorm_swish: given the primal sources and their tangents (same layouts) writes
 * out = act(GN(x)) and dout = d/dh act(GN(x + h*dx)) at h = 0, both bf16 [B,HW,C0+C1]:
 *     xhat = (x-mean)*rstd,  du = rstd*gamma*(dx - mean_g(dx) - xhat*mean_g(xhat*dx)),  dout = swish'(u)*du.
 * One link of jax.jvp through cifar/models/ddpm.py (reference cifar/dynamics.py:84).  scratch >= 4*B*64*(C0+C1) floats. */
int sd_groupnorm_swish_jvp(const void* x0, const void* dx0, int C0, const void* x1, const void* dx1, int C1, int B, int HW,
                           const float* gamma, const float* beta, float eps, int apply_swish, float* scratch,
                           size_t scratch_floats, void* out, void* dout, void* stream);

/* JVP of a row softmax from its output: dP = P * (dS - sum_k P_k dS_k) with dS = scale*(dS1 + dS2) (dS2 may be NULL).
 * P: bf16 [rows][cols], dS1/dS2: fp32, dP: bf16 (attention tangent, cifar/models/layers.py:505-509). */
int sd_softmax_jvp(const void* P, const float* dS1, const float* dS2, float scale, void* dP, long rows, int cols, void* stream);

/* Single-head self-attention over HW tokens (cifar/models/layers.py:505-509):
 * out = softmax_{HW}(q k^T * C^-1/2) v.  qkv: bf16 [B, S, 3C] (q | k | v
 * channel blocks), out: bf16 [B, S, C].  CUDA-core kernel for the low-resolution
 * blocks (S in {4, 16, 64}); S = 256 runs as sd_batched_gemm + sd_softmax_rows. */
int sd_attention(const void* qkv, int B, int S, int C, void* out, void* stream);

/* Nearest-neighbour x2 upsample of an NHWC bf16 tensor
 * (jax.image.resize 'nearest', cifar/models/layers.py:520). */
int sd_upsample2x(const void* x, int B, int H, int W, int C, void* out, void* stream);

/* Gather for the stride-2 SAME conv (cifar/models/layers.py:533, pad (0,1)):
 * out[b,ho,wo, tap*C + c] = x[b, 2ho+kh, 2wo+kw, c] (0 outside). bf16. */
int sd_im2col_s2(const void* x, int B, int H, int W, int C, void* out, void* stream);

/* out[0..row_floats) = table[row][:] with row = *counter (device int, clamped to [0, rows)): the per-timestep row of a
 * precomputed table inside a captured CUDA graph.  Used for the time-embedding biases: Dense_i(act(temb(t))) of
 * cifar/models/ddpm.py:64-66 + layers.py:556 depends only on t, so the sampler tabulates it once for its n_steps times. */
int sd_gather_row(const float* table, int rows, int row_floats, const int* counter, float* out, void* stream);

/* Operand gather for the first conv (cifar/models/ddpm.py:71) on the tensor cores: out[pixel] = 64 bf16 =
 * [hi(9*Cin) | lo(9*Cin) | zeros], hi = bf16(v), lo = bf16(v - hi) of the zero-padded 3x3xCin fp32 neighbourhood (Cin <= 3).
 * The conv is then sd_conv_gemm with one 1-tap source of 64 channels against weights [w | w | 0] (bf16 [Cout, 64]). */
int sd_im2col_in(const float* x, int B, int H, int W, int Cin, void* out /* bf16 [B,H,W,64] */, void* stream);

/* Same with flags: SD_GEMM_SPLIT3 writes 128-wide rows [block(64) | zeros(64)], i.e. the block as the hi half of a hi|lo pair,
 * for sd_conv_gemm(SD_GEMM_SPLIT3) against weights [Cout, 128] = [w_hi | w_hi | 0 || w_lo | 0 | 0]. */
int sd_im2col_in_ex(const float* x, int B, int H, int W, int Cin, void* out, unsigned flags, void* stream);

/* First conv (cifar/models/ddpm.py:71): fp32 NHWC [B,H,W,Cin<=4] -> bf16
 * [B,H,W,Cout], 3x3 SAME, fp32 weights [3,3,Cin,Cout] (Flax HWIO) + bias. */
int sd_conv_in(const float* x, int B, int H, int W, int Cin, const float* w_hwio, const float* bias,
               int Cout, void* out, void* stream);

/* Sinusoidal time embedding + Dense + swish + Dense (+ class embedding) and the
 * swish that feeds every ResBlock's Dense (cifar/models/ddpm.py:64-68,
 * layers.py:450-461,556): writes act_temb = swish(temb) as bf16 [B, 4nf].
 * t comes from t_dev[0] scaled per sample, or from sched/step_counter
 * (sigma column == t for this SDE). */
int sd_time_embedding(const float* t_dev, int t_stride, const float* sched, const int* step_counter,
                      int B, int nf,
                      const float* w0 /*[nf,4nf]*/, const float* b0, const float* w1 /*[4nf,4nf]*/, const float* b1,
                      const float* class_emb /*[ncls,4nf] or NULL*/, const int* labels,
                      float* temb_scratch /* fp32 [B,4nf] ([1,4nf] when t is shared) */,
                      void* act_temb_out /* bf16 [B,4nf] */, void* stream);

/* Same with flags: SD_GEMM_SPLIT3 = exact swish everywhere and act_temb_out as a hi|lo bf16 pair [B, 2*4nf]. */
int sd_time_embedding_ex(const float* t_dev, int t_stride, const float* sched, const int* step_counter,
                         int B, int nf, const float* w0, const float* b0, const float* w1, const float* b1,
                         const float* class_emb, const int* labels, float* temb_scratch, void* act_temb_out,
                         unsigned flags, void* stream);

/* fp32 -> bf16 and bf16 -> fp32 converts (weights upload, debugging). */
int sd_cast_f32_to_bf16(const float* in, void* out, size_t n, void* stream);
int sd_cast_bf16_to_f32(const void* in, float* out, size_t n, void* stream);

/* ------------------------------------------------------------------------
 * Whole score network as one call (the reference's model seam: model_fn(t, x, y) from get_model_fn,
 * cifar/models/utils.py:86-96, over ScoreNet.__call__, cifar/models/ddpm.py:47-101).
 * The network structure follows the reference's config fields (configs/sm/cifar/vpsde.py: model.nf, ch_mult,
 * num_res_blocks, attn_resolutions, conditioned; data.image_size, num_channels, num_classes); the weights are ONE device
 * blob in GEMM layout, packed from a reference parameter tree by super_diffusion_b200/native.py (order and shapes:
 * csrc/scorenet_forward.cu header).  Runs the same kernels in the same order as the Python-driven forward.
 * ------------------------------------------------------------------------ */
typedef struct sd_scorenet_desc {
  int image_size;              /* config.data.image_size (multiple of 16) */
  int channels;                /* config.data.num_channels (<= 3) */
  int nf;                      /* config.model.nf (multiple of 64) */
  int num_res_blocks;          /* config.model.num_res_blocks */
  int n_levels;                /* len(config.model.ch_mult), <= 8 */
  int ch_mult[8];
  int n_attn_res;              /* len(config.model.attn_resolutions), <= 8 */
  int attn_resolutions[8];
  int conditioned;             /* config.model.conditioned */
  int num_classes;             /* config.data.num_classes (conditioned models) */
  const void* weights;         /* device blob, 256-byte aligned */
  size_t weights_bytes;        /* must equal sd_scorenet_weights_bytes() */
  int precision;               /* SD_PRECISION_*; 0 = SD_PRECISION_BF16.  Decides the blob layout (GEMM weights are [N, K] bf16
                                  or [N, 2K] hi|lo pairs) and the workspace size. */
} sd_scorenet_desc;

#define SD_PRECISION_BF16 1           /* bf16 operands / activations, fp32 accumulation */
#define SD_PRECISION_FP32_FAITHFUL 2  /* the reference's fp32 arithmetic (cifar/models/ddpm.py runs in fp32) to ~1e-5: every
                                         operand a hi|lo bf16 pair, products hi*hi + lo*hi + hi*lo on the tensor cores with fp32
                                         accumulation (SD_GEMM_SPLIT3), GroupNorm / swish / softmax in fp32; ~3x the tensor work */

/* size of the weight blob for a configuration (desc->weights is ignored) */
int sd_scorenet_weights_bytes(const sd_scorenet_desc* desc, size_t* bytes_out);
/* workspace a forward at batch B needs (activation arena, 1.5 GiB at batch 512 for the CIFAR configuration; t_stride as in
 * sd_scorenet_forward) */
int sd_scorenet_workspace_bytes(const sd_scorenet_desc* desc, int B, int t_stride, size_t* bytes_out);
/* out_nhwc[B,H,W,C] (fp32) = model_fn(t, x, y) = sigma_t * grad log q_t(x).
 * t_dev: device fp32, one value (t_stride = 0) or one per sample (t_stride = 1); x_nhwc: fp32 [B,H,W,C];
 * y: int32 labels [B] (conditioned models) or NULL.  Asynchronous on `stream`, no host synchronisation. */
int sd_scorenet_forward(const sd_scorenet_desc* desc, const float* t_dev, int t_stride, const float* x_nhwc, const int* y,
                        int B, float* out_nhwc, void* workspace, size_t workspace_bytes, int precision, void* stream);

/* Same network with the time read on the device: t = sched[*step_counter].sigma (rows (a_t, b_t, sigma_t, dt) as in
 * sd_step_vpsde; sigma_t = t for this SDE).  Every argument is then fixed across timesteps, so one captured CUDA graph of
 * M x sd_scorenet_forward_sched + sd_step_vpsde(sched, step_counter) + sd_counter_add replays for the whole sampling loop. */
int sd_scorenet_forward_sched(const sd_scorenet_desc* desc, const float* sched, const int* step_counter, const float* x_nhwc,
                              const int* y, int B, float* out_nhwc, void* workspace, size_t workspace_bytes, int precision,
                              void* stream);

const char* sd_last_error(void);
int sd_version(void);
/* 1 when the current device is compute capability 10.x (sm_100 family). */
int sd_device_ok(void);

#ifdef __cplusplus
}
#endif
#endif /* SUPERDIFF_B200_H_ */
