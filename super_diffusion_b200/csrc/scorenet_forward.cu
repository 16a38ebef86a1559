// Native whole-network entry: sd_scorenet_forward runs the reference's DDPM U-Net (cifar/models/ddpm.py:47-101) as one
// C call - the op plan that super_diffusion_b200/models/ddpm.py drives from Python, restated in C++ over the same C-ABI
// kernels, so a host written in any language gets the score network with three calls (weights_bytes / workspace_bytes /
// forward) and no Python.  Same kernels in the same order with the same arguments: the output is bit-identical to the
// Python-driven forward (tests/test_scorenet_native_gpu.py).
//
// Weights: one device blob in the order `layout()` walks (super_diffusion_b200/native.py packs a bound model into it):
//   temb Dense_0 / Dense_1 (fp32) | class embedding (fp32, conditioned models) | first conv as a [nf, 64] bf16 K-block + bias |
//   concatenated per-block Dense(temb) weights bf16 [sum cout, 4nf] + biases (conv bias folded in) |
//   per ResnetBlock: GN0 scale, bias | conv0 bf16 [cout, 9 cin] | GN1 scale, bias | conv1 (+ NIN shortcut or identity) bf16
//     [cout, 9 cout + cin] | bias   |  per AttnBlock (projections folded, see models/ddpm.py::add_attn): GN scale, bias |
//     Wq Wk^T bf16 [c, c] + Wk bq | (Wv Wo)^T bf16 [c, c] + bv Wo + bo  |  per Downsample: bf16 [c, 9c] + bias  |
//   per Upsample: four 2x2-tap phase matrices bf16 [4, c, 4c] + bias  |  output GN scale, bias | output conv bf16 [16, 9 nf] + bias.
// Every array starts on a 256-byte boundary.  Workspace: a best-fit arena; a block is reused as soon as its last reader is enqueued.
#include "common.cuh"
#include "../../include/superdiff_b200.h"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace sdb {
namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct ResW { const float *g1, *be1, *g2, *be2, *b2; const void *w1, *w2; int cin, cout, off, nparts, parts[2]; };
struct AttnW { const float *g, *be, *b_q2, *b_vo; const void *w_q2, *w_voT; int c; };
struct DownW { const void* w; const float* b; int c; };
struct UpW { const void* w4; const float* b; int c; };
enum OpKind { OP_DOWN_BLOCK, OP_DOWNSAMPLE, OP_MID, OP_UP_BLOCK, OP_ATTN, OP_UPSAMPLE };
struct Op { OpKind kind; int a, b, c; };

struct Net {
  const float *temb_w0, *temb_b0, *temb_w1, *temb_b1, *class_emb, *conv_in_b, *dense_b, *out_g, *out_be, *out_b;
  const void *conv_in_w64, *dense_w, *out_w;
  int dense_n, out_c;
  std::vector<ResW> res;
  std::vector<AttnW> attn;
  std::vector<DownW> down;
  std::vector<UpW> up;
  std::vector<Op> plan;
  size_t bytes;
};

// Walks the blob in its canonical order; with base == nullptr only the size is computed.
struct Cursor {
  const char* base;
  size_t off = 0;
  const void* take(size_t bytes) {
    const void* p = base ? base + off : nullptr;
    off = align_up(off + bytes);
    return p;
  }
  const float* f32(size_t n) { return static_cast<const float*>(take(n * 4)); }
  const void* bf16(size_t n) { return take(n * 2); }
  int kmul = 1;          // 2 in the FP32-faithful arm: every GEMM weight is [N, 2K] = [hi | lo] (SD_GEMM_SPLIT3)
  const void* gemm_w(size_t n) { return take(n * 2 * kmul); }
};

inline bool desc_split(const sd_scorenet_desc& d) { return d.precision == SD_PRECISION_FP32_FAITHFUL; }

bool in_list(const int* v, int n, int x) {
  for (int i = 0; i < n; ++i)
    if (v[i] == x) return true;
  return false;
}

// Structure of the network (control flow of cifar/models/ddpm.py:70-99) and the blob offsets of every array.
int layout(const sd_scorenet_desc& d, Net& net) {
  if (d.n_levels < 1 || d.n_levels > 8 || d.num_res_blocks < 1 || d.nf < 64 || d.nf % 64 || d.channels < 1 || d.channels > 3 ||
      d.image_size < 16 || d.image_size % 16 || d.n_attn_res < 0 || d.n_attn_res > 8)
    return fail(kErrUnsupported, "sd_scorenet: unsupported configuration (nf multiple of 64, <= 3 image channels, image size multiple of 16)");
  if (d.conditioned && d.num_classes < 1) return fail(kErrInvalidArg, "sd_scorenet: conditioned model needs num_classes");
  if (d.precision != 0 && d.precision != SD_PRECISION_BF16 && d.precision != SD_PRECISION_FP32_FAITHFUL)
    return fail(kErrUnsupported, "sd_scorenet: desc.precision must be 0 (= bf16), SD_PRECISION_BF16 or SD_PRECISION_FP32_FAITHFUL");
  const int nf = d.nf;
  Cursor cur{static_cast<const char*>(d.weights)};
  cur.kmul = desc_split(d) ? 2 : 1;
  net.temb_w0 = cur.f32((size_t)nf * 4 * nf); net.temb_b0 = cur.f32(4 * nf);
  net.temb_w1 = cur.f32((size_t)4 * nf * 4 * nf); net.temb_b1 = cur.f32(4 * nf);
  net.class_emb = d.conditioned ? cur.f32((size_t)d.num_classes * 4 * nf) : nullptr;
  net.conv_in_w64 = cur.gemm_w((size_t)nf * 64); net.conv_in_b = cur.f32(nf);

  // pass 1: block shapes in creation order (add_res / add_attn of models/ddpm.py)
  int off = 0;
  auto add_res = [&](int c0, int c1, int cout) {
    ResW r{};
    r.nparts = c1 ? 2 : 1; r.parts[0] = c0; r.parts[1] = c1; r.cin = c0 + c1; r.cout = cout; r.off = off;
    off += cout;
    net.res.push_back(r);
    return (int)net.res.size() - 1;
  };
  auto add_attn = [&](int c) { AttnW a{}; a.c = c; net.attn.push_back(a); return (int)net.attn.size() - 1; };
  int size = d.image_size, c = nf;
  std::vector<int> chans{nf};
  for (int lvl = 0; lvl < d.n_levels; ++lvl) {
    for (int i = 0; i < d.num_res_blocks; ++i) {
      const int ri = add_res(c, 0, nf * d.ch_mult[lvl]);
      c = nf * d.ch_mult[lvl];
      const int ai = in_list(d.attn_resolutions, d.n_attn_res, size) ? add_attn(c) : -1;
      net.plan.push_back({OP_DOWN_BLOCK, ri, ai, 0});
      chans.push_back(c);
    }
    if (lvl != d.n_levels - 1) {
      net.down.push_back({nullptr, nullptr, c});
      net.plan.push_back({OP_DOWNSAMPLE, (int)net.down.size() - 1, 0, 0});
      size /= 2;
      chans.push_back(c);
    }
  }
  {
    const int r0 = add_res(c, 0, c), a0 = add_attn(c), r1 = add_res(c, 0, c);
    net.plan.push_back({OP_MID, r0, a0, r1});
  }
  for (int lvl = d.n_levels - 1; lvl >= 0; --lvl) {
    for (int i = 0; i < d.num_res_blocks + 1; ++i) {
      const int skip = chans.back();
      chans.pop_back();
      const int ri = add_res(c, skip, nf * d.ch_mult[lvl]);
      c = nf * d.ch_mult[lvl];
      net.plan.push_back({OP_UP_BLOCK, ri, 0, 0});
    }
    if (in_list(d.attn_resolutions, d.n_attn_res, size)) net.plan.push_back({OP_ATTN, add_attn(c), 0, 0});
    if (lvl != 0) {
      net.up.push_back({nullptr, nullptr, c});
      net.plan.push_back({OP_UPSAMPLE, (int)net.up.size() - 1, 0, 0});
      size *= 2;
    }
  }
  if (!chans.empty()) return fail(kErrInvalidArg, "sd_scorenet: inconsistent skip bookkeeping");
  net.dense_n = off;
  net.out_c = c;

  // pass 2: blob offsets
  net.dense_w = cur.gemm_w((size_t)off * 4 * nf); net.dense_b = cur.f32(off);
  for (ResW& r : net.res) {
    r.g1 = cur.f32(r.cin); r.be1 = cur.f32(r.cin);
    r.w1 = cur.gemm_w((size_t)r.cout * 9 * r.cin);
    r.g2 = cur.f32(r.cout); r.be2 = cur.f32(r.cout);
    // conv1 + shortcut: NIN over the block input when C_in != C_out (layers.py:560-564), identity segment otherwise (:565)
    r.w2 = cur.gemm_w((size_t)r.cout * (9 * r.cout + r.cin));
    r.b2 = cur.f32(r.cout);
    if (r.cin == r.cout && r.nparts != 1) return fail(kErrUnsupported, "sd_scorenet: identity shortcut over a concatenation");
  }
  for (AttnW& a : net.attn) {
    a.g = cur.f32(a.c); a.be = cur.f32(a.c);
    a.w_q2 = cur.gemm_w((size_t)a.c * a.c); a.b_q2 = cur.f32(a.c);
    a.w_voT = cur.gemm_w((size_t)a.c * a.c); a.b_vo = cur.f32(a.c);
  }
  for (DownW& w : net.down) { w.w = cur.gemm_w((size_t)w.c * 9 * w.c); w.b = cur.f32(w.c); }
  for (UpW& w : net.up) { w.w4 = cur.gemm_w((size_t)4 * w.c * 4 * w.c); w.b = cur.f32(w.c); }
  net.out_g = cur.f32(c); net.out_be = cur.f32(c);
  net.out_w = cur.gemm_w((size_t)16 * 9 * c); net.out_b = cur.f32(d.channels);
  net.bytes = cur.off;
  return SD_OK;
}

// Activation in the workspace: NHWC bf16 plus the per-128-pixel-tile channel sums its producer emitted (if any).
struct Act {
  void* p = nullptr; int H = 0, W = 0, C = 0; float* stats = nullptr; int nchunk = 0;
  // GroupNorm of THIS tensor pre-computed (or to be computed) for its first consumer: norm_key = that GroupNorm's scale vector;
  // norm_p = its output buffer (== p when the tensor itself was normalised in place: no other reader of the raw form);
  // norm_valid = the producing GEMM's epilogue already filled it (sd_conv_gemm_gn fused)
  void* norm_p = nullptr; const float* norm_key = nullptr; bool norm_valid = false;
};
// the GroupNorm that consumes a conv's output first (models/ddpm.py::gn_after)
struct GnNext { const float* gamma = nullptr; const float* beta = nullptr; bool swish = true; bool keep_raw = true; };

struct Runner {
  const sd_scorenet_desc& d;
  const Net& net;
  int B;
  char* ws;             // nullptr: dry run, only the workspace size is computed
  size_t ws_bytes, used = 0;
  cudaStream_t st;
  float* gn_scratch = nullptr;
  size_t gn_scratch_floats = 0;
  int rc = SD_OK;
  // FP32-faithful arm: activations are hi|lo bf16 pairs [.., 2C] and every kernel gets SD_GEMM_SPLIT3 (see the header); the
  // attention runs unfused (fp32 scores, fp32 softmax).  Same plan otherwise.
  bool split() const { return desc_split(d); }
  int sm() const { return split() ? 2 : 1; }
  unsigned fl() const { return split() ? SD_GEMM_SPLIT3 : 0u; }

  // Workspace allocator: best-fit from the blocks released so far, else bump.  Everything is enqueued on ONE stream, so a block
  // may be handed out again as soon as the launches that read it have been enqueued.  Reuse matters for speed, not only for
  // size: with pure bump allocation (6.6 GiB of distinct addresses at batch 512) the same kernels ran 3-6 % slower than the
  // Python-driven plan, whose caching allocator keeps re-using a few hot buffers.
  struct Block { size_t off, size; };
  std::vector<Block> free_list;
  static size_t ws_align() {     // SDB_NATIVE_WS_ALIGN: granularity of workspace blocks (tuning knob)
    static const size_t v = [] { const char* e = getenv("SDB_NATIVE_WS_ALIGN"); const long a = e && *e ? atol(e) : 0; return a >= 256 ? (size_t)a : kAlign; }();
    return v;
  }
  void* alloc(size_t bytes) {
    bytes = (bytes ? bytes : 1);
    bytes = (bytes + ws_align() - 1) / ws_align() * ws_align();
    int best = -1;
    for (int i = 0; i < (int)free_list.size(); ++i)
      if (free_list[i].size >= bytes && (best < 0 || free_list[i].size < free_list[best].size)) best = i;
    size_t off;
    if (best >= 0) {
      off = free_list[best].off;
      if (free_list[best].size == bytes) free_list.erase(free_list.begin() + best);
      else { free_list[best].off += bytes; free_list[best].size -= bytes; }
    } else {
      off = used;
      used += bytes;
      if (ws && used > ws_bytes && rc == SD_OK) rc = fail(kErrInvalidArg, "sd_scorenet_forward: workspace too small (see sd_scorenet_workspace_bytes)");
    }
    sizes.push_back({off, bytes});
    return ws ? ws + off : reinterpret_cast<char*>(kAlign) + off;      // dry run: fake addresses keep the bookkeeping identical
  }
  std::vector<Block> sizes;                     // live allocations (offset -> size)
  void release(const void* p) {
    if (!p) return;
    const size_t off = (size_t)(static_cast<const char*>(p) - (ws ? ws : reinterpret_cast<char*>(kAlign)));
    for (int i = (int)sizes.size() - 1; i >= 0; --i)
      if (sizes[i].off == off) {
        Block b = sizes[i];
        sizes.erase(sizes.begin() + i);
        // coalesce with neighbours
        for (int j = 0; j < (int)free_list.size();) {
          if (free_list[j].off + free_list[j].size == b.off) { b.off = free_list[j].off; b.size += free_list[j].size; free_list.erase(free_list.begin() + j); }
          else if (b.off + b.size == free_list[j].off) { b.size += free_list[j].size; free_list.erase(free_list.begin() + j); }
          else ++j;
        }
        free_list.push_back(b);
        return;
      }
  }
  void release(Act& a) { release(a.p); release(a.stats); a.p = nullptr; a.stats = nullptr; }
  bool live() const { return ws != nullptr && rc == SD_OK; }
  void check(int r) { if (rc == SD_OK && r != SD_OK) rc = r; }

  // every activation has room for the batch rounded up to a multiple of 8 images: low-resolution attention packs up to 8 images
  // per 128-row tile and pads the last tile with zero images (attn_block)
  int Bp() const { return (B + 7) / 8 * 8; }
  Act new_act(int H, int W, int C) { Act a; a.H = H; a.W = W; a.C = C; a.p = alloc((size_t)Bp() * H * W * C * 2 * sm()); return a; }
  void want_stats(Act& a, int tiles_per_img) {
    if ((a.H * a.W) % 128 == 0 && a.C % 16 == 0 && B > 0) {
      a.nchunk = tiles_per_img;
      a.stats = static_cast<float*>(alloc((size_t)B * tiles_per_img * 2 * a.C * 4));
    }
  }

  // act(GroupNorm(concat(x0, x1)))  (layers.py:552,557,498; ddpm.py:98)
  Act gn(const Act& x0, const Act* x1, const float* gamma, const float* beta, bool swish) {
    Act o = new_act(x0.H, x0.W, x0.C + (x1 ? x1->C : 0));
    if (live())
      check(sd_groupnorm_swish_ex(x0.p, x0.C, x1 ? x1->p : nullptr, x1 ? x1->C : 0, B, x0.H * x0.W, gamma, beta, 1e-6f, swish ? 1 : 0,
                                  x0.stats, x0.nchunk, x1 ? x1->stats : nullptr, x1 ? x1->nchunk : 0, gn_scratch, gn_scratch_floats,
                                  o.p, fl(), st));
    return o;
  }

  // act(GroupNorm(x)) for x's first consumer: the buffer its producer reserved (already filled when the producer's epilogue
  // fused the normalisation), else a separate pass.  The returned Act owns the buffer like any GroupNorm output.
  Act gn_first(Act& x, const float* gamma, const float* beta, bool swish) {
    if (x.norm_p && x.norm_key == gamma && x.norm_p != x.p) {
      Act o; o.H = x.H; o.W = x.W; o.C = x.C; o.p = x.norm_p;
      if (!x.norm_valid && live())
        check(sd_groupnorm_swish_ex(x.p, x.C, nullptr, 0, B, x.H * x.W, gamma, beta, 1e-6f, swish ? 1 : 0, x.stats, x.nchunk, nullptr, 0,
                                    gn_scratch, gn_scratch_floats, o.p, fl(), st));
      x.norm_p = nullptr;
      return o;
    }
    return gn(x, nullptr, gamma, beta, swish);
  }

  // conv whose output's first consumer is the GroupNorm `next` (may be null): sd_conv_gemm_gn with the raw tensor kept as a second
  // output (keep_raw) or normalised in place.  The normalised buffer is reserved whether or not the launch fuses, so the arena sees
  // one allocation sequence (the workspace dry run cannot know the tile shape).
  Act conv_next(const sd_gemm_src* srcs, int nsrc, int H, int W, const void* Wt, int N, const float* bias, const GnNext* next) {
    if (!next || !next->gamma) return conv(srcs, nsrc, H, W, Wt, N, bias, nullptr, 0, true);
    Act o = new_act(H, W, N);                 // normalised output (keep_raw) or the only output
    Act raw;
    if (next->keep_raw) raw = new_act(H, W, N);
    Act& st_owner = next->keep_raw ? raw : o;
    want_stats(st_owner, H * W / 128);
    int fused = 0;
    if (live())
      check(sd_conv_gemm_gn(srcs, nsrc, B, H, W, Wt, N, bias, nullptr, 0, fl(), o.p, N * sm(), st_owner.stats, next->gamma, next->beta, 1e-6f,
                            next->swish ? 1 : 0, next->keep_raw ? raw.p : nullptr, &fused, st));
    if (next->keep_raw) {
      raw.norm_p = o.p; raw.norm_key = next->gamma; raw.norm_valid = fused != 0;
      return raw;
    }
    if (fused) { o.norm_p = o.p; o.norm_key = next->gamma; o.norm_valid = true; release(o.stats); o.stats = nullptr; }
    return o;
  }

  Act conv(const sd_gemm_src* srcs, int nsrc, int H, int W, const void* Wt, int N, const float* bias, const float* rowbias,
           int rb_ld, bool stats) {
    Act o = new_act(H, W, N);
    if (stats) want_stats(o, H * W / 128);
    if (live()) check(sd_conv_gemm(srcs, nsrc, B, H, W, Wt, N, bias, rowbias, rb_ld, nullptr, fl(), o.p, N * sm(), o.stats, st));
    return o;
  }

  // ResnetBlockDDPM (layers.py:540-565): GN+swish -> conv3x3 + temb bias -> GN+swish -> conv3x3 + shortcut.  The second GroupNorm
  // runs inside conv1's epilogue where the tile shape allows it (sd_conv_gemm_gn): 3 launches instead of 4.
  Act res_block(Act& x0, const Act* x1, int i, const float* rowbias, const GnNext* next = nullptr) {
    const ResW& r = net.res[i];
    Act a1 = x1 ? gn(x0, x1, r.g1, r.be1, true) : gn_first(x0, r.g1, r.be1, true);
    sd_gemm_src s1[1] = {{a1.p, a1.C, 9}};
    Act h1 = new_act(x0.H, x0.W, r.cout);
    want_stats(h1, x0.H * x0.W / 128);
    int fused = 0;
    if (live())
      check(sd_conv_gemm_gn(s1, 1, B, x0.H, x0.W, r.w1, r.cout, nullptr, rowbias ? rowbias + r.off : nullptr, net.dense_n, fl(), h1.p,
                            r.cout * sm(), h1.stats, r.g2, r.be2, 1e-6f, 1, nullptr, &fused, st));
    release(a1);
    Act a2;
    if (fused) {
      // h1 already holds act(normalize(conv1)).  The buffer the separate GroupNorm pass would have written is still taken and given
      // back, so the arena sees the same allocation sequence as the workspace dry run (which cannot know the tile shape).
      Act dummy = new_act(h1.H, h1.W, h1.C);
      release(dummy);
      a2 = h1;
      release(a2.stats);
      a2.stats = nullptr;
    } else {
      a2 = gn(h1, nullptr, r.g2, r.be2, true);
      release(h1);
    }
    sd_gemm_src s2[3] = {{a2.p, a2.C, 9}, {x0.p, x0.C, 1}, {x1 ? x1->p : nullptr, x1 ? x1->C : 0, 1}};
    Act o = conv_next(s2, x1 ? 3 : 2, x0.H, x0.W, r.w2, r.cout, r.b2, next);
    release(a2);
    return o;
  }

  // AttnBlock (layers.py:493-511) with the projections folded at export time: GN -> q' = NIN(h) -> V'^T = (Wv Wo)^T h^T ->
  // fused softmax(q' h^T) V' + bias + x
  Act attn_block(Act& x, int i) {
    const AttnW& a = net.attn[i];
    const int S = x.H * x.W, C = x.C;
    const int g = S >= 128 ? 1 : 128 / S;
    // a batch that does not fill its last tile (B % g != 0, e.g. the reference's eval batch of 100 at the 4x4 level) is padded
    // with zero images: the block-diagonal softmax keeps images independent, and zero keys / values keep the 0 * v products of
    // the masked blocks finite.  The padded output rows land in the spare room every activation has (new_act) and are ignored.
    const int Sp = g * S, nb = (B + g - 1) / g;
    if (S < 16 || !(Sp == 128 || Sp == 256) || C % 64 || C > 256) {
      if (rc == SD_OK)
        rc = fail(kErrUnsupported, "sd_scorenet_forward: attention needs 16..256 pixels per image and C <= 256 (multiple of 64)");
      return x;
    }
    Act h = gn_first(x, a.g, a.be, false);
    if (live() && nb * g > B)
      check(check_cuda(cudaMemsetAsync(static_cast<char*>(h.p) + (size_t)B * S * C * 2 * sm(), 0, (size_t)(nb * g - B) * S * C * 2 * sm(), st),
                       "sd_scorenet_forward: zeroing the attention padding"));
    sd_gemm_src sq[1] = {{h.p, C, 1}};
    Act q2 = conv(sq, 1, x.H, x.W, a.w_q2, C, a.b_q2, nullptr, 0, false);
    if (split()) {
      // five launches: V'^T, fp32 scores, fp32 block softmax -> hi|lo probabilities, P V' + bias + x (all products 3 x bf16)
      void* vt = alloc((size_t)nb * C * Sp * 2 * 2);
      float* sc = static_cast<float*>(alloc((size_t)nb * Sp * Sp * 4));
      void* pr = alloc((size_t)nb * Sp * Sp * 2 * 2);
      Act o = new_act(x.H, x.W, C);
      if (g == 1) { o.nchunk = Sp / 128; o.stats = static_cast<float*>(alloc((size_t)nb * o.nchunk * 2 * C * 4)); }
      if (live()) {
        const long long sA = (long long)Sp * 2 * C;
        check(sd_batched_gemm(a.w_voT, 2 * C, 0, h.p, 2 * C, sA, nb, C, Sp, C, nullptr, nullptr, SD_GEMM_SPLIT3, vt, 2 * Sp,
                              (long long)C * 2 * Sp, st));
        check(sd_batched_gemm(q2.p, 2 * C, sA, h.p, 2 * C, sA, nb, Sp, Sp, C, nullptr, nullptr, SD_GEMM_SPLIT3 | SD_EPI_OUT_F32, sc, Sp,
                              (long long)Sp * Sp, st));
        check(sd_softmax_rows_split(sc, pr, (long)nb * Sp, Sp, (float)std::pow((double)C, -0.5), S, Sp, st));
        check(sd_batched_gemm_stats(pr, 2 * Sp, (long long)Sp * 2 * Sp, vt, 2 * Sp, (long long)C * 2 * Sp, nb, Sp, C, Sp, a.b_vo, x.p,
                                    SD_GEMM_SPLIT3, o.p, 2 * C, sA, o.stats, st));
      }
      release(h); release(q2); release(vt); release(sc); release(pr);
      return o;
    }
    void* vt = alloc((size_t)nb * C * Sp * 2);
    Act o = new_act(x.H, x.W, C);
    if (g == 1) { o.nchunk = Sp / 128; o.stats = static_cast<float*>(alloc((size_t)nb * o.nchunk * 2 * C * 4)); }
    if (live()) {
      check(sd_batched_gemm(a.w_voT, C, 0, h.p, C, (long long)Sp * C, nb, C, Sp, C, nullptr, nullptr, 0u, vt, Sp, (long long)C * Sp, st));
      check(sd_attention_core(q2.p, C, (long long)Sp * C, h.p, C, (long long)Sp * C, vt, Sp, (long long)C * Sp, nb, Sp, C,
                              (float)std::pow((double)C, -0.5), S, a.b_vo, x.p, o.p, o.stats, st));
    }
    release(h); release(q2); release(vt);
    return o;
  }

  const float* sched = nullptr;         // optional device schedule table + step counter instead of t_dev (sd_scorenet_forward_sched)
  const int* step_counter = nullptr;
  void run(const float* t_dev, int t_stride, const float* x, const int* y, float* out) {
    const int nf = d.nf, H0 = d.image_size;
    int max_c = nf;                       // widest GroupNorm input (a skip concatenation)
    for (const ResW& r : net.res) max_c = r.cin > max_c ? r.cin : (r.cout > max_c ? r.cout : max_c);
    gn_scratch_floats = (size_t)(4736 + B) * 2 * (size_t)max_c + (size_t)64 * B;     // one buffer, reused in stream order
    gn_scratch = static_cast<float*>(alloc(gn_scratch_floats * 4));
    // time embedding (ddpm.py:64-68) -> per-sample bias of every block's Dense(temb) (layers.py:556): one [B, 4nf] x [4nf, sum cout] GEMM
    const bool shared_t = t_stride == 0;
    float* temb_scratch = static_cast<float*>(alloc((size_t)(shared_t ? 1 : B) * 4 * nf * 4));
    void* act_temb = alloc((size_t)B * 4 * nf * 2 * sm());
    float* rowbias = static_cast<float*>(alloc((size_t)B * net.dense_n * 4));
    if (live()) {
      check(sd_time_embedding_ex(t_dev, t_stride, sched, step_counter, B, nf, net.temb_w0, net.temb_b0, net.temb_w1, net.temb_b1,
                                 net.class_emb, net.class_emb ? y : nullptr, temb_scratch, act_temb, fl(), st));
      check(sd_batched_gemm(act_temb, 4 * nf * sm(), 0, net.dense_w, 4 * nf * sm(), 0, 1, B, net.dense_n, 4 * nf, net.dense_b, nullptr,
                            SD_EPI_OUT_F32 | fl(), rowbias, net.dense_n, (long long)B * net.dense_n, st));
    }
    // first conv (ddpm.py:71) on the tensor cores: hi/lo-split im2col K-block
    void* cols = alloc((size_t)B * H0 * H0 * 64 * 2 * sm());
    if (live()) check(sd_im2col_in_ex(x, B, H0, H0, d.channels, cols, fl(), st));
    // the GroupNorm that first consumes the tensor produced by plan[k] (stage: which ResBlock of a mid op; k = -1: the first conv);
    // gamma == nullptr when that consumer is not a plain GroupNorm of this tensor alone -- same rule as models/ddpm.py::gn_after
    const std::vector<Op>& plan = net.plan;
    auto gn_after = [&](int k, int stage) {
      GnNext g;
      auto attn_gn = [&](int ai) { g.gamma = net.attn[ai].g; g.beta = net.attn[ai].be; g.swish = false; g.keep_raw = true; };
      if (k >= 0) {
        const Op& op = plan[k];
        if (op.kind == OP_DOWN_BLOCK && op.b >= 0) { attn_gn(op.b); return g; }
        if (op.kind == OP_MID) { if (stage == 0) attn_gn(op.b); return g; }
        if (op.kind == OP_UP_BLOCK) {
          if (k + 1 == (int)plan.size()) { g.gamma = net.out_g; g.beta = net.out_be; g.swish = true; g.keep_raw = false; }
          else if (plan[k + 1].kind == OP_ATTN) attn_gn(plan[k + 1].a);
          return g;
        }
      }
      if (k + 1 < (int)plan.size() && (plan[k + 1].kind == OP_DOWN_BLOCK || plan[k + 1].kind == OP_MID) &&
          (k < 0 || plan[k].kind == OP_DOWN_BLOCK)) {
        const ResW& r = net.res[plan[k + 1].a];
        g.gamma = r.g1; g.beta = r.be1; g.swish = true; g.keep_raw = true;
      }
      return g;
    };
    sd_gemm_src s0[1] = {{cols, 64, 1}};
    GnNext n0 = gn_after(-1, 0);
    Act h = conv_next(s0, 1, H0, H0, net.conv_in_w64, nf, net.conv_in_b, &n0);
    release(cols);
    std::vector<Act> hs{h};
    for (int k = 0; k < (int)plan.size(); ++k) {
      const Op& op = plan[k];
      if (rc != SD_OK) return;
      switch (op.kind) {
        case OP_DOWN_BLOCK: {
          GnNext nx = gn_after(k, 0);
          h = res_block(hs.back(), nullptr, op.a, rowbias, &nx);
          if (op.b >= 0) { Act r = h; h = attn_block(r, op.b); if (h.p != r.p) release(r); }
          hs.push_back(h);
          break;
        }
        case OP_DOWNSAMPLE: {
          const DownW& w = net.down[op.a];
          const Act& src = hs.back();
          Act o = new_act(src.H / 2, src.W / 2, w.c);
          want_stats(o, o.H * o.W / 128);
          if (live()) check(sd_conv_gemm_s2(src.p, B, src.H, src.W, src.C, w.w, w.c, w.b, fl(), o.p, o.stats, st));
          h = o;
          hs.push_back(h);
          break;
        }
        case OP_MID: {
          GnNext n1 = gn_after(k, 0), n2 = gn_after(k, 1);
          Act r0 = res_block(hs.back(), nullptr, op.a, rowbias, &n1);        // its input stays alive as a skip connection
          Act a0 = attn_block(r0, op.b);
          if (a0.p != r0.p) release(r0);
          h = res_block(a0, nullptr, op.c, rowbias, &n2);
          release(a0);
          break;
        }
        case OP_UP_BLOCK: {
          Act skip = hs.back();
          hs.pop_back();
          Act prev = h;
          GnNext nx = gn_after(k, 0);
          h = res_block(prev, &skip, op.a, rowbias, &nx);
          if (prev.p != skip.p) release(prev);
          release(skip);
          break;
        }
        case OP_ATTN: {
          Act prev = h;
          h = attn_block(prev, op.a);
          if (h.p != prev.p) release(prev);
          break;
        }
        case OP_UPSAMPLE: {
          const UpW& w = net.up[op.a];
          Act o = new_act(2 * h.H, 2 * h.W, w.c);
          if ((h.H * h.W) % 128 == 0 && w.c % 16 == 0) {
            o.nchunk = 4 * h.H * h.W / 128;
            o.stats = static_cast<float*>(alloc((size_t)B * o.nchunk * 2 * w.c * 4));
          }
          if (live()) check(sd_upconv_gemm(h.p, B, h.H, h.W, h.C, w.w4, w.c, w.b, fl(), o.p, o.stats, st));
          release(h);
          h = o;
          break;
        }
      }
    }
    if (rc != SD_OK) return;
    Act a;
    if (h.norm_p == h.p && h.norm_key == net.out_g && h.norm_valid) {
      a = h;                                   // normalised in the last conv's epilogue: the raw tensor had no other reader
      h.p = nullptr; h.stats = nullptr;
    } else {
      a = gn(h, nullptr, net.out_g, net.out_be, true);
      release(h);
    }
    sd_gemm_src so[1] = {{a.p, a.C, 9}};
    if (live())
      check(sd_conv_gemm(so, 1, B, a.H, a.W, net.out_w, d.channels, net.out_b, nullptr, 0, nullptr, SD_EPI_OUT_F32 | fl(), out, d.channels,
                         nullptr, st));
  }
};

int prepare(const sd_scorenet_desc* desc, Net& net, bool need_weights) {
  if (!desc) return fail(kErrInvalidArg, "sd_scorenet: null descriptor");
  int rc = layout(*desc, net);
  if (rc != SD_OK) return rc;
  if (need_weights) {
    if (!desc->weights) return fail(kErrInvalidArg, "sd_scorenet_forward: null weights");
    if (desc->weights_bytes != net.bytes)
      return fail(kErrInvalidArg, "sd_scorenet_forward: weights_bytes does not match the configuration (see sd_scorenet_weights_bytes)");
    if ((uintptr_t)desc->weights % kAlign) return fail(kErrInvalidArg, "sd_scorenet_forward: weights must be 256-byte aligned");
  }
  return SD_OK;
}

}  // namespace
}  // namespace sdb

extern "C" {

int sd_scorenet_weights_bytes(const sd_scorenet_desc* desc, size_t* bytes_out) {
  using namespace sdb;
  if (!bytes_out) return fail(kErrInvalidArg, "sd_scorenet_weights_bytes: null output");
  sd_scorenet_desc d = desc ? *desc : sd_scorenet_desc{};
  d.weights = nullptr;
  Net net;
  int rc = prepare(desc ? &d : nullptr, net, false);
  if (rc == SD_OK) *bytes_out = net.bytes;
  return rc;
}

int sd_scorenet_workspace_bytes(const sd_scorenet_desc* desc, int B, int t_stride, size_t* bytes_out) {
  using namespace sdb;
  if (!bytes_out || B < 0) return fail(kErrInvalidArg, "sd_scorenet_workspace_bytes: bad arguments");
  sd_scorenet_desc d = desc ? *desc : sd_scorenet_desc{};
  d.weights = nullptr;
  Net net;
  int rc = prepare(desc ? &d : nullptr, net, false);
  if (rc != SD_OK) return rc;
  Runner r{d, net, B, nullptr, 0, 0, nullptr};
  r.run(nullptr, t_stride, nullptr, nullptr, nullptr);
  if (r.rc != SD_OK) return r.rc;
  *bytes_out = r.used;
  return SD_OK;
}

int sd_scorenet_forward(const sd_scorenet_desc* desc, const float* t_dev, int t_stride, const float* x_nhwc, const int* y, int B,
                        float* out_nhwc, void* workspace, size_t workspace_bytes, int precision, void* stream) {
  using namespace sdb;
  if (precision != SD_PRECISION_BF16 && precision != SD_PRECISION_FP32_FAITHFUL)
    return fail(kErrUnsupported, "sd_scorenet_forward: precision must be SD_PRECISION_BF16 or SD_PRECISION_FP32_FAITHFUL");
  if (desc && (desc->precision ? desc->precision : SD_PRECISION_BF16) != precision)
    return fail(kErrInvalidArg, "sd_scorenet_forward: `precision` does not match desc->precision (the weight blob layout depends on it)");
  if (B < 0 || (t_stride != 0 && t_stride != 1)) return fail(kErrInvalidArg, "sd_scorenet_forward: B >= 0 and t_stride in {0, 1} required");
  if (B == 0) return SD_OK;
  if (!t_dev || !x_nhwc || !out_nhwc || !workspace) return fail(kErrInvalidArg, "sd_scorenet_forward: null pointer argument");
  if ((uintptr_t)workspace % 256) return fail(kErrInvalidArg, "sd_scorenet_forward: workspace must be 256-byte aligned");
  Net net;
  int rc = prepare(desc, net, true);
  if (rc != SD_OK) return rc;
  if (desc->conditioned && !y) return fail(kErrInvalidArg, "sd_scorenet_forward: conditioned score-net needs labels");
  Runner r{*desc, net, B, static_cast<char*>(workspace), workspace_bytes, 0, (cudaStream_t)stream};
  r.run(t_dev, t_stride, x_nhwc, y, out_nhwc);
  return r.rc;
}

int sd_scorenet_forward_sched(const sd_scorenet_desc* desc, const float* sched, const int* step_counter, const float* x_nhwc,
                              const int* y, int B, float* out_nhwc, void* workspace, size_t workspace_bytes, int precision,
                              void* stream) {
  using namespace sdb;
  if (precision != SD_PRECISION_BF16 && precision != SD_PRECISION_FP32_FAITHFUL)
    return fail(kErrUnsupported, "sd_scorenet_forward_sched: precision must be SD_PRECISION_BF16 or SD_PRECISION_FP32_FAITHFUL");
  if (desc && (desc->precision ? desc->precision : SD_PRECISION_BF16) != precision)
    return fail(kErrInvalidArg, "sd_scorenet_forward_sched: `precision` does not match desc->precision (the weight blob layout depends on it)");
  if (B < 0) return fail(kErrInvalidArg, "sd_scorenet_forward_sched: B >= 0 required");
  if (B == 0) return SD_OK;
  if (!sched || !step_counter || !x_nhwc || !out_nhwc || !workspace) return fail(kErrInvalidArg, "sd_scorenet_forward_sched: null pointer argument");
  if ((uintptr_t)workspace % 256) return fail(kErrInvalidArg, "sd_scorenet_forward_sched: workspace must be 256-byte aligned");
  Net net;
  int rc = prepare(desc, net, true);
  if (rc != SD_OK) return rc;
  if (desc->conditioned && !y) return fail(kErrInvalidArg, "sd_scorenet_forward_sched: conditioned score-net needs labels");
  Runner r{*desc, net, B, static_cast<char*>(workspace), workspace_bytes, 0, (cudaStream_t)stream};
  r.sched = sched;
  r.step_counter = step_counter;
  r.run(nullptr, 0, x_nhwc, y, out_nhwc);
  return r.rc;
}

}  // extern "C"
