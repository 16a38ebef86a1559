"""B200-native 'score-net' (the reference's DDPM U-Net, cifar/models/ddpm.py:41-101).

Same registry name, config fields and call signature as the reference
(``model_fn(t, x, y)`` with t (B,1,1,1), x (B,H,W,C) NHWC fp32, y (B,) labels;
returns (B,H,W,C) fp32 = sigma_t * grad log q_t), but the forward is a fixed
sequence of hand-written sm_100a kernels reached through the C ABI:

* every 3x3 conv / NIN / Dense is the tcgen05 implicit GEMM (``ops.conv_gemm``,
  ``ops.batched_gemm``) with bf16 operands and fp32 accumulation in TMEM;
  conv bias, the time-embedding projection (layers.py:556), the NIN shortcut
  (:560-564) and the residual add (:565) are extra K segments / epilogue terms
  of the same GEMM, which also emits the next GroupNorm's channel sums; the
  3-channel input conv is a hi/lo-split im2col K-block on the same kernel;
* GroupNorm+swish (normalization.py:38-39 + layers.py:552,557) is a streaming
  apply pass over producer-emitted statistics (one register-resident kernel at
  the 8x8 / 4x4 levels) that also performs the skip concat of ddpm.py:90;
* attention (layers.py:493-511): projections folded at bind time
  (q' = NIN(h; Wq Wk^T, Wk bq) against keys h, V' = h Wv Wo), then ONE fused
  kernel softmax(q' h^T) V' + bias + x with the probabilities kept in shared
  memory (``ops.attention_core``); low resolutions pack several images per tile;
* ``jvp()`` is the forward-mode derivative (cifar/dynamics.py:84) built from the
  same GEMMs plus tangent kernels for GroupNorm+swish and the softmax.

Activations are NHWC bf16 between kernels; the output score is fp32.
There is no PyTorch / CPU fallback for the forward.
"""
import math
import os

import torch

from .. import _lib, ops
from . import utils


_FUSED_ATTENTION = os.environ.get("SDB_FUSED_ATTENTION", "1") != "0"      # tuning knob: 0 = probabilities + P V as two launches


def _conv_w(p):
    """Flax HWIO [3,3,Cin,Cout] -> bf16 [Cout, 9*Cin] (tap-major, then channel)."""
    k = p["kernel"]
    return k.permute(3, 0, 1, 2).reshape(k.shape[3], -1).contiguous()


class _Bound:
    """ScoreNet bound to one parameter tree, weights resident on the device in GEMM layout."""

    def __init__(self, model, params, device, precision="bf16"):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model = model
        self.config = model.config
        self.device = device
        self.precision = precision
        # 'fp32' = the FP32-faithful arm (SD_PRECISION_FP32_FAITHFUL): every activation and weight is a hi|lo pair of bf16
        # tensors, products are hi*hi + lo*hi + hi*lo on the tensor cores with fp32 accumulation (SD_GEMM_SPLIT3 in
        # include/superdiff_b200.h), GroupNorm / swish / softmax run in fp32 with exact transcendental forms.  ~3x the
        # tensor work of the bf16 arm for ~2^-17 relative operand error instead of 2^-9.
        self.split = precision == "fp32"
        self._prepare(params)

    # -- weight upload -------------------------------------------------------
    def _dev(self, t, dtype=torch.float32):
        return t.detach().to(device=self.device, dtype=dtype).contiguous()

    def _w(self, t):
        """GEMM weight [N, K] fp32 -> device bf16 [N, K] (bf16 arm) or [N, 2K] = [hi | lo] (FP32-faithful arm)."""
        if self.split:
            return self._dev(ops.split_pair(t.detach().to(torch.float32)), torch.bfloat16)
        return self._dev(t, torch.bfloat16)

    def _prepare(self, P):
        cfg = self.config
        m = cfg.model
        nf, ch_mult, nrb = m.nf, tuple(m.ch_mult), m.num_res_blocks
        attn_res = tuple(m.attn_resolutions)
        bf = torch.bfloat16
        self.nf = nf
        self.temb_w0 = self._dev(P["Dense_0"]["kernel"]); self.temb_b0 = self._dev(P["Dense_0"]["bias"])
        self.temb_w1 = self._dev(P["Dense_1"]["kernel"]); self.temb_b1 = self._dev(P["Dense_1"]["bias"])
        self.class_emb = self._dev(P["Embed_0"]["embedding"]) if m.conditioned else None
        self.conv_in_w = self._dev(P["Conv_0"]["kernel"]); self.conv_in_b = self._dev(P["Conv_0"]["bias"])
        # tensor-core form of the first conv: one 64-wide K-block of hi/lo-split 3x3 neighbourhoods (ops.im2col_in)
        self.conv_in_tc = cfg.data.num_channels <= 3 and cfg.data.image_size % 16 == 0
        if self.split and not self.conv_in_tc:
            raise NotImplementedError("the FP32-faithful arm needs the tensor-core first conv (<= 3 channels, image size % 16 == 0)")
        if self.split:
            self.conv_in_w64 = self._dev(ops.conv_in_weights_split(P["Conv_0"]["kernel"]), bf)
        else:
            self.conv_in_w64 = self._dev(ops.conv_in_weights(P["Conv_0"]["kernel"]), bf) if self.conv_in_tc else None

        dense_w, dense_b = [], []
        self.res = []
        self.attn = []
        self.down = []
        self.up = []
        self.plan = []     # op list mirroring the control flow of ddpm.py:70-99
        off = 0
        counters = {"res": 0, "attn": 0, "down": 0, "up": 0}

        def add_res(cin_parts, cout):
            nonlocal off
            blk = P[f"ResnetBlockDDPM_{counters['res']}"]
            cin = sum(cin_parts)
            w2 = _conv_w(blk["Conv_1"])
            b2 = blk["Conv_1"]["bias"].clone()
            nin = "NIN_0" in blk
            if nin:
                w2 = torch.cat([w2, blk["NIN_0"]["W"].T], dim=1)
                b2 = b2 + blk["NIN_0"]["b"]
            else:   # x + h (layers.py:565): the residual rides the MMA as an identity-weight 1x1 segment (exact in bf16)
                w2 = torch.cat([w2, torch.eye(cout)], dim=1)
            dense_w.append(blk["Dense_0"]["kernel"].T)                       # [cout, 4nf]
            dense_b.append(blk["Dense_0"]["bias"] + blk["Conv_0"]["bias"])   # conv1 bias folded into the row bias
            r = dict(g1=self._dev(blk["GroupNorm_0"]["scale"]), be1=self._dev(blk["GroupNorm_0"]["bias"]),
                     w1=self._w(_conv_w(blk["Conv_0"])),
                     g2=self._dev(blk["GroupNorm_1"]["scale"]), be2=self._dev(blk["GroupNorm_1"]["bias"]),
                     w2=self._w(w2), b2=self._dev(b2), nin=nin, cout=cout, off=off, cin=cin)
            off += cout
            self.res.append(r)
            counters["res"] += 1
            return len(self.res) - 1

        def add_attn(c):
            blk = P[f"AttnBlock_{counters['attn']}"]
            wq, wk, wv, wo = (blk[f"NIN_{i}"]["W"] for i in range(4))
            bq, bk, bv, bo = (blk[f"NIN_{i}"]["b"] for i in range(4))
            # Folded projections (exact algebra, evaluated in fp64 at bind time; layers.py:500-511):
            #   q k^T = (h Wq + bq)(h Wk + bk)^T = (h Wq Wk^T + (Wk bq)^T) h^T + [terms constant along a row, which the row
            #           softmax cancels]                  ->  q' = NIN(h; Wq Wk^T, Wk bq), keys = h itself (no k projection)
            #   NIN_3(P v) + x = P (h Wv Wo) + (bv Wo + bo) + x   (rows of P sum to 1)   -> V' = h (Wv Wo), no out projection
            wq64, wk64, wv64, wo64 = (w.double() for w in (wq, wk, wv, wo))
            w_q2 = (wq64 @ wk64.T).T.float()                      # GEMM weight layout [out, in]
            b_q2 = (wk64 @ bq.double()).float()
            w_vo = (wv64 @ wo64).float()                          # [C_in, C_out]
            b_vo = (bv.double() @ wo64 + bo.double()).float()
            a = dict(g=self._dev(blk["GroupNorm_0"]["scale"]), be=self._dev(blk["GroupNorm_0"]["bias"]),
                     w_qkv=self._dev(torch.cat([wq.T, wk.T, wv.T], 0), bf), b_qkv=self._dev(torch.cat([bq, bk, bv])),
                     w_q2=self._w(w_q2), b_q2=self._dev(b_q2), w_voT=self._w(w_vo.T.contiguous()), b_vo=self._dev(b_vo),
                     w_o=self._dev(torch.cat([wo.T, torch.eye(c)], dim=1), bf), b_o=self._dev(bo), c=c)   # + x (layers.py:511)
            self.attn.append(a)
            counters["attn"] += 1
            return len(self.attn) - 1

        size = cfg.data.image_size
        c = nf
        chans = [nf]
        nres = len(ch_mult)
        for lvl in range(nres):
            for _ in range(nrb):
                ri = add_res([c], nf * ch_mult[lvl])
                c = nf * ch_mult[lvl]
                ai = add_attn(c) if size in attn_res else None
                self.plan.append(("down_block", ri, ai))
                chans.append(c)
            if lvl != nres - 1:
                blk = P[f"Downsample_{counters['down']}"]["Conv_0"]
                self.down.append(dict(w=self._w(_conv_w(blk)), b=self._dev(blk["bias"])))
                self.plan.append(("downsample", len(self.down) - 1))
                counters["down"] += 1
                size //= 2
                chans.append(c)
        r0 = add_res([c], c); a0 = add_attn(c); r1 = add_res([c], c)
        self.plan.append(("mid", r0, a0, r1))
        for lvl in reversed(range(nres)):
            for _ in range(nrb + 1):
                ri = add_res([c, chans.pop()], nf * ch_mult[lvl])
                c = nf * ch_mult[lvl]
                self.plan.append(("up_block", ri))
            if size in attn_res:
                self.plan.append(("attn", add_attn(c)))
            if lvl != 0:
                blk = P[f"Upsample_{counters['up']}"]["Conv_0"]
                self.up.append(dict(w4=self._w(ops.upconv_weights(blk["kernel"])), b=self._dev(blk["bias"])))
                self.plan.append(("upsample", len(self.up) - 1))
                counters["up"] += 1
                size *= 2
        assert not chans
        self.out_g = self._dev(P["GroupNorm_0"]["scale"]); self.out_be = self._dev(P["GroupNorm_0"]["bias"])
        wout = _conv_w(P["Conv_1"])                       # [C_img, 9*nf]
        self.n_img = wout.shape[0]
        pad = (-wout.shape[0]) % 16
        self.out_w = self._w(torch.cat([wout, torch.zeros(pad, wout.shape[1])], 0))
        self.out_b = self._dev(P["Conv_1"]["bias"])
        self.dense_w = self._w(torch.cat(dense_w, 0))         # [sum cout, 4nf]
        self.dense_b = self._dev(torch.cat(dense_b, 0))
        self.dense_n = off

    # -- forward ---------------------------------------------------------------
    @staticmethod
    def _pre_normalised(x, gamma):
        """The tensor a producing GEMM already normalised for the GroupNorm whose scale is `gamma` (fused epilogue), or None."""
        if getattr(x, "gn_norm", None) is not None and getattr(x, "gn_norm_key", None) == id(gamma):
            return x.gn_norm
        return None

    def _conv_then_gn(self, srcs, weight, next_gn, **kw):
        """conv_gemm whose output's first consumer is a GroupNorm (`next_gn` = (gamma, beta, swish, keep_raw) or None): tags the
        result so that the consumer finds the pre-normalised tensor (`_pre_normalised`)."""
        if next_gn is None:
            return ops.conv_gemm(srcs, weight, want_stats=True, split=self.split, **kw)
        out = ops.conv_gemm(srcs, weight, want_stats=True, split=self.split, gn=next_gn, **kw)
        if next_gn[3]:                       # raw tensor returned, normalised one attached (None when the launch did not fuse)
            out.gn_norm_key = id(next_gn[0])
        elif out.gn_fused:                   # the tensor itself is normalised (its raw form has no other reader)
            out.gn_norm, out.gn_norm_key = out, id(next_gn[0])
        return out

    def _res(self, srcs, i, rowbias, next_gn=None):
        r = self.res[i]
        x0 = srcs[0]
        x1 = srcs[1] if len(srcs) > 1 else None
        sp = self.split
        a1 = self._pre_normalised(x0, r["g1"]) if x1 is None else None
        if a1 is None:
            a1 = ops.groupnorm_swish(x0, r["g1"], r["be1"], x1=x1, split=sp)
        # conv1 + temb bias, then act(normalize(.)) (layers.py:553-557): the GroupNorm runs inside the GEMM epilogue where the tile
        # shape lets one CTA / cluster see the whole image (sd_conv_gemm_gn), as its own pass otherwise
        h1 = ops.conv_gemm([(a1, 9)], r["w1"], rowbias=rowbias[:, r["off"]:r["off"] + r["cout"]], want_stats=True, split=sp,
                           gn=(r["g2"], r["be2"]))
        a2 = h1 if h1.gn_fused else ops.groupnorm_swish(h1, r["g2"], r["be2"], split=sp)
        # NIN shortcut (C_in != C_out) or identity residual: both are extra 1-tap K segments of the same GEMM; the GroupNorm of the
        # layer that follows (next block, attention block, output head) runs in this GEMM's epilogue where the tile shape allows
        return self._conv_then_gn([(a2, 9)] + [(s, 1) for s in srcs], r["w2"], next_gn, bias=r["b2"])

    def _attn_split(self, x, i):
        """AttnBlock in the FP32-faithful arm: same folded algebra, every product as a 3 x bf16 split GEMM, the softmax in fp32
        on the fp32 scores (five launches; the fused bf16 core keeps its probabilities in bf16 and is not used here)."""
        a = self.attn[i]
        B, H, W, C2 = x.shape
        C, S = C2 // 2, H * W
        g = max(1, 128 // S)
        if S < 16 or (g * S) % 16:
            raise NotImplementedError("the FP32-faithful attention needs at least 16 pixels per image")
        if B % g:      # pad the last packed tile with zero images (block-diagonal softmax keeps images independent)
            z = x.new_zeros((g - B % g,) + tuple(x.shape[1:]))
            return self._attn_split(torch.cat([x, z]), i)[:B]
        h = ops.groupnorm_swish(x, a["g"], a["be"], swish=False, split=True)
        nb, Sp = B // g, g * S
        hb = h.view(nb, Sp, C2)
        q2 = ops.conv_gemm([(h, 1)], a["w_q2"], bias=a["b_q2"], split=True).view(nb, Sp, C2)
        vt = ops.batched_gemm(a["w_voT"], hb, split=True)                                # [nb, C, 2 Sp]
        sc = ops.batched_gemm(q2, hb, out_f32=True, split=True)                          # fp32 scores [nb, Sp, Sp]
        p = ops.softmax_rows_split(sc, C ** -0.5, block=S)                               # [nb, Sp, 2 Sp], block diagonal
        out = ops.batched_gemm(p, vt, bias=a["b_vo"], residual=x.view(nb, Sp, C2), want_stats=(g == 1), split=True)
        res = out.view(B, H, W, C2)
        if hasattr(out, "gn_stats"):
            res.gn_stats = out.gn_stats
        return res

    def _attn(self, x, i):
        if self.split:
            return self._attn_split(x, i)
        a = self.attn[i]
        B, H, W, C = x.shape
        S = H * W
        h = self._pre_normalised(x, a["g"])
        if h is None:
            h = ops.groupnorm_swish(x, a["g"], a["be"], swish=False)
        g = max(1, 128 // S)              # low resolution: pack g images per 128-row tensor-core tile
        if S < 16 or B % g or (g * S) % 16:
            qkv = ops.conv_gemm([(h, 1)], a["w_qkv"], bias=a["b_qkv"])
            o = ops.attention_small(qkv.view(B, S, 3 * C), C)
            return ops.conv_gemm([(o.view(B, H, W, C), 1), (x, 1)], a["w_o"], bias=a["b_o"], want_stats=True)
        nb, Sp = B // g, g * S
        hb = h.view(nb, Sp, C)
        q2 = ops.conv_gemm([(h, 1)], a["w_q2"], bias=a["b_q2"]).view(nb, Sp, C)      # q' = h Wq Wk^T + Wk bq
        vt = ops.batched_gemm(a["w_voT"], hb)                                        # [nb, C, Sp] = (h Wv Wo)^T
        if Sp in (128, 256) and C % 64 == 0 and C <= 256 and _FUSED_ATTENTION:
            # softmax(q' h^T) V' + bias + x in one kernel: the probabilities stay in shared memory (csrc/attn_core.cu)
            out = ops.attention_core(q2, hb, vt, C ** -0.5, block=S, bias=a["b_vo"], residual=x.view(nb, Sp, C),
                                     want_stats=(g == 1), C=C)
        else:
            p = ops.attention_probs(q2, hb, C ** -0.5, block=S, C=C)                 # [nb, Sp, Sp], block diagonal; keys = h
            out = ops.batched_gemm(p, vt, bias=a["b_vo"], residual=x.view(nb, Sp, C), want_stats=(g == 1))
        res = out.view(B, H, W, C)
        if hasattr(out, "gn_stats"):
            res.gn_stats = out.gn_stats
        return res

    def __call__(self, t, x, y=None, *, sched=None, step_counter=None, out=None):
        """t: (B,1,1,1)/(B,)/scalar tensor or float; x: (B,H,W,C) fp32 NHWC on the device."""
        _lib.require_device()
        B = x.shape[0]
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise ValueError("x must be a contiguous float32 CUDA tensor (NHWC)")
        rowbias = self._rowbias(t, x, y, sched, step_counter)
        plan = self.plan

        def gn_after(k, stage=0):
            """(gamma, beta, swish, keep_raw) of the GroupNorm that first consumes the tensor produced by plan[k] (stage: which of a
            'mid' op's two ResBlocks; k = -1: the first conv), or None when that consumer is not a plain GroupNorm of this tensor
            alone (skip concatenation, down / upsample, an attention block's output)."""
            if k >= 0:
                op = plan[k]
                if op[0] == "down_block" and op[2] is not None:
                    a = self.attn[op[2]]
                    return (a["g"], a["be"], False, True)
                if op[0] == "mid":
                    if stage == 0:
                        a = self.attn[op[2]]
                        return (a["g"], a["be"], False, True)
                    return None
                if op[0] == "up_block":
                    if k + 1 == len(plan):
                        return (self.out_g, self.out_be, True, False)          # output head: the raw tensor has no other reader
                    if plan[k + 1][0] == "attn":
                        a = self.attn[plan[k + 1][1]]
                        return (a["g"], a["be"], False, True)
                    return None
            nxt = plan[k + 1] if k + 1 < len(plan) else None
            if nxt is not None and nxt[0] in ("down_block", "mid") and (k < 0 or plan[k][0] == "down_block"):
                r = self.res[nxt[1]]
                return (r["g1"], r["be1"], True, True)
            return None

        h = self._conv_in(x, self.conv_in_b, True, gn_after(-1))
        hs = [h]
        for k, op in enumerate(plan):
            kind = op[0]
            if kind == "down_block":
                h = self._res([hs[-1]], op[1], rowbias, gn_after(k))
                if op[2] is not None:
                    h = self._attn(h, op[2])
                hs.append(h)
            elif kind == "downsample":
                d = self.down[op[1]]
                h = ops.conv_gemm_s2(hs[-1], d["w"], bias=d["b"], want_stats=True, split=self.split)
                hs.append(h)
            elif kind == "mid":
                h = self._res([hs[-1]], op[1], rowbias, gn_after(k, 0))
                h = self._attn(h, op[2])
                h = self._res([h], op[3], rowbias, gn_after(k, 1))
            elif kind == "up_block":
                h = self._res([h, hs.pop()], op[1], rowbias, gn_after(k))
            elif kind == "attn":
                h = self._attn(h, op[1])
            elif kind == "upsample":
                u = self.up[op[1]]
                h = ops.upconv_gemm(h, u["w4"], bias=u["b"], want_stats=True, split=self.split)
        assert not hs
        a = self._pre_normalised(h, self.out_g)
        if a is None:
            a = ops.groupnorm_swish(h, self.out_g, self.out_be, split=self.split)
        return ops.conv_gemm([(a, 9)], self.out_w, bias=self.out_b, out_f32=True, n_out=self.n_img, out=out, split=self.split)

    def _conv_in(self, x, bias, want_stats, next_gn=None):
        """conv3x3(x, nf) (ddpm.py:71).  Tensor-core form: 166 -> ~55 us at batch 512 and the GEMM epilogue emits the channel
        sums the first GroupNorm needs; the CUDA-core kernel remains for inputs the gather does not cover."""
        if self.conv_in_tc:
            cols = ops.im2col_in(x, split=self.split)
            if next_gn is not None and want_stats:
                return self._conv_then_gn([(cols, 1)], self.conv_in_w64, next_gn, bias=bias)
            return ops.conv_gemm([(cols, 1)], self.conv_in_w64, bias=bias, want_stats=want_stats, split=self.split)
        return ops.conv_in(x, self.conv_in_w, bias)

    def _rowbias(self, t, x, y, sched, step_counter):
        """Time embedding (ddpm.py:64-68) -> the per-sample bias of every ResBlock's Dense(temb) (layers.py:556): [B, sum cout]."""
        B = x.shape[0]
        if sched is not None and self.class_emb is None and step_counter is not None:
            # sampler mode, unconditioned model: the biases depend only on t, so they are tabulated once for the schedule's
            # n_steps times (one time-embedding launch + one GEMM at set-up) and each timestep gathers its row -- instead of a
            # single-CTA MLP (58 us), an activation pass and a [B, 4 nf] x [sum cout] GEMM on the critical path of every forward
            # keyed on the schedule tensor itself (kept alive by the cache entry, so its address cannot be recycled) and
            # its version counter (in-place edits of the table re-tabulate); old tables stay alive because a captured
            # CUDA graph of another sampler may still read them
            cache = self.__dict__.setdefault("_rb_cache", {})
            key = (id(sched), sched._version)
            ent = cache.get(key)
            if ent is None or ent[0] is not sched:
                ts = sched[:, 2].contiguous()                     # sigma_t = t (cifar/dynamics.py:105)
                act = ops.time_embedding(ts.shape[0], self.nf, self.temb_w0, self.temb_b0, self.temb_w1, self.temb_b1,
                                         t=ts, t_stride=1, split=self.split)
                table = ops.batched_gemm(act, self.dense_w, bias=self.dense_b, out_f32=True, split=self.split)[0].contiguous()
                row = torch.empty(1, table.shape[1], device=x.device, dtype=torch.float32)
                ent = cache[key] = (sched, table, row)
            self._rb_table, self._rb_row = ent[1], ent[2]
            ops.gather_row(self._rb_table, step_counter, out=self._rb_row)
            return self._rb_row.expand(B, -1)                     # row stride 0: every sample reads the same biases
        labels = None
        if self.class_emb is not None:
            if y is None:
                raise ValueError("conditioned score-net needs labels")
            labels = y.to(device=x.device, dtype=torch.int32).contiguous()
        if sched is not None:
            act = ops.time_embedding(B, self.nf, self.temb_w0, self.temb_b0, self.temb_w1, self.temb_b1,
                                     sched=sched, step_counter=step_counter, class_emb=self.class_emb, labels=labels,
                                     split=self.split)
        else:
            if not torch.is_tensor(t):
                t = torch.full((1,), float(t), device=x.device, dtype=torch.float32)
            t = t.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
            stride = 0 if t.numel() == 1 else 1
            if stride and t.numel() != B:
                raise ValueError("t must have one entry per sample")
            act = ops.time_embedding(B, self.nf, self.temb_w0, self.temb_b0, self.temb_w1, self.temb_b1,
                                     t=t, t_stride=stride, class_emb=self.class_emb, labels=labels, split=self.split)
        return ops.batched_gemm(act, self.dense_w, bias=self.dense_b, out_f32=True, split=self.split)[0]    # [B, sum cout]

    # -- forward-mode derivative (Hutchinson probes of the deterministic sampler, cifar/dynamics.py:84) -------------
    def _res_jvp(self, srcs, dsrcs, i, rowbias):
        r = self.res[i]
        two = len(srcs) > 1
        a1, da1 = ops.groupnorm_swish_jvp(srcs[0], dsrcs[0], r["g1"], r["be1"], x1=srcs[1] if two else None,
                                          dx1=dsrcs[1] if two else None)
        h1 = ops.conv_gemm([(a1, 9)], r["w1"], rowbias=rowbias[:, r["off"]:r["off"] + r["cout"]])
        dh1 = ops.conv_gemm([(da1, 9)], r["w1"])                     # the time-embedding bias does not depend on x
        a2, da2 = ops.groupnorm_swish_jvp(h1, dh1, r["g2"], r["be2"])
        out = ops.conv_gemm([(a2, 9)] + [(s, 1) for s in srcs], r["w2"], bias=r["b2"])
        dout = ops.conv_gemm([(da2, 9)] + [(d, 1) for d in dsrcs], r["w2"])
        return out, dout

    def _attn_jvp(self, x, dx, i):
        a = self.attn[i]
        B, H, W, C = x.shape
        S = H * W
        g = max(1, 128 // S)
        if S < 16 or (g * S) % 16:
            raise NotImplementedError("score-net JVP needs at least 16 pixels per image at the attention resolutions")
        if B % g:
            # low-resolution blocks pack g images per 128-row tile: pad the batch with zero images (the block-diagonal
            # softmax keeps images independent, zeros keep the padded keys / values finite) and drop them again.
            # The reference's eval batch (100; 12 per GPU on 8 GPUs) is not a multiple of 8.
            pad = g - B % g
            z = x.new_zeros((pad,) + tuple(x.shape[1:]))
            o, do = self._attn_jvp(torch.cat([x, z]), torch.cat([dx, z]), i)
            return o[:B], do[:B]
        h, dh = ops.groupnorm_swish_jvp(x, dx, a["g"], a["be"], swish=False)
        nb, Sp = B // g, g * S
        scale = C ** -0.5
        hb, dhb = h.view(nb, Sp, C), dh.view(nb, Sp, C)
        q2 = ops.conv_gemm([(h, 1)], a["w_q2"], bias=a["b_q2"]).view(nb, Sp, C)      # folded projections, see add_attn
        dq2 = ops.conv_gemm([(dh, 1)], a["w_q2"]).view(nb, Sp, C)
        vt = ops.batched_gemm(a["w_voT"], hb)
        dvt = ops.batched_gemm(a["w_voT"], dhb)
        p = ops.attention_probs(q2, hb, scale, block=S, C=C)
        ds1 = ops.batched_gemm(dq2, hb, out_f32=True, K=C)                           # dq' h^T
        ds2 = ops.batched_gemm(q2, dhb, out_f32=True, K=C)                           # q' dh^T
        dp = ops.softmax_jvp(p, ds1, ds2, scale)       # off-block entries of a packed tile: p = 0 -> dp = 0
        out = ops.batched_gemm(p, vt, bias=a["b_vo"], residual=x.view(nb, Sp, C))
        do = ops.batched_gemm(dp, vt, residual=dx.view(nb, Sp, C))                   # dP V' + dx ...
        dout = ops.batched_gemm(p, dvt, residual=do)                                 # ... + P dV'
        return out.view(B, H, W, C), dout.view(B, H, W, C)

    def jvp(self, t, x, y, v, *, sched=None, step_counter=None, out=None, jvp_out=None):
        """(score(x), d/dh score(x + h v)|_0): what jax.jvp(sdlogdx_fn, (x,), (eps,)) returns at cifar/dynamics.py:84.
        Linear layers reuse the tcgen05 GEMMs on the tangent stream (no bias); GroupNorm+swish and the attention softmax
        have tangent kernels (csrc/scorenet_jvp.cu).  Tangents travel in bf16 like the activations."""
        _lib.require_device()
        if self.split:
            raise NotImplementedError("the tangent kernels exist for the bf16 arm only (bind with precision='bf16')")
        for name, ten in (("x", x), ("v", v)):
            if not (ten.is_cuda and ten.dtype == torch.float32 and ten.is_contiguous()):
                raise ValueError(f"{name} must be a contiguous float32 CUDA tensor (NHWC)")
        if v.shape != x.shape:
            raise ValueError("tangent must have the shape of x")
        rowbias = self._rowbias(t, x, y, sched, step_counter)
        h, dh = self._conv_in(x, self.conv_in_b, False), self._conv_in(v, None, False)
        hs = [(h, dh)]
        for op in self.plan:
            kind = op[0]
            if kind == "down_block":
                h, dh = self._res_jvp([hs[-1][0]], [hs[-1][1]], op[1], rowbias)
                if op[2] is not None:
                    h, dh = self._attn_jvp(h, dh, op[2])
                hs.append((h, dh))
            elif kind == "downsample":
                d = self.down[op[1]]
                h, dh = ops.conv_gemm_s2(hs[-1][0], d["w"], bias=d["b"]), ops.conv_gemm_s2(hs[-1][1], d["w"])
                hs.append((h, dh))
            elif kind == "mid":
                h, dh = self._res_jvp([hs[-1][0]], [hs[-1][1]], op[1], rowbias)
                h, dh = self._attn_jvp(h, dh, op[2])
                h, dh = self._res_jvp([h], [dh], op[3], rowbias)
            elif kind == "up_block":
                sk, dsk = hs.pop()
                h, dh = self._res_jvp([h, sk], [dh, dsk], op[1], rowbias)
            elif kind == "attn":
                h, dh = self._attn_jvp(h, dh, op[1])
            elif kind == "upsample":
                u = self.up[op[1]]
                h, dh = ops.upconv_gemm(h, u["w4"], bias=u["b"]), ops.upconv_gemm(dh, u["w4"])
        assert not hs
        a, da = ops.groupnorm_swish_jvp(h, dh, self.out_g, self.out_be)
        score = ops.conv_gemm([(a, 9)], self.out_w, bias=self.out_b, out_f32=True, n_out=self.n_img, out=out)
        tangent = ops.conv_gemm([(da, 9)], self.out_w, out_f32=True, n_out=self.n_img, out=jvp_out)
        return score, tangent


@utils.register_model(name="score-net")
class ScoreNet:
    """cifar/models/ddpm.py:41-101.  ``ScoreNet(config).bind(params)`` gives the callable."""

    def __init__(self, config):
        self.config = config
        m = config.model
        if m.normalization != "GroupNorm" or m.nonlinearity.lower() != "swish" or not m.resamp_with_conv:
            raise NotImplementedError("the B200 score-net implements the reference's CIFAR config family: "
                                      "GroupNorm + swish + conv resampling (cifar/configs/sm/cifar/vpsde*.py)")
        if m.nf % 64 or config.data.num_channels > 4:
            raise NotImplementedError("nf must be a multiple of 64 and the image must have <= 4 channels")

    def bind(self, params, device=None, precision=None):
        """precision: 'bf16' (default; bf16 operands / activations, fp32 accumulation) or 'fp32' (the FP32-faithful arm);
        None reads config.model.precision when the config carries it."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if precision is None:
            precision = getattr(self.config.model, "precision", "bf16") if hasattr(self.config, "model") else "bf16"
        return _Bound(self, params, device, precision)

    def apply(self, variables, t, x, y, train=False, mutable=False, rngs=None):
        """Flax-style entry used by the reference's get_model_fn (models/utils.py:91-95)."""
        if train:
            raise NotImplementedError("train=True (dropout) is outside the sampling path")
        return self.bound_for(variables["params"], x.device)(t, x, y)

    def bound_for(self, params, device=None, precision=None):
        """The bound net of a parameter tree, uploaded once per (tree object, device).  The entry holds the tree, so an
        id() can never be recycled for another tree while it is cached; a handful of trees at most (the M models)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cache = self.__dict__.setdefault("_bound_cache", {})
        if precision is None:
            precision = getattr(self.config.model, "precision", "bf16")
        key = (id(params), str(device), precision)
        ent = cache.get(key)
        if ent is None or ent[0] is not params:
            if len(cache) >= 16:
                cache.clear()
            ent = cache[key] = (params, self.bind(params, device, precision))
        return ent[1]
