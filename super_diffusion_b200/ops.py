"""Tensor-level wrappers over the C ABI (one Python call = one kernel launch)."""
import ctypes

import torch

from . import _lib

MODE_OR, MODE_AND, MODE_AVG, MODE_FIXED = 0, 1, 2, 3
DLOGQ_CIFAR_MAXSUB, DLOGQ_ITO, DLOGQ_NONE = 0, 1, 2
MODES = {"or": MODE_OR, "and": MODE_AND, "avg": MODE_AVG, "fixed": MODE_FIXED}
DLOGQ_MODES = {"cifar": DLOGQ_CIFAR_MAXSUB, "ito": DLOGQ_ITO, "none": DLOGQ_NONE}

_launches = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t, name):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor")
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def step_vpsde(x, noise, scores, logq, a, b, sigma, dt, mode, dlogq_mode, temperature=1.0,
               logp_bias=None, ito_scale=0.0, x_out=None, weights=None, sched=None, step_counter=None,
               launch_shape=None):
    """Fused VP-SDE SuperDiff step (sd_step_vpsde).  x, noise: (B, ...); scores:
    list of M tensors shaped like x (or one stacked (M, B, ...) tensor); logq,
    weights: (B, M), updated in place.  Returns (x_out, logq, weights)."""
    lib = _lib.load()
    if isinstance(scores, torch.Tensor):
        scores = list(scores.unbind(0))
    M = len(scores)
    B = x.shape[0]
    D = x[0].numel() if B > 0 else int(torch.tensor(x.shape[1:]).prod())
    _f32c(x, "x"); _f32c(noise, "noise"); _f32c(logq, "logq")
    for i, s in enumerate(scores):
        _f32c(s, f"scores[{i}]")
        if s.shape != x.shape:
            raise ValueError("score shape mismatch")
    if noise.shape != x.shape or logq.shape != (B, M):
        raise ValueError("noise must match x and logq must be (B, M)")
    if x_out is None:
        x_out = torch.empty_like(x)
    if weights is None:
        weights = torch.empty(B, M, device=x.device, dtype=torch.float32)
    _f32c(x_out, "x_out"); _f32c(weights, "weights")
    if isinstance(mode, str):
        mode = MODES[mode]
    if isinstance(dlogq_mode, str):
        dlogq_mode = DLOGQ_MODES[dlogq_mode]
    sp = (ctypes.c_void_p * M)(*[s.data_ptr() for s in scores])
    common = (_ptr(x), _ptr(noise), sp, M, B, D, float(a), float(b), float(sigma), float(dt),
              _ptr(sched), _ptr(step_counter), int(mode), int(dlogq_mode), float(temperature),
              _ptr(logp_bias), float(ito_scale), _ptr(logq), _ptr(x_out), _ptr(weights), _stream())
    if launch_shape is None:
        rc = lib.sd_step_vpsde(*common)
    else:
        rc = lib.sd_step_vpsde_ex(*common, *[int(v) for v in launch_shape])
    _lib.check(rc, "sd_step_vpsde")
    if B > 0:
        _count()
    return x_out, logq, weights


def step_vpsde_ode(x, scores, logq, a, b, sigma_eps, dt, mode, dlogq_mode, temperature=1.0, logp_bias=None,
                   dlogq_add=None, x_out=None, weights=None, sched=None, step_counter=None):
    """Deterministic SuperDiff / averaged / single-model ODE step (sd_step_vpsde_ode): no noise tensor, drift
    -dt*(a x - b s_mix), optional per-(sample, model) additive term ``dlogq_add`` (B, M) = dt * divergence estimate."""
    lib = _lib.load()
    if isinstance(scores, torch.Tensor):
        scores = list(scores.unbind(0))
    M, B = len(scores), x.shape[0]
    D = x[0].numel()
    _f32c(x, "x")
    for i, s in enumerate(scores):
        _f32c(s, f"scores[{i}]")
        if s.shape != x.shape:
            raise ValueError("score shape mismatch")
    if logq is not None:
        _f32c(logq, "logq")
        if logq.shape != (B, M):
            raise ValueError("logq must be (B, M)")
    if dlogq_add is not None and (_f32c(dlogq_add, "dlogq_add").shape != (B, M)):
        raise ValueError("dlogq_add must be (B, M)")
    if x_out is None:
        x_out = torch.empty_like(x)
    if weights is None:
        weights = torch.empty(B, M, device=x.device, dtype=torch.float32)
    mode = MODES[mode] if isinstance(mode, str) else mode
    dlogq_mode = DLOGQ_MODES[dlogq_mode] if isinstance(dlogq_mode, str) else dlogq_mode
    sp = (ctypes.c_void_p * M)(*[s.data_ptr() for s in scores])
    rc = lib.sd_step_vpsde_ode(_ptr(x), sp, M, B, D, float(a), float(b), float(sigma_eps), float(dt), _ptr(sched),
                               _ptr(step_counter), int(mode), int(dlogq_mode), float(temperature), _ptr(logp_bias),
                               _ptr(dlogq_add), _ptr(logq), _ptr(x_out), _ptr(weights), _stream())
    _lib.check(rc, "sd_step_vpsde_ode")
    if B > 0:
        _count()
    return x_out, logq, weights


def rowdot(a, b, scale=1.0, out=None, column=0):
    """out[:, column] = scale * sum over all but the first axis of a*b (sd_rowdot); out: (B, M) fp32 or None -> (B,)."""
    lib = _lib.load()
    _f32c(a, "a"); _f32c(b, "b")
    B, D = a.shape[0], a[0].numel()
    if out is None:
        out = torch.empty(B, device=a.device, dtype=torch.float32)
        ptr, stride = out.data_ptr(), 1
    else:
        _f32c(out, "out")
        ptr, stride = out.data_ptr() + 4 * column, out.shape[1]
    _lib.check(lib.sd_rowdot(_ptr(a), _ptr(b), B, D, float(scale), ctypes.c_void_p(ptr), stride, _stream()), "sd_rowdot")
    if B > 0:
        _count()
    return out


def step_edm_cfg(latents, z, v_obj, v_bg, v_unc, ll, sigma, dsigma, mode, guidance=7.5, lift_term=0.0,
                 temperature=1.0, logp=0.0, kappa_fixed=0.5, latents_out=None, kappa_out=None):
    """Fused EDM/CFG SuperDiff step on SD latents (sd_step_edm_cfg).  ll: (B, 2)
    updated in place.  Returns (latents_out, ll, kappa)."""
    lib = _lib.load()
    B = latents.shape[0]
    D = latents[0].numel()
    for n, t in (("latents", latents), ("z", z), ("v_obj", v_obj), ("v_bg", v_bg), ("v_unc", v_unc), ("ll", ll)):
        _f32c(t, n)
    if latents_out is None:
        latents_out = torch.empty_like(latents)
    if kappa_out is None:
        kappa_out = torch.empty(B, device=latents.device, dtype=torch.float32)
    if isinstance(mode, str):
        mode = MODES[mode]
    rc = lib.sd_step_edm_cfg(_ptr(latents), _ptr(z), _ptr(v_obj), _ptr(v_bg), _ptr(v_unc), B, D,
                             float(sigma), float(dsigma), float(guidance), float(lift_term), int(mode),
                             float(temperature), float(logp), float(kappa_fixed),
                             _ptr(ll), _ptr(latents_out), _ptr(kappa_out), _stream())
    _lib.check(rc, "sd_step_edm_cfg")
    if B > 0:
        _count()
    return latents_out, ll, kappa_out


def step_edm_ode(latents, v_obj, v_bg, v_unc, dlog, ll, sigma, dsigma, guidance=7.5, lift_term=0.0, latents_out=None,
                 kappa_out=None):
    """Deterministic AND step on SD latents (sd_step_edm_ode).  dlog: (B, 2) Hutchinson divergences; ll: (B, 2) in place."""
    lib = _lib.load()
    B, D = latents.shape[0], latents[0].numel()
    for n, t in (("latents", latents), ("v_obj", v_obj), ("v_bg", v_bg), ("v_unc", v_unc), ("dlog", dlog), ("ll", ll)):
        _f32c(t, n)
    if dlog.shape != (B, 2) or ll.shape != (B, 2):
        raise ValueError("dlog and ll must be (B, 2)")
    if latents_out is None:
        latents_out = torch.empty_like(latents)
    if kappa_out is None:
        kappa_out = torch.empty(B, device=latents.device, dtype=torch.float32)
    rc = lib.sd_step_edm_ode(_ptr(latents), _ptr(v_obj), _ptr(v_bg), _ptr(v_unc), _ptr(dlog), B, D, float(sigma), float(dsigma),
                             float(guidance), float(lift_term), _ptr(ll), _ptr(latents_out), _ptr(kappa_out), _stream())
    _lib.check(rc, "sd_step_edm_ode")
    if B > 0:
        _count()
    return latents_out, ll, kappa_out


def counter_add(counter, delta=1, rows=None):
    """*counter += delta on the device; with ``rows`` the result saturates at rows - 1 (sd_counter_add_sat) so a schedule
    table of that many rows is never indexed out of bounds."""
    if rows is None:
        _lib.check(_lib.load().sd_counter_add(_ptr(counter), int(delta), _stream()), "sd_counter_add")
    else:
        _lib.check(_lib.load().sd_counter_add_sat(_ptr(counter), int(delta), int(rows), _stream()), "sd_counter_add_sat")
    _count()


# ---------------------------------------------------------------------------
# score-net ops (NHWC bf16 activations)
# ---------------------------------------------------------------------------
EPI_SWISH, EPI_OUT_F32, GEMM_SPLIT3 = 1, 2, 8


def split_pair(x):
    """fp32 [..., C] -> bf16 [..., 2C] = [hi | lo], hi = bf16(x), lo = bf16(x - hi): the operand format of the FP32-faithful
    arm (SD_GEMM_SPLIT3 in include/superdiff_b200.h).  Host / torch helper for weights and test inputs."""
    hi = x.to(torch.bfloat16)
    lo = (x.to(torch.float32) - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def merge_pair(x):
    """Inverse of split_pair up to ~2^-17 relative: bf16 [..., 2C] -> fp32 [..., C] = hi + lo."""
    C = x.shape[-1] // 2
    return x[..., :C].to(torch.float32) + x[..., C:].to(torch.float32)


def _bf16c(t, name):
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous bfloat16 CUDA tensor")
    return t


def conv_gemm(srcs, weight, bias=None, rowbias=None, residual=None, swish=False, out_f32=False, out=None,
              n_out=None, want_stats=False, split=False, gn=None):
    """Implicit GEMM over NHWC bf16 sources (sd_conv_gemm).  srcs: list of
    (tensor [B,H,W,C], taps) with taps in {1, 9}; weight: bf16 [N, K].  want_stats: also emit per-128-pixel-tile
    channel sums of the output (attached as ``out.gn_stats = (tensor [B, HW/128, 2, N], HW/128)``) so a following
    groupnorm_swish skips its statistics pass; silently ignored where the layout does not allow it.
    split: 3 x bf16 split precision (SD_GEMM_SPLIT3) -- sources / residual / out are hi|lo pairs [B,H,W,2C], weight [N, 2K].
    gn: (gamma, beta) of a GroupNorm + swish that is the ONLY consumer of this output (sd_conv_gemm_gn): where the tile shape
    allows it the normalised, activated tensor is written directly and ``out.gn_fused`` is True; otherwise the raw output (with
    ``gn_stats``) comes back with ``out.gn_fused`` False and the caller applies groupnorm_swish."""
    lib = _lib.load()
    x0 = srcs[0][0]
    B, H, W = x0.shape[0], x0.shape[1], x0.shape[2]
    arr = (_lib.GemmSrc * len(srcs))()
    K = 0
    sm = 2 if split else 1
    for i, (t, taps) in enumerate(srcs):
        _bf16c(t, f"srcs[{i}]")
        if t.shape[:3] != x0.shape[:3]:
            raise ValueError("all sources must share (B, H, W)")
        arr[i].ptr = t.data_ptr()
        arr[i].C = t.shape[3] // sm
        arr[i].taps = taps
        arr[i].ld = t.shape[3]
        K += taps * t.shape[3]
    _bf16c(weight, "weight")
    N = weight.shape[0] if n_out is None else n_out
    if weight.shape[1] != K:
        raise ValueError(f"weight K {weight.shape[1]} != {K}")
    if out is None:
        out = torch.empty(B, H, W, N if out_f32 else sm * N, device=x0.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    flags = (EPI_SWISH if swish else 0) | (EPI_OUT_F32 if out_f32 else 0) | (GEMM_SPLIT3 if split else 0)
    rb_ld = rowbias.stride(0) if rowbias is not None else 0
    if residual is not None:
        _bf16c(residual, "residual")
    stats = None
    if want_stats and (H * W) % 128 == 0 and N % 16 == 0 and not out_f32 and B > 0:
        stats = torch.empty(B, (H * W) // 128, 2, N, device=x0.device, dtype=torch.float32)
    if gn is not None:
        # gn = (gamma, beta[, swish[, keep_raw]]): keep_raw asks for the raw tensor as a second output (it stays on the residual
        # stream); the return value is then the RAW tensor with `.gn_norm` = the normalised one (None when the launch did not fuse)
        if residual is not None or out_f32 or swish:
            raise ValueError("gn fusion: no residual / fp32 output / extra swish")
        gn_swish = bool(gn[2]) if len(gn) > 2 else True
        keep_raw = bool(gn[3]) if len(gn) > 3 else False
        raw = torch.empty_like(out) if keep_raw else None
        fused = ctypes.c_int(0)
        rc = lib.sd_conv_gemm_gn(arr, len(srcs), B, H, W, _ptr(weight), N, _ptr(bias), _ptr(rowbias), rb_ld, flags, _ptr(out),
                                 out.shape[-1], _ptr(stats), _ptr(_f32c(gn[0], "gamma")), _ptr(_f32c(gn[1], "beta")), 1e-6, int(gn_swish),
                                 _ptr(raw), ctypes.byref(fused), _stream())
        _lib.check(rc, "sd_conv_gemm_gn")
        if keep_raw:
            norm = out if fused.value else None
            out = raw
            out.gn_norm = norm
            out.gn_fused = False           # `out` is the raw tensor: its gn_stats (below) are valid either way
        else:
            out.gn_fused = bool(fused.value)
    else:
        rc = lib.sd_conv_gemm(arr, len(srcs), B, H, W, _ptr(weight), N, _ptr(bias), _ptr(rowbias), rb_ld,
                              _ptr(residual), flags, _ptr(out), out.shape[-1], _ptr(stats), _stream())
        _lib.check(rc, "sd_conv_gemm")
    if B > 0:
        _count()
    if stats is not None and not getattr(out, "gn_fused", False):
        out.gn_stats = (stats, (H * W) // 128)
    return out


def set_gn_fuse(level):
    """sd_set_gn_fuse: 0 / 1 / 2 (see the header); returns the previous level."""
    prev = _lib.load().sd_set_gn_fuse(int(level))
    if prev < 0:
        _lib.check(prev, "sd_set_gn_fuse")
    return prev


def conv_gemm_s2(x, weight, bias=None, out=None, want_stats=False, split=False):
    """3x3 stride-2 SAME conv (flax pad (0,1)) with the stride in the TMA descriptor (sd_conv_gemm_s2).
    x: bf16 [B,H,W,C]; weight: bf16 [N, 9C]."""
    lib = _lib.load()
    _bf16c(x, "x"); _bf16c(weight, "weight")
    B, H, W, C = x.shape
    sm = 2 if split else 1
    C //= sm
    N = weight.shape[0]
    Ho, Wo = H // 2, W // 2
    if out is None:
        out = torch.empty(B, Ho, Wo, sm * N, device=x.device, dtype=torch.bfloat16)
    stats = None
    if want_stats and (Ho * Wo) % 128 == 0 and N % 16 == 0 and B > 0:
        stats = torch.empty(B, (Ho * Wo) // 128, 2, N, device=x.device, dtype=torch.float32)
    rc = lib.sd_conv_gemm_s2(_ptr(x), B, H, W, C, _ptr(weight), N, _ptr(bias), GEMM_SPLIT3 if split else 0, _ptr(out), _ptr(stats),
                             _stream())
    _lib.check(rc, "sd_conv_gemm_s2")
    if B > 0:
        _count()
    if stats is not None:
        out.gn_stats = (stats, (Ho * Wo) // 128)
    return out


def upconv_weights(kernel_hwio):
    """Flax HWIO [3,3,C,N] kernel of the conv that follows a nearest x2 upsample -> fp32 [4, N, 4*C] phase weights:
    for output phase (a, b) the 3x3 taps that land on the same low-resolution source pixel are pre-summed."""
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}      # phase -> taps hitting source offset (phase-1), (phase)
    out = []
    for a in (0, 1):
        for b in (0, 1):
            taps = []
            for r in (0, 1):
                for c in (0, 1):
                    w = sum(kernel_hwio[kh, kw] for kh in rows[a][r] for kw in rows[b][c])    # [C, N]
                    taps.append(w.T)                                                            # [N, C]
            out.append(torch.cat(taps, dim=1))                                                  # [N, 4C]
    return torch.stack(out)


def upconv_gemm(x, w4, bias=None, out=None, want_stats=False, split=False):
    """Fused nearest-x2 upsample + 3x3 conv as four 2x2-tap implicit GEMMs (sd_upconv_gemm).
    x: bf16 [B,H,W,C]; w4: bf16 [4, N, 4C] from upconv_weights()."""
    lib = _lib.load()
    _bf16c(x, "x"); _bf16c(w4, "w4")
    B, H, W, C = x.shape
    sm = 2 if split else 1
    C //= sm
    N = w4.shape[1]
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, sm * N, device=x.device, dtype=torch.bfloat16)
    stats = None
    if want_stats and (H * W) % 128 == 0 and N % 16 == 0 and B > 0:
        stats = torch.empty(B, 4 * (H * W) // 128, 2, N, device=x.device, dtype=torch.float32)
    rc = lib.sd_upconv_gemm(_ptr(x), B, H, W, C, _ptr(w4), N, _ptr(bias), GEMM_SPLIT3 if split else 0, _ptr(out), _ptr(stats),
                            _stream())
    _lib.check(rc, "sd_upconv_gemm")
    if B > 0:
        _count(1)         # the four phases run as one launch (phase = slowest digit of the tile index)
    if stats is not None:
        out.gn_stats = (stats, 4 * (H * W) // 128)
    return out


def batched_gemm(A, Bt, bias=None, residual=None, swish=False, out_f32=False, out=None, K=None, want_stats=False, split=False):
    """out[b] = A[b] @ Bt[b]^T (sd_batched_gemm).  A: bf16 [batch, M, >=K] or [M, K]
    (shared), Bt: bf16 [batch, N, >=K] or [N, K] (shared); row strides may exceed K.
    split: A, Bt, residual and (unless out_f32) out are hi|lo pairs along their last axis (SD_GEMM_SPLIT3); K, N logical."""
    lib = _lib.load()
    a3 = A if A.dim() == 3 else A.unsqueeze(0)
    b3 = Bt if Bt.dim() == 3 else Bt.unsqueeze(0)
    batch = max(a3.shape[0], b3.shape[0])
    M, N = a3.shape[1], b3.shape[1]
    if K is None:
        K = a3.shape[2] // (2 if split else 1)
    for t in (a3, b3):
        if t.dtype != torch.bfloat16 or not t.is_cuda or t.stride(2) != 1:
            raise ValueError("operands must be bf16 CUDA tensors with unit inner stride")
    sA = a3.stride(0) if (A.dim() == 3 and a3.shape[0] > 1) else 0
    sB = b3.stride(0) if (Bt.dim() == 3 and b3.shape[0] > 1) else 0
    if out is None:
        out = torch.empty(batch, M, N if (out_f32 or not split) else 2 * N, device=A.device,
                          dtype=torch.float32 if out_f32 else torch.bfloat16)
    flags = (EPI_SWISH if swish else 0) | (EPI_OUT_F32 if out_f32 else 0) | (GEMM_SPLIT3 if split else 0)
    stats = None
    if want_stats and M % 128 == 0 and N % 16 == 0 and not out_f32 and batch > 0:
        stats = torch.empty(batch, M // 128, 2, N, device=A.device, dtype=torch.float32)
    if stats is not None:
        rc = lib.sd_batched_gemm_stats(_ptr(a3), a3.stride(1), sA, _ptr(b3), b3.stride(1), sB, batch, M, N, K,
                                       _ptr(bias), _ptr(residual), flags, _ptr(out), out.stride(-2),
                                       out.stride(0) if out.dim() == 3 else 0, _ptr(stats), _stream())
    else:
        rc = lib.sd_batched_gemm(_ptr(a3), a3.stride(1), sA, _ptr(b3), b3.stride(1), sB, batch, M, N, K,
                                 _ptr(bias), _ptr(residual), flags, _ptr(out), out.stride(-2),
                                 out.stride(0) if out.dim() == 3 else 0, _stream())
    _lib.check(rc, "sd_batched_gemm")
    if batch > 0:
        _count()
    if stats is not None:
        out.gn_stats = (stats, M // 128)
    return out


def attention_probs(q, k, scale, block=None, out=None, C=None):
    """P = blockdiag-softmax(scale * q k^T) in one tcgen05 launch (sd_attention_probs).  q, k: bf16 [batch, S, >=C]
    views with unit inner stride; returns bf16 [batch, S, S]."""
    lib = _lib.load()
    batch, S = q.shape[0], q.shape[1]
    C = q.shape[2] if C is None else C
    block = S if block is None else block
    if out is None:
        out = torch.empty(batch, S, S, device=q.device, dtype=torch.bfloat16)
    rc = lib.sd_attention_probs(_ptr(q), q.stride(1), q.stride(0), _ptr(k), k.stride(1), k.stride(0), batch, S, C,
                                float(scale), int(block), _ptr(out), _stream())
    _lib.check(rc, "sd_attention_probs")
    if batch > 0:
        _count()
    return out


def attention_core(q, k, vt, scale, block=None, bias=None, residual=None, want_stats=False, C=None):
    """out = blockdiag-softmax(scale * q k^T) v + bias + residual in ONE launch (sd_attention_core).  q, k: bf16 [batch, S, >=C]
    views with unit inner stride, vt: bf16 [batch, C, S] (V transposed), residual: bf16 [batch, S, C]; returns bf16 [batch, S, C]."""
    lib = _lib.load()
    batch, S = q.shape[0], q.shape[1]
    C = q.shape[2] if C is None else C
    block = S if block is None else block
    for t in (q, k, vt):
        if t.dtype != torch.bfloat16 or not t.is_cuda or t.stride(2) != 1:
            raise ValueError("operands must be bf16 CUDA tensors with unit inner stride")
    if residual is not None:
        _bf16c(residual, "residual")
    out = torch.empty(batch, S, C, device=q.device, dtype=torch.bfloat16)
    stats = None
    if want_stats and block == S and batch > 0:
        stats = torch.empty(batch, S // 128, 2, C, device=q.device, dtype=torch.float32)
    rc = lib.sd_attention_core(_ptr(q), q.stride(1), q.stride(0), _ptr(k), k.stride(1), k.stride(0), _ptr(vt), vt.stride(1),
                               vt.stride(0), batch, S, C, float(scale), int(block), _ptr(bias), _ptr(residual), _ptr(out),
                               _ptr(stats), _stream())
    _lib.check(rc, "sd_attention_core")
    if batch > 0:
        _count()
    if stats is not None:
        out.gn_stats = (stats, S // 128)
    return out


def softmax_rows(x, scale, out=None):
    lib = _lib.load()
    _f32c(x, "x")
    cols = x.shape[-1]
    rows = x.numel() // cols
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_softmax_rows(_ptr(x), _ptr(out), rows, cols, float(scale), _stream()), "sd_softmax_rows")
    _count()
    return out


def softmax_rows_split(x, scale, block=None, out=None):
    """FP32-faithful attention probabilities (sd_softmax_rows_split): x fp32 [batch, S, S] -> hi|lo bf16 pair [batch, S, 2S];
    softmax over the diagonal block of ``block`` columns each row belongs to, zeros elsewhere."""
    lib = _lib.load()
    _f32c(x, "x")
    batch, S, cols = x.shape
    block = cols if block is None else block
    if out is None:
        out = torch.empty(batch, S, 2 * cols, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_softmax_rows_split(_ptr(x), _ptr(out), batch * S, cols, float(scale), int(block), S, _stream()),
               "sd_softmax_rows_split")
    if batch > 0:
        _count()
    return out


_gn_scratch = {}


def _gn_scratch_for(device, floats):
    # one buffer per (device, stream): stats -> apply are back to back on a stream, but two score-nets may run on
    # two streams concurrently (SuperDiffSampler)
    key = (device, torch.cuda.current_stream().cuda_stream)
    buf = _gn_scratch.get(key)
    if buf is None or buf.numel() < floats:
        buf = torch.empty(max(floats, 1 << 20), device=device, dtype=torch.float32)
        _gn_scratch[key] = buf
    return buf


def groupnorm_swish(x0, gamma, beta, x1=None, eps=1e-6, swish=True, out=None, split=False):
    """GroupNorm(32) + swish over concat(x0, x1) (sd_groupnorm_swish).  Sources carrying ``gn_stats`` (set by
    conv_gemm(want_stats=True)) skip the statistics pass.  split: sources and out are hi|lo pairs [B,H,W,2C], exact swish."""
    lib = _lib.load()
    _bf16c(x0, "x0")
    B = x0.shape[0]
    HW = x0.shape[1] * x0.shape[2]
    sm = 2 if split else 1
    C0 = x0.shape[3] // sm
    C1 = 0
    if x1 is not None:
        _bf16c(x1, "x1")
        C1 = x1.shape[3] // sm
    if out is None:
        out = torch.empty(B, x0.shape[1], x0.shape[2], sm * (C0 + C1), device=x0.device, dtype=torch.bfloat16)
    st0, n0 = getattr(x0, "gn_stats", (None, 0))
    st1, n1 = getattr(x1, "gn_stats", (None, 0)) if x1 is not None else (None, 0)
    # channel sums of sources without stats + group stats; stream-ordered reuse of one buffer per (device, stream)
    scratch = _gn_scratch_for(x0.device, (4736 + B) * 2 * (C0 + C1) + 64 * B)
    rc = lib.sd_groupnorm_swish_ex(_ptr(x0), C0, _ptr(x1), C1, B, HW, _ptr(_f32c(gamma, "gamma")),
                                   _ptr(_f32c(beta, "beta")), float(eps), int(bool(swish)), _ptr(st0), int(n0),
                                   _ptr(st1), int(n1), _ptr(scratch), scratch.numel(), _ptr(out),
                                   GEMM_SPLIT3 if split else 0, _stream())
    _lib.check(rc, "sd_groupnorm_swish")
    if B > 0:
        C, nv = C0 + C1, HW * (C0 + C1) // 8
        small = st0 is None and st1 is None and C in (256, 512) and nv % 256 == 0 and nv // 256 in (1, 2, 4, 8, 16)
        _count(1 if small else 1 + (st0 is None) + (x1 is not None and st1 is None))
    return out


def groupnorm_swish_jvp(x0, dx0, gamma, beta, x1=None, dx1=None, eps=1e-6, swish=True):
    """(out, dout) = JVP of GroupNorm(32)+swish over concat(x0, x1) in direction concat(dx0, dx1) (sd_groupnorm_swish_jvp)."""
    lib = _lib.load()
    _bf16c(x0, "x0"); _bf16c(dx0, "dx0")
    B, HW, C0 = x0.shape[0], x0.shape[1] * x0.shape[2], x0.shape[3]
    C1 = 0
    if x1 is not None:
        _bf16c(x1, "x1"); _bf16c(dx1, "dx1")
        C1 = x1.shape[3]
    out = torch.empty(B, x0.shape[1], x0.shape[2], C0 + C1, device=x0.device, dtype=torch.bfloat16)
    dout = torch.empty_like(out)
    scratch = _gn_scratch_for(x0.device, 4 * max(B, 1) * 64 * (C0 + C1))
    rc = lib.sd_groupnorm_swish_jvp(_ptr(x0), _ptr(dx0), C0, _ptr(x1), _ptr(dx1), C1, B, HW, _ptr(_f32c(gamma, "gamma")),
                                    _ptr(_f32c(beta, "beta")), float(eps), int(bool(swish)), _ptr(scratch), scratch.numel(),
                                    _ptr(out), _ptr(dout), _stream())
    _lib.check(rc, "sd_groupnorm_swish_jvp")
    if B > 0:
        _count(2)
    return out, dout


def softmax_jvp(P, dS1, dS2, scale):
    """dP = P * (dS - rowsum(P * dS)), dS = scale * (dS1 + dS2) (sd_softmax_jvp); P bf16, dS fp32, same shape."""
    lib = _lib.load()
    _bf16c(P, "P"); _f32c(dS1, "dS1")
    if dS2 is not None:
        _f32c(dS2, "dS2")
    cols = P.shape[-1]
    rows = P.numel() // cols
    dP = torch.empty_like(P)
    _lib.check(lib.sd_softmax_jvp(_ptr(P), _ptr(dS1), _ptr(dS2), float(scale), _ptr(dP), rows, cols, _stream()), "sd_softmax_jvp")
    if rows > 0:
        _count()
    return dP


def attention_small(qkv, C, out=None):
    lib = _lib.load()
    _bf16c(qkv, "qkv")
    B, S = qkv.shape[0], qkv.shape[1]
    if out is None:
        out = torch.empty(B, S, C, device=qkv.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_attention(_ptr(qkv), B, S, C, _ptr(out), _stream()), "sd_attention")
    if B > 0:
        _count()
    return out


def upsample2x(x, out=None):
    lib = _lib.load()
    _bf16c(x, "x")
    B, H, W, C = x.shape
    if out is None:
        out = torch.empty(B, 2 * H, 2 * W, C, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_upsample2x(_ptr(x), B, H, W, C, _ptr(out), _stream()), "sd_upsample2x")
    if B > 0:
        _count()
    return out


def im2col_s2(x, out=None):
    lib = _lib.load()
    _bf16c(x, "x")
    B, H, W, C = x.shape
    if out is None:
        out = torch.empty(B, H // 2, W // 2, 9 * C, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_im2col_s2(_ptr(x), B, H, W, C, _ptr(out), _stream()), "sd_im2col_s2")
    if B > 0:
        _count()
    return out


def gather_row(table, counter, out=None):
    """out = table[counter] (sd_gather_row); table: fp32 [rows, n] on the device, counter: int32 device scalar."""
    lib = _lib.load()
    _f32c(table, "table")
    rows, n = table.shape
    if out is None:
        out = torch.empty(1, n, device=table.device, dtype=torch.float32)
    _lib.check(lib.sd_gather_row(_ptr(table), rows, n, _ptr(counter), _ptr(out), _stream()), "sd_gather_row")
    _count()
    return out


def im2col_in(x, out=None, split=False):
    """fp32 NHWC [B,H,W,Cin<=3] -> bf16 [B,H,W,64] hi/lo-split 3x3 neighbourhoods (sd_im2col_in); split: [B,H,W,128] with a
    zero upper half (the block as the hi half of a hi|lo pair)."""
    lib = _lib.load()
    _f32c(x, "x")
    B, H, W, Cin = x.shape
    if out is None:
        out = torch.empty(B, H, W, 128 if split else 64, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_im2col_in_ex(_ptr(x), B, H, W, Cin, _ptr(out), GEMM_SPLIT3 if split else 0, _stream()), "sd_im2col_in")
    if B > 0:
        _count()
    return out


def conv_in_weights(w_hwio):
    """Flax HWIO [3,3,Cin,Cout] fp32 -> fp32 [Cout, 64] = [w | w | 0] matching im2col_in's [hi | lo | 0] rows."""
    cout = w_hwio.shape[-1]
    w = w_hwio.reshape(-1, cout).T                       # [Cout, 9*Cin], (kh, kw, c) order
    pad = torch.zeros(cout, 64 - 2 * w.shape[1], dtype=w.dtype, device=w.device)
    return torch.cat([w, w, pad], dim=1)


def conv_in_weights_split(w_hwio):
    """Flax HWIO [3,3,Cin,Cout] fp32 -> bf16 [Cout, 128] = [w_hi | w_hi | 0 || w_lo | 0 | 0] for im2col_in(split=True) rows
    [x_hi | x_lo | 0 || 0]: x_hi w_hi + x_lo w_hi + x_hi w_lo."""
    cout = w_hwio.shape[-1]
    w = w_hwio.reshape(-1, cout).T.to(torch.float32)     # [Cout, 9*Cin]
    k = w.shape[1]
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    z = lambda n: torch.zeros(cout, n, dtype=torch.bfloat16, device=w.device)
    return torch.cat([hi, hi, z(64 - 2 * k), lo, z(64 - k)], dim=1).contiguous()


def conv_in(x, w_hwio, bias, out=None):
    lib = _lib.load()
    _f32c(x, "x")
    B, H, W, Cin = x.shape
    Cout = w_hwio.shape[-1]
    if out is None:
        out = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.bfloat16)
    rc = lib.sd_conv_in(_ptr(x), B, H, W, Cin, _ptr(_f32c(w_hwio, "w")), _ptr(bias), Cout, _ptr(out), _stream())
    _lib.check(rc, "sd_conv_in")
    if B > 0:
        _count()
    return out


def time_embedding(B, nf, w0, b0, w1, b1, t=None, t_stride=0, sched=None, step_counter=None, class_emb=None,
                   labels=None, scratch=None, out=None, split=False):
    lib = _lib.load()
    dev = w0.device
    shared = sched is not None or t_stride == 0
    if scratch is None:
        scratch = torch.empty(1 if shared else B, 4 * nf, device=dev, dtype=torch.float32)
    if out is None:
        out = torch.empty(B, (8 if split else 4) * nf, device=dev, dtype=torch.bfloat16)
    rc = lib.sd_time_embedding_ex(_ptr(t), int(t_stride), _ptr(sched), _ptr(step_counter), B, nf, _ptr(w0), _ptr(b0),
                                  _ptr(w1), _ptr(b1), _ptr(class_emb), _ptr(labels), _ptr(scratch), _ptr(out),
                                  GEMM_SPLIT3 if split else 0, _stream())
    _lib.check(rc, "sd_time_embedding")
    if B > 0:
        _count(2)
    return out


def cast_bf16(x, out=None):
    lib = _lib.load()
    _f32c(x, "x")
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.sd_cast_f32_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "sd_cast_f32_to_bf16")
    _count()
    return out
