"""GPU parity of the score-net building blocks against plain torch math on the
same bf16-rounded inputs (fp32/fp64 reference; tolerance = bf16 output rounding,
2^-8 relative, plus fp32 accumulation-order noise)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from super_diffusion_b200 import ops

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


def _close(got, ref, rtol=1.2e-2, atol=None):
    got = got.float().cpu().double()
    ref = ref.double().cpu()
    if atol is None:
        atol = 1.2e-2 * ref.abs().max().item() / 4 + 1e-6
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert not bad.any(), (err.max().item(), ref.abs().max().item(), int(bad.sum()))


def _conv_ref(x, w_nk, taps, C):
    # x: [B,H,W,C] float; w_nk: [N, taps*C] float (tap-major, then channel)
    N = w_nk.shape[0]
    if taps == 1:
        return torch.einsum("bhwc,nc->bhwn", x, w_nk)
    w = w_nk.reshape(N, 3, 3, C).permute(0, 3, 1, 2)  # OIHW
    return F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)


@pytest.mark.parametrize("B,H,W,C,N,taps", [
    (2, 32, 32, 128, 128, 9), (3, 16, 16, 256, 256, 9), (5, 8, 8, 128, 256, 9), (9, 4, 4, 256, 256, 9),
    (2, 32, 32, 64, 16, 9), (2, 32, 32, 128, 3, 9), (4, 16, 16, 256, 768, 1), (2, 32, 32, 256, 128, 1),
    (1, 32, 32, 64, 64, 9), (16, 4, 4, 512, 256, 1), (3, 8, 8, 256, 48, 1),
    (601, 8, 8, 64, 64, 9),     # 301 m-tiles (odd) -> dual m-tile mode with a masked tail tile
    (40, 32, 32, 64, 128, 9),   # dual m-tile mode, 320 m-tiles
])
def test_conv_gemm_single_source(cuda, B, H, W, C, N, taps):
    g = torch.Generator().manual_seed(B * 1000 + H + C + N)
    x = _bf(torch.randn(B, H, W, C, generator=g))
    w = _bf(torch.randn(N, taps * C, generator=g) / math.sqrt(taps * C))
    bias = torch.randn(N, generator=g)
    out = ops.conv_gemm([(x.to(cuda), taps)], w.to(cuda), bias=bias.to(cuda))
    torch.cuda.synchronize()
    ref = _conv_ref(x.double(), w.double(), taps, C) + bias.double()
    assert out.shape == (B, H, W, N)
    _close(out, ref)


def test_conv_gemm_fused_resblock_tail(cuda):
    """conv2 (9 taps) + NIN shortcut over a 2-tensor concat (1 tap each) + bias + time-embedding row bias,
    i.e. the tail of ResnetBlockDDPM (cifar/models/layers.py:556-565) as ONE GEMM; then residual/swish/f32 flags."""
    B, H, W, Co, C0, C1 = 3, 16, 16, 256, 256, 128
    g = torch.Generator().manual_seed(7)
    a2 = _bf(torch.randn(B, H, W, Co, generator=g))
    xa = _bf(torch.randn(B, H, W, C0, generator=g))
    xb = _bf(torch.randn(B, H, W, C1, generator=g))
    w_conv = torch.randn(Co, 9 * Co, generator=g) / math.sqrt(9 * Co)
    w_nin = torch.randn(Co, C0 + C1, generator=g) / math.sqrt(C0 + C1)
    w = _bf(torch.cat([w_conv, w_nin], dim=1))
    bias = torch.randn(Co, generator=g)
    rowbias = torch.randn(B, 640, generator=g)
    out = ops.conv_gemm([(a2.to(cuda), 9), (xa.to(cuda), 1), (xb.to(cuda), 1)], w.to(cuda), bias=bias.to(cuda),
                        rowbias=rowbias.to(cuda)[:, 128:128 + Co])
    wd = w.double()
    ref = _conv_ref(a2.double(), wd[:, :9 * Co], 9, Co) \
        + torch.einsum("bhwc,nc->bhwn", torch.cat([xa, xb], -1).double(), wd[:, 9 * Co:]) \
        + bias.double() + rowbias.double()[:, None, None, 128:128 + Co]
    _close(out, ref)
    res = _bf(torch.randn(B, H, W, Co, generator=g))
    out2 = ops.conv_gemm([(a2.to(cuda), 9)], w[:, :9 * Co].contiguous().to(cuda), bias=bias.to(cuda),
                         residual=res.to(cuda), swish=True, out_f32=True)
    r2 = _conv_ref(a2.double(), wd[:, :9 * Co], 9, Co) + bias.double() + res.double()
    r2 = r2 * torch.sigmoid(r2)
    assert out2.dtype == torch.float32
    _close(out2, r2, rtol=2e-3, atol=2e-3)


def test_conv_gemm_many_tiles_persistent(cuda):
    """More tiles than SMs (B=64 at 32x32 -> 512 tiles): exercises the persistent loop, the smem ring
    wrap-around and the two TMEM accumulators."""
    B, H, W, C, N = 64, 32, 32, 128, 128
    g = torch.Generator().manual_seed(3)
    x = _bf(torch.randn(B, H, W, C, generator=g))
    w = _bf(torch.randn(N, 9 * C, generator=g) / math.sqrt(9 * C))
    out = ops.conv_gemm([(x.to(cuda), 9)], w.to(cuda))
    ref = F.conv2d(x.to(cuda).float().permute(0, 3, 1, 2), w.to(cuda).float().reshape(N, 3, 3, C).permute(0, 3, 1, 2),
                   padding=1).permute(0, 2, 3, 1)
    _close(out, ref.cpu())


@pytest.mark.parametrize("batch,M,N,K,shareA,shareB", [
    (1, 512, 640, 512, False, True), (1, 100, 256, 128, False, True), (6, 256, 256, 256, False, False),
    (5, 256, 256, 256, True, False), (3, 128, 64, 64, False, False), (2, 200, 40, 192, False, False),
    (1, 128 * 299, 64, 64, False, True), (160, 256, 128, 64, False, False),   # dual m-tile mode (odd tail / batched)
])
def test_batched_gemm(cuda, batch, M, N, K, shareA, shareB):
    g = torch.Generator().manual_seed(batch + M + N + K)
    A = _bf(torch.randn(*( (M, K) if shareA else (batch, M, K)), generator=g))
    Bt = _bf(torch.randn(*( (N, K) if shareB else (batch, N, K)), generator=g) / math.sqrt(K))
    bias = torch.randn(N, generator=g)
    out = ops.batched_gemm(A.to(cuda), Bt.to(cuda), bias=bias.to(cuda), out_f32=True)
    Ad = A.double() if A.dim() == 3 else A.double()[None]
    Bd = Bt.double() if Bt.dim() == 3 else Bt.double()[None]
    ref = torch.einsum("bmk,bnk->bmn", Ad.expand(batch, -1, -1), Bd.expand(batch, -1, -1)) + bias.double()
    _close(out, ref, rtol=1e-4, atol=1e-4)


def test_batched_gemm_strided_qkv_views(cuda):
    """q k^T on strided views of a packed [B, S, 3C] tensor (the attention layout)."""
    B, S, C = 4, 256, 256
    g = torch.Generator().manual_seed(11)
    qkv = _bf(torch.randn(B, S, 3 * C, generator=g)).to(cuda)
    q, k = qkv[:, :, :C], qkv[:, :, C:2 * C]
    out = ops.batched_gemm(q, k, out_f32=True, K=C)
    ref = torch.einsum("bsc,btc->bst", q.double(), k.double())
    _close(out, ref.cpu(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("B,H,C0,C1,swish", [(3, 32, 128, 0, True), (2, 32, 256, 128, True), (2, 16, 256, 256, True),
                                            (5, 8, 256, 0, False), (9, 4, 256, 256, True), (2, 32, 128, 128, True),
                                            (2, 16, 128, 0, True), (1, 16, 256, 128, True),
                                            # single-kernel register-resident form (8x8 / 4x4 levels, C = 256 / 512)
                                            (7, 8, 256, 0, True), (5, 8, 256, 256, True), (6, 4, 256, 0, True), (3, 8, 512, 0, True)])
def test_groupnorm_swish(cuda, B, H, C0, C1, swish):
    g = torch.Generator().manual_seed(B + H + C0 + C1)
    x0 = _bf(1.5 * torch.randn(B, H, H, C0, generator=g) + 0.3)
    x1 = _bf(0.7 * torch.randn(B, H, H, C1, generator=g) - 0.2) if C1 else None
    C = C0 + C1
    gamma = 1 + 0.1 * torch.randn(C, generator=g)
    beta = 0.1 * torch.randn(C, generator=g)
    out = ops.groupnorm_swish(x0.to(cuda), gamma.to(cuda), beta.to(cuda), x1=x1.to(cuda) if C1 else None, swish=swish)
    x = torch.cat([x0, x1], -1).double() if C1 else x0.double()
    xr = x.reshape(B, H * H, 32, C // 32)
    mean = xr.mean(dim=(1, 3), keepdim=True)
    var = (xr * xr).mean(dim=(1, 3), keepdim=True) - mean * mean
    y = ((xr - mean) / torch.sqrt(var + 1e-6)).reshape(B, H, H, C) * gamma.double() + beta.double()
    if swish:
        y = y * torch.sigmoid(y)
    _close(out, y, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("S", [16, 64])
def test_attention_small(cuda, S):
    B, C = 5, 256
    g = torch.Generator().manual_seed(S)
    qkv = _bf(torch.randn(B, S, 3 * C, generator=g))
    out = ops.attention_small(qkv.to(cuda), C)
    q, k, v = qkv.double().split(C, dim=-1)
    w = torch.softmax(torch.einsum("bsc,btc->bst", q, k) * C ** -0.5, dim=-1)
    _close(out, torch.einsum("bst,btc->bsc", w, v), rtol=1e-2, atol=1e-2)


def test_softmax_rows(cuda):
    x = torch.randn(3, 256, 256) * 8
    out = ops.softmax_rows(x.to(cuda), 1 / 16)
    _close(out, torch.softmax(x.double() / 16, -1), rtol=1e-2, atol=1e-4)


def test_upsample_im2col_convin_temb(cuda):
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(3, 8, 8, 256, generator=g))
    up = ops.upsample2x(x.to(cuda))
    assert torch.equal(up.cpu(), x.repeat_interleave(2, 1).repeat_interleave(2, 2))
    col = ops.im2col_s2(x.to(cuda)).cpu()
    xp = F.pad(x, (0, 0, 0, 1, 0, 1))
    ref = torch.stack([xp[:, kh:kh + 8:2, kw:kw + 8:2, :] for kh in range(3) for kw in range(3)], dim=3).reshape(3, 4, 4, 9 * 256)
    assert torch.equal(col, ref)
    xi = torch.randn(4, 32, 32, 3, generator=g)
    w = torch.randn(3, 3, 3, 128, generator=g) / math.sqrt(27)
    b = torch.randn(128, generator=g)
    ci = ops.conv_in(xi.to(cuda), w.to(cuda), b.to(cuda))
    ref = F.conv2d(xi.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(), padding=1).permute(0, 2, 3, 1)
    _close(ci, ref, rtol=1e-2, atol=1e-2)
    # time embedding (cifar/models/layers.py:450-461 + ddpm.py:64-68) -> swish(temb) in bf16
    nf, B = 128, 6
    w0, b0 = torch.randn(nf, 4 * nf, generator=g) / math.sqrt(nf), 0.1 * torch.randn(4 * nf, generator=g)
    w1, b1 = torch.randn(4 * nf, 4 * nf, generator=g) / math.sqrt(4 * nf), 0.1 * torch.randn(4 * nf, generator=g)
    emb_tab = torch.randn(10, 4 * nf, generator=g)
    labels = torch.arange(B, dtype=torch.int32) % 10
    for t_val, per_sample in ((0.37, False), (None, True)):
        tt = torch.rand(B, generator=g) if per_sample else torch.full((1,), t_val)
        got = ops.time_embedding(B, nf, w0.to(cuda), b0.to(cuda), w1.to(cuda), b1.to(cuda), t=tt.to(cuda),
                                 t_stride=1 if per_sample else 0, class_emb=emb_tab.to(cuda), labels=labels.to(cuda))
        tv = (tt if per_sample else tt.expand(B)).double()
        half = nf // 2
        f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).double()
        e = torch.cat([torch.sin(tv[:, None] * f), torch.cos(tv[:, None] * f)], 1)
        h = e @ w0.double() + b0.double()
        h = (h * torch.sigmoid(h)) @ w1.double() + b1.double() + emb_tab.double()[labels.long()]
        _close(got, h * torch.sigmoid(h), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("nb,S,block", [(4, 256, 256), (6, 128, 64), (5, 128, 16), (3, 128, 128)])
def test_attention_probs_fused_softmax(cuda, nb, S, block):
    """q k^T -> scaled, block-diagonal softmax in the GEMM epilogue (scores never leave tensor memory)."""
    C = 256
    g = torch.Generator().manual_seed(nb + S + block)
    qk = _bf(torch.randn(nb, S, 2 * C, generator=g)).to(cuda)
    p = ops.attention_probs(qk[:, :, :C], qk[:, :, C:], C ** -0.5, block=block, C=C)
    sc = torch.einsum("bsc,btc->bst", qk[:, :, :C].double(), qk[:, :, C:].double()) * C ** -0.5
    idx = torch.arange(S, device=cuda) // block
    mask = idx[:, None] == idx[None, :]
    ref = torch.softmax(sc.masked_fill(~mask, float("-inf")), dim=-1)
    assert (p.double()[:, ~mask] == 0).all()
    _close(p, ref.cpu(), rtol=1e-2, atol=1e-4)
    assert torch.allclose(p.float().sum(-1), torch.ones(nb, S, device=cuda), atol=2e-2)


def test_conv_gemm_swapped_operands_epilogue(cuda):
    """N = 128 layers run with swapped operands (D[cout][pixel]); exercise its epilogue: multi-source K,
    bias, per-image row bias, residual, swish, fp32 output, odd pair tail (B = 37 -> 296 m-tiles + ...)."""
    B, H, W, Co, C0, C1 = 39, 32, 32, 128, 128, 64
    g = torch.Generator().manual_seed(21)
    a2 = _bf(torch.randn(B, H, W, Co, generator=g))
    xa = _bf(torch.randn(B, H, W, C0, generator=g))
    xb = _bf(torch.randn(B, H, W, C1, generator=g))
    w = _bf(torch.cat([torch.randn(Co, 9 * Co, generator=g) / math.sqrt(9 * Co),
                       torch.randn(Co, C0 + C1, generator=g) / math.sqrt(C0 + C1)], dim=1))
    bias = torch.randn(Co, generator=g)
    rowbias = torch.randn(B, 384, generator=g)
    res = _bf(torch.randn(B, H, W, Co, generator=g))
    wd = w.to(cuda).float()
    a2d, xad, xbd = a2.to(cuda), xa.to(cuda), xb.to(cuda)
    ref = F.conv2d(a2d.float().permute(0, 3, 1, 2), wd[:, :9 * Co].reshape(Co, 3, 3, Co).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) \
        + torch.einsum("bhwc,nc->bhwn", torch.cat([xad, xbd], -1).float(), wd[:, 9 * Co:]) \
        + bias.to(cuda) + rowbias.to(cuda)[:, None, None, 256:256 + Co]
    out = ops.conv_gemm([(a2d, 9), (xad, 1), (xbd, 1)], w.to(cuda), bias=bias.to(cuda), rowbias=rowbias.to(cuda)[:, 256:256 + Co])
    _close(out, ref.cpu())
    ref2 = ref + res.to(cuda).float()
    ref2 = ref2 * torch.sigmoid(ref2)
    out2 = ops.conv_gemm([(a2d, 9), (xad, 1), (xbd, 1)], w.to(cuda), bias=bias.to(cuda), rowbias=rowbias.to(cuda)[:, 256:256 + Co],
                         residual=res.to(cuda), swish=True, out_f32=True)
    _close(out2, ref2.cpu(), rtol=2e-3, atol=5e-3)


@pytest.mark.parametrize("B,H,C,N,taps", [(40, 32, 64, 128, 9), (6, 16, 128, 256, 9), (3, 32, 128, 128, 1)])
def test_groupnorm_from_gemm_emitted_statistics(cuda, B, H, C, N, taps):
    """conv_gemm(want_stats=True) emits per-tile channel sums in its epilogue; GroupNorm fed with them (no statistics
    pass) equals GroupNorm that measures the tensor itself, also for a concat of one source with and one without stats."""
    g = torch.Generator().manual_seed(B + H + N)
    x = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    w = _bf(torch.randn(N, taps * C, generator=g) / math.sqrt(taps * C)).to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    res = _bf(torch.randn(B, H, H, N, generator=g)).to(cuda)
    y = ops.conv_gemm([(x, taps)], w, bias=bias, residual=res, want_stats=True)
    assert hasattr(y, "gn_stats") and y.gn_stats[0].shape == (B, H * H // 128, 2, N)
    st = y.gn_stats[0]
    yf = y.float().reshape(B, H * H // 128, 128, N)
    assert torch.allclose(st[:, :, 0], yf.sum(2), rtol=2e-2, atol=0.5)          # sums of the (bf16-rounded) output
    assert torch.allclose(st[:, :, 1], (yf * yf).sum(2), rtol=2e-2, atol=0.5)
    gamma = (1 + 0.1 * torch.randn(N, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(N, generator=g)).to(cuda)
    a = ops.groupnorm_swish(y, gamma, beta)
    y_plain = y.clone()                                   # same data, no gn_stats attribute
    b = ops.groupnorm_swish(y_plain, gamma, beta)
    assert (a.float() - b.float()).abs().max().item() <= 6.5e-2      # at most one bf16 ulp at |y| < 8
    skip = _bf(torch.randn(B, H, H, 64, generator=g)).to(cuda)
    g2 = (1 + 0.1 * torch.randn(N + 64, generator=g)).to(cuda)
    b2 = (0.1 * torch.randn(N + 64, generator=g)).to(cuda)
    a2 = ops.groupnorm_swish(y, g2, b2, x1=skip)
    r2 = ops.groupnorm_swish(y_plain, g2, b2, x1=skip)
    assert (a2.float() - r2.float()).abs().max().item() <= 6.5e-2


@pytest.mark.parametrize("B,H,C,N", [(3, 16, 128, 256), (5, 8, 256, 256), (9, 4, 64, 64), (40, 16, 64, 128)])
def test_upconv_gemm_equals_upsample_then_conv(cuda, B, H, C, N):
    """Nearest x2 upsample + 3x3 SAME conv (cifar/models/layers.py:514-523) as four 2x2-tap phase GEMMs."""
    g = torch.Generator().manual_seed(B + H + C + N)
    x = _bf(torch.randn(B, H, H, C, generator=g))
    k = torch.randn(3, 3, C, N, generator=g) / math.sqrt(9 * C)          # Flax HWIO
    bias = torch.randn(N, generator=g)
    w4 = ops.upconv_weights(k)
    assert w4.shape == (4, N, 4 * C)
    out = ops.upconv_gemm(x.to(cuda), w4.bfloat16().to(cuda), bias=bias.to(cuda), want_stats=True)
    up = x.double().repeat_interleave(2, 1).repeat_interleave(2, 2)
    ref = F.conv2d(up.permute(0, 3, 1, 2), k.double().permute(3, 2, 0, 1), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert out.shape == (B, 2 * H, 2 * H, N)
    _close(out, ref, rtol=1.5e-2)
    if (H * H) % 128 == 0:
        st, n = out.gn_stats
        assert n == 4 * H * H // 128
        assert torch.allclose(st[:, :, 0].sum(1), out.float().sum(dim=(1, 2)), rtol=2e-2, atol=1.0)
        gamma, beta = torch.ones(N, device=cuda), torch.zeros(N, device=cuda)
        a = ops.groupnorm_swish(out, gamma, beta)
        b = ops.groupnorm_swish(out.clone(), gamma, beta)
        assert (a.float() - b.float()).abs().max().item() <= 6.5e-2


@pytest.mark.parametrize("B,H,C,N", [(3, 32, 128, 128), (5, 16, 256, 256), (9, 8, 256, 256), (40, 32, 64, 64),
                                     (600, 32, 64, 128),      # 1200 output tiles: two-tile units, operands swapped ([channel][pixel] epilogue)
                                     (600, 16, 64, 256)])     # 8x8 outputs, 300 tiles: cta_group::2 pairs, swapped, a unit spans four images
def test_conv_gemm_stride2_tma(cuda, B, H, C, N):
    """Downsample conv (cifar/models/layers.py:533): 3x3, stride 2, SAME => pad (0,1); stride carried by the TMA descriptor.
    Also equals the explicit im2col gather + 1x1 GEMM path."""
    g = torch.Generator().manual_seed(B + H + C)
    x = _bf(torch.randn(B, H, H, C, generator=g))
    w = _bf(torch.randn(N, 9 * C, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(N, generator=g)
    out = ops.conv_gemm_s2(x.to(cuda), w.to(cuda), bias=bias.to(cuda), want_stats=True)
    xp = F.pad(x.double().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, w.double().reshape(N, 3, 3, C).permute(0, 3, 1, 2), bias.double(), stride=2).permute(0, 2, 3, 1)
    assert out.shape == (B, H // 2, H // 2, N)
    _close(out, ref)
    alt = ops.conv_gemm([(ops.im2col_s2(x.to(cuda)), 1)], w.to(cuda), bias=bias.to(cuda))
    assert torch.equal(out, alt)


def test_conv_gemm_pair_slab_multisegment(cuda):
    """Wide (N = 256) layer large enough for the cta_group::2 pair kernel with activation slabs (three vertical taps
    share one TMA load), followed by 1-tap segments (NIN shortcut over a concat), row bias and emitted GN statistics."""
    B, H, W, Co, C0, C1 = 152, 16, 16, 256, 128, 64
    g = torch.Generator().manual_seed(31)
    a2 = _bf(torch.randn(B, H, W, Co, generator=g)).to(cuda)
    xa = _bf(torch.randn(B, H, W, C0, generator=g)).to(cuda)
    xb = _bf(torch.randn(B, H, W, C1, generator=g)).to(cuda)
    w = _bf(torch.cat([torch.randn(Co, 9 * Co, generator=g) / math.sqrt(9 * Co),
                       torch.randn(Co, C0 + C1, generator=g) / math.sqrt(C0 + C1)], dim=1)).to(cuda)
    bias = torch.randn(Co, generator=g).to(cuda)
    rowbias = torch.randn(B, 512, generator=g).to(cuda)
    out = ops.conv_gemm([(a2, 9), (xa, 1), (xb, 1)], w, bias=bias, rowbias=rowbias[:, 256:256 + Co], want_stats=True)
    wf = w.float()
    ref = F.conv2d(a2.float().permute(0, 3, 1, 2), wf[:, :9 * Co].reshape(Co, 3, 3, Co).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) \
        + torch.einsum("bhwc,nc->bhwn", torch.cat([xa, xb], -1).float(), wf[:, 9 * Co:]) + bias + rowbias[:, None, None, 256:256 + Co]
    _close(out, ref.cpu())
    st = out.gn_stats[0]
    assert torch.allclose(st[:, :, 0].sum(1), out.float().sum(dim=(1, 2)), rtol=2e-2, atol=1.0)
    # odd number of m-tiles (pair tail) and a different channel count
    B2 = 149
    x = _bf(torch.randn(B2, H, W, 64, generator=g)).to(cuda)
    w2 = _bf(torch.randn(256, 9 * 64, generator=g) / math.sqrt(9 * 64)).to(cuda)
    out2 = ops.conv_gemm([(x, 9)], w2)
    ref2 = F.conv2d(x.float().permute(0, 3, 1, 2), w2.float().reshape(256, 3, 3, 64).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    _close(out2, ref2.cpu())


def test_conv_gemm_swapped_n128(cuda):
    """N = 128 conv layers with >= 2*148 tiles run operand-swapped (weights = M operand, 256 pixels = N operand, accumulator
    [channel][pixel]): activation slabs, 1-tap segments, bias + row bias, thread-local GN statistics, lane-pair packed stores."""
    B, H, W, Co, C0, C1 = 40, 32, 32, 128, 128, 64
    g = torch.Generator().manual_seed(47)
    a2 = _bf(torch.randn(B, H, W, Co, generator=g)).to(cuda)
    xa = _bf(torch.randn(B, H, W, C0, generator=g)).to(cuda)
    xb = _bf(torch.randn(B, H, W, C1, generator=g)).to(cuda)
    w = _bf(torch.cat([torch.randn(Co, 9 * Co, generator=g) / math.sqrt(9 * Co),
                       torch.randn(Co, C0 + C1, generator=g) / math.sqrt(C0 + C1)], dim=1)).to(cuda)
    bias = torch.randn(Co, generator=g).to(cuda)
    rowbias = torch.randn(B, 512, generator=g).to(cuda)
    out = ops.conv_gemm([(a2, 9), (xa, 1), (xb, 1)], w, bias=bias, rowbias=rowbias[:, 256:256 + Co], want_stats=True)
    wf = w.float()
    ref = F.conv2d(a2.float().permute(0, 3, 1, 2), wf[:, :9 * Co].reshape(Co, 3, 3, Co).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) \
        + torch.einsum("bhwc,nc->bhwn", torch.cat([xa, xb], -1).float(), wf[:, 9 * Co:]) + bias + rowbias[:, None, None, 256:256 + Co]
    _close(out, ref.cpu())
    st, nt = out.gn_stats
    assert nt == 8 and st.shape == (B, 8, 2, Co)
    tiles = out.float().reshape(B, 8, 128, Co)
    assert torch.allclose(st[:, :, 0], tiles.sum(2), rtol=2e-2, atol=0.5)
    assert torch.allclose(st[:, :, 1], (tiles ** 2).sum(2), rtol=2e-2, atol=0.5)
    # swish epilogue, no bias, 16x16 images (2 m-tiles per image = one unit per image), 1-tap only
    x = _bf(torch.randn(301, 16, 16, 64, generator=g)).to(cuda)
    w2 = _bf(torch.randn(128, 9 * 64, generator=g) / math.sqrt(9 * 64)).to(cuda)
    out2 = ops.conv_gemm([(x, 9)], w2, swish=True)
    ref2 = F.conv2d(x.float().permute(0, 3, 1, 2), w2.float().reshape(128, 3, 3, 64).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    _close(out2, (ref2 * torch.sigmoid(ref2)).cpu())
    w3 = _bf(torch.randn(128, 64, generator=g) / 8).to(cuda)
    out3 = ops.conv_gemm([(x, 1)], w3, bias=bias)
    _close(out3, (torch.einsum("bhwc,nc->bhwn", x.float(), w3.float()) + bias).cpu())


def test_conv_gemm_swapped_1x1_multi_ntile(cuda):
    """1x1 layers with N = 256 / 512 run as N = 128 column blocks through the operand-swapped path (n_tiles > 1):
    attention q,k projection (bias) and out-projection + identity residual segment with GN statistics."""
    B, H, C = 160, 16, 256
    g = torch.Generator().manual_seed(53)
    h = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    x = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    wqk = _bf(torch.randn(2 * C, C, generator=g) / 16).to(cuda)
    bqk = torch.randn(2 * C, generator=g).to(cuda)
    qk = ops.conv_gemm([(h, 1)], wqk, bias=bqk)
    _close(qk, (torch.einsum("bhwc,nc->bhwn", h.float(), wqk.float()) + bqk).cpu())
    wo = _bf(torch.cat([torch.randn(C, C, generator=g) / 16, torch.eye(C)], dim=1)).to(cuda)
    bo = torch.randn(C, generator=g).to(cuda)
    out = ops.conv_gemm([(h, 1), (x, 1)], wo, bias=bo, want_stats=True)
    ref = torch.einsum("bhwc,nc->bhwn", torch.cat([h, x], -1).float(), wo.float()) + bo
    _close(out, ref.cpu())
    st, nt = out.gn_stats
    tiles = out.float().reshape(B, nt, 128, C)
    assert torch.allclose(st[:, :, 0], tiles.sum(2), rtol=2e-2, atol=0.5)
    assert torch.allclose(st[:, :, 1], (tiles ** 2).sum(2), rtol=2e-2, atol=0.5)


@pytest.mark.parametrize("B,H,Cin,Cout", [(3, 32, 3, 128), (40, 32, 3, 128), (5, 16, 1, 64)])
def test_conv_in_tensor_core_form(cuda, B, H, Cin, Cout):
    """First conv as im2col_in (hi/lo split of the fp32 input, one 64-wide K-block) + 1-tap implicit GEMM, with bias and the
    GroupNorm channel sums of the output; the fp32 input keeps ~16 mantissa bits, the weights are bf16 like every other layer."""
    g = torch.Generator().manual_seed(B + H + Cin)
    x = torch.randn(B, H, H, Cin, generator=g) * 3.0
    w = torch.randn(3, 3, Cin, Cout, generator=g) / math.sqrt(9 * Cin)
    bias = torch.randn(Cout, generator=g)
    w64 = ops.conv_in_weights(w)
    assert w64.shape == (Cout, 64)
    a = ops.im2col_in(x.to(cuda))
    # the gather itself: hi + lo reproduces the fp32 neighbourhood to 2^-16 relative
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))
    nb = torch.stack([xp[:, :, kh:kh + H, kw:kw + H] for kh in range(3) for kw in range(3)], dim=1)   # [B, 9, Cin, H, H]
    nb = nb.permute(0, 3, 4, 1, 2).reshape(B, H, H, 9 * Cin)
    af = a.float().cpu()
    assert torch.allclose(af[..., :9 * Cin] + af[..., 9 * Cin:18 * Cin], nb, rtol=2e-5, atol=1e-6)
    assert (af[..., 18 * Cin:] == 0).all()
    out = ops.conv_gemm([(a, 1)], _bf(w64).to(cuda), bias=bias.to(cuda), want_stats=(H * H) % 128 == 0)
    wb = _bf(w).float()
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), wb.permute(3, 2, 0, 1).double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    _close(out, ref)
    old = ops.conv_in(x.to(cuda), w.to(cuda), bias.to(cuda))
    _close(old, ref, rtol=2e-2)


@pytest.mark.parametrize("B,H,W,Cin", [(3, 12, 12, 3), (2, 32, 32, 2), (600, 32, 32, 3), (5, 8, 64, 1)])
def test_im2col_in_gather_any_geometry(cuda, B, H, W, Cin):
    """sd_im2col_in alone: tile form (image rows staged in shared memory, eight lanes per pixel) where 128 % W == 0 and whole rows
    make a 128-pixel tile, one thread per pixel otherwise; more tiles than the grid; hi + lo reproduces the fp32 neighbourhood."""
    g = torch.Generator().manual_seed(B + H + W + Cin)
    x = torch.randn(B, H, W, Cin, generator=g) * 3.0
    a = ops.im2col_in(x.to(cuda)).float().cpu()
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))
    nb = torch.stack([xp[:, :, kh:kh + H, kw:kw + W] for kh in range(3) for kw in range(3)], dim=1)   # [B, 9, Cin, H, W]
    nb = nb.permute(0, 3, 4, 1, 2).reshape(B, H, W, 9 * Cin)
    assert torch.allclose(a[..., :9 * Cin] + a[..., 9 * Cin:18 * Cin], nb, rtol=2e-5, atol=1e-6)
    assert torch.equal(a[..., :9 * Cin], nb.bfloat16().float())
    assert (a[..., 18 * Cin:] == 0).all()


@pytest.mark.parametrize("nb,Sp,block,C,stats", [(5, 256, 256, 256, True), (150, 256, 256, 256, True), (6, 128, 64, 256, False),
                                                  (7, 128, 16, 256, False), (3, 128, 128, 128, True)])
def test_attention_core_fused(cuda, nb, Sp, block, C, stats):
    """softmax_blocks(scale q k^T) v + bias + residual in one launch (probabilities stay in shared memory) vs torch math on the
    same bf16 inputs; persistent over more tiles than SMs (nb = 150 -> 300 tiles); GroupNorm channel sums of the output."""
    g = torch.Generator().manual_seed(nb * 7 + Sp + block)
    q = _bf(torch.randn(nb, Sp, C, generator=g))
    k = _bf(torch.randn(nb, Sp, C, generator=g))
    v = _bf(torch.randn(nb, Sp, C, generator=g))
    res = _bf(torch.randn(nb, Sp, C, generator=g))
    bias = torch.randn(C, generator=g)
    scale = C ** -0.5
    vt = v.transpose(1, 2).contiguous()
    out = ops.attention_core(q.to(cuda), k.to(cuda), vt.to(cuda), scale, block=block, bias=bias.to(cuda), residual=res.to(cuda),
                             want_stats=stats)
    torch.cuda.synchronize()
    s = torch.einsum("bic,bjc->bij", q.double(), k.double()) * scale
    blk = torch.arange(Sp) // block
    s = s.masked_fill(blk[:, None] != blk[None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    ref = torch.einsum("bij,bjc->bic", _bf(p.float()).double(), v.double()) + bias.double() + res.double()
    _close(out, ref)
    if stats:
        st, nt = out.gn_stats
        tiles = out.float().reshape(nb, nt, 128, C)
        assert torch.allclose(st[:, :, 0], tiles.sum(2), rtol=2e-2, atol=0.5)
        assert torch.allclose(st[:, :, 1], (tiles ** 2).sum(2), rtol=2e-2, atol=1.0)
    else:
        assert not hasattr(out, "gn_stats")
    # strided q / k views (row stride 2C), no bias / residual
    qk = _bf(torch.randn(nb, Sp, 2 * C, generator=g)).to(cuda)
    out2 = ops.attention_core(qk[:, :, :C], qk[:, :, C:], vt.to(cuda), scale, block=block, C=C)
    s2 = torch.einsum("bic,bjc->bij", qk[:, :, :C].double().cpu(), qk[:, :, C:].double().cpu()) * scale
    s2 = s2.masked_fill(blk[:, None] != blk[None, :], float("-inf"))
    ref2 = torch.einsum("bij,bjc->bic", _bf(torch.softmax(s2, -1).float()).double(), v.double())
    _close(out2, ref2)


@pytest.mark.parametrize("B,H,C,N,taps", [(600, 8, 64, 256, 9), (601, 8, 128, 256, 1), (2400, 4, 64, 256, 9), (1200, 8, 64, 128, 1)])
def test_conv_gemm_swapped_units_spanning_images(cuda, B, H, C, N, taps):
    """Low-resolution layers at batches that give >= 148 tiles: operand-swapped 256-pixel units that span several images (8x8: four,
    4x4: sixteen) -- allowed because nothing in these epilogues is per image (bias only); 3x3 + identity-free 1-tap segment, 1x1
    layers as two 128-channel column blocks, an odd tile count; against the fp64 convolution."""
    g = torch.Generator().manual_seed(B + H + C + N + taps)
    x = _bf(torch.randn(B, H, H, C, generator=g))
    x1 = _bf(torch.randn(B, H, H, C, generator=g))
    K = taps * C
    w = _bf(torch.cat([torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, C, generator=g) / math.sqrt(C)], dim=1))
    bias = torch.randn(N, generator=g)
    out = ops.conv_gemm([(x.to(cuda), taps), (x1.to(cuda), 1)], w.to(cuda), bias=bias.to(cuda))
    torch.cuda.synchronize()
    wd = w.double()
    if taps == 9:
        ref = F.conv2d(x.double().permute(0, 3, 1, 2), wd[:, :K].reshape(N, 3, 3, C).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1)
    else:
        ref = torch.einsum("bhwc,nc->bhwn", x.double(), wd[:, :K])
    ref = ref + torch.einsum("bhwc,nc->bhwn", x1.double(), wd[:, K:]) + bias.double()
    _close(out, ref)


@pytest.mark.parametrize("B,N", [(300, 3), (37, 3), (301, 10)])
def test_conv_gemm_narrow_fp32_output(cuda, B, N):
    """The output head (3x3 conv to `num_channels` = 3 fp32 channels, cifar/models/ddpm.py:98-100) at batches with two-tile
    units and below, and a 10-channel variant: against the fp64 convolution, and the large batch against itself in small batches.
    (Padding the three weight rows to the M = 128 operand of the swapped form was measured and dropped: 202 us vs 121 us at
    batch 512 -- the zero-filled weight tiles cost the same shared-memory traffic as a full 128-channel layer.)"""
    H, C = 32, 64
    g = torch.Generator().manual_seed(B + N)
    x = _bf(torch.randn(B, H, H, C, generator=g))
    w = _bf(torch.randn(16, 9 * C, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(16, generator=g)
    out = ops.conv_gemm([(x.to(cuda), 9)], w.to(cuda), bias=bias.to(cuda), out_f32=True, n_out=N)
    torch.cuda.synchronize()
    assert out.shape == (B, H, H, N) and out.dtype == torch.float32
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w[:N].double().reshape(N, 3, 3, C).permute(0, 3, 1, 2), bias[:N].double(),
                   padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(out.cpu().double(), ref, rtol=1e-4, atol=1e-4)
    small = ops.conv_gemm([(x[:5].contiguous().to(cuda), 9)], w.to(cuda), bias=bias.to(cuda), out_f32=True, n_out=N)
    assert torch.allclose(out[:5], small, rtol=1e-5, atol=1e-5)


def test_conv_gemm_pair_two_images_per_tile(cuda):
    """8x8 level at a batch that gives 148..295 m-tiles: cta_group::2 pairs with TWO images per 128-row tile (row-bias table
    with two rows per CTA), 1-tap segments, odd pair count."""
    B, H, C = 322, 8, 64                 # 161 m-tiles -> 81 pairs, the last one half empty
    g = torch.Generator().manual_seed(61)
    a2 = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    xa = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    w = _bf(torch.cat([torch.randn(256, 9 * C, generator=g) / math.sqrt(9 * C), torch.randn(256, C, generator=g) / 8], dim=1)).to(cuda)
    bias = torch.randn(256, generator=g).to(cuda)
    rowbias = torch.randn(B, 256, generator=g).to(cuda)
    out = ops.conv_gemm([(a2, 9), (xa, 1)], w, bias=bias, rowbias=rowbias)
    wf = w.float()
    ref = F.conv2d(a2.float().permute(0, 3, 1, 2), wf[:, :9 * C].reshape(256, 3, 3, C).permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) \
        + torch.einsum("bhwc,nc->bhwn", xa.float(), wf[:, 9 * C:]) + bias + rowbias[:, None, None, :]
    _close(out, ref.cpu())


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("B,H,C,N,expect_fused", [
    (300, 16, 256, 256, True),     # pair-swapped tiles: the pair unit is the image, each CTA holds it for 128 channels
    (296, 32, 128, 128, None),     # dual-swapped tiles: four 256-pixel units per image = a cluster of 4 exchanging sums through DSMEM
    (297, 32, 128, 128, None),     # ... with a ragged last cluster round   (None: fused unless SDB_GN_FUSE < 2, see launch_gemm)
    (600, 16, 128, 128, True),     # dual-swapped, unit == image (no exchange)
    (512, 32, 256, 128, None),     # K = 2304 (the widest 32x32 conv1 of the up path takes 384 channels; 256 here)
    (8, 32, 128, 128, False),      # too few tiles for the swapped shapes: unfused fallback, raw output + stats
    (64, 8, 256, 256, True),       # thread = pixel-row tiles holding whole images: 8x8, two per 128-row tile, 128-column tiles
    (512, 8, 512, 256, True),      # ... at >= 148 tiles: cta_group::2 pairs, operands swapped, four images per unit with per-chunk statistics
    (37, 8, 256, 256, True),       # ... ragged last tile
    (512, 4, 512, 256, True),      # 4x4: eight images per tile, statistics inside 16-lane segments
    (2048, 4, 256, 256, True),     # ... in pair mode
    (9, 4, 256, 256, True),        # ... one image in the last tile
    (16, 8, 128, 128, False),      # N = 128: groups of 4 channels are not covered by the pixel-row form
])
def test_conv_gemm_with_fused_groupnorm(cuda, B, H, C, N, expect_fused, split):
    """sd_conv_gemm_gn: conv3x3 + per-sample (time-embedding) bias -> GroupNorm(32) -> swish in one launch, against the fp64
    composition and against the unfused two-launch path on the same inputs."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(B + H + C + N)
    x = torch.randn(B, H, H, C, generator=g)
    w = torch.randn(N, 9, C, generator=g) / (9 * C) ** 0.5
    rowbias = torch.randn(B, N, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(N, generator=g), 0.2 * torch.randn(N, generator=g)
    prep = ops.split_pair if split else (lambda t: t.to(torch.bfloat16))
    xs, ws = prep(x).to(cuda), prep(w.view(N, -1)).to(cuda)
    out = ops.conv_gemm([(xs, 9)], ws, rowbias=rowbias.to(cuda), want_stats=True, split=split, gn=(gamma.to(cuda), beta.to(cuda)))
    torch.cuda.synchronize()
    if expect_fused is None:
        expect_fused = os.environ.get("SDB_GN_FUSE", "2") == "2"
    expect_fused = expect_fused and not split        # the FP32-faithful arm keeps the separate (fp32) GroupNorm pass
    assert out.gn_fused == expect_fused
    a2 = out if out.gn_fused else ops.groupnorm_swish(out, gamma.to(cuda), beta.to(cuda), split=split)
    raw = ops.conv_gemm([(xs, 9)], ws, rowbias=rowbias.to(cuda), want_stats=True, split=split)
    unfused = ops.groupnorm_swish(raw, gamma.to(cuda), beta.to(cuda), split=split)
    merge = ops.merge_pair if split else (lambda t: t.float())
    # fp64 composition on the operands the kernel saw (bf16-rounded in the bf16 arm)
    xe, we = merge(xs).double().cpu(), merge(ws).double().cpu().view(N, 3, 3, C)
    h = F.conv2d(xe.permute(0, 3, 1, 2), we.permute(0, 3, 1, 2), padding=1).permute(0, 2, 3, 1) + rowbias.double()[:, None, None, :]
    hg = h.reshape(B, H * H, 32, N // 32)
    mean, var = hg.mean((1, 3), keepdim=True), hg.var((1, 3), unbiased=False, keepdim=True)
    y = ((hg - mean) / (var + 1e-6).sqrt()).reshape(B, H, H, N) * gamma.double() + beta.double()
    ref = y * torch.sigmoid(y)
    tol = 3e-5 if split else 1.2e-2        # bf16 arm: output rounding (2^-9) + the unfused path's bf16 round trip of the raw tensor
    err = ((merge(a2).double().cpu() - ref).abs().max() / ref.abs().max()).item()
    err_unfused = ((merge(unfused).double().cpu() - ref).abs().max() / ref.abs().max()).item()
    assert err <= tol, (err, err_unfused)
    if out.gn_fused and not split:
        assert err <= err_unfused * 1.05 + 1e-3      # normalising the fp32 accumulator is at least as accurate as the bf16 round trip
    # second form: the raw tensor is kept as well (it stays on the residual stream), no swish (an attention block's GroupNorm)
    both = ops.conv_gemm([(xs, 9)], ws, rowbias=rowbias.to(cuda), want_stats=True, split=split, gn=(gamma.to(cuda), beta.to(cuda), False, True))
    torch.cuda.synchronize()
    assert (both.gn_norm is not None) == expect_fused
    assert torch.equal(both, raw), "the raw second output must be the unfused conv output bit for bit"
    href = y                                          # GroupNorm without swish
    hn = both.gn_norm if both.gn_norm is not None else ops.groupnorm_swish(both, gamma.to(cuda), beta.to(cuda), swish=False, split=split)
    assert ((merge(hn).double().cpu() - href).abs().max() / href.abs().max()).item() <= tol
    if hasattr(both, "gn_stats") and hasattr(raw, "gn_stats"):
        sb, sr = both.gn_stats[0].double().sum(1), raw.gn_stats[0].double().sum(1)       # per image: [B, 2, N]
        assert torch.allclose(sb, sr, rtol=1e-5, atol=1e-2)
        a_raw = ops.groupnorm_swish(both, gamma.to(cuda), beta.to(cuda), split=split)      # a later GroupNorm over the kept raw tensor
        assert (merge(a_raw).float() - merge(unfused).float()).abs().max().item() <= 2e-2 * merge(unfused).float().abs().max().item()


@pytest.mark.parametrize("B,H,C,N", [(512, 16, 256, 256), (512, 8, 256, 256), (512, 4, 256, 256), (300, 16, 64, 128), (333, 8, 128, 64)])
def test_upconv_gemm_large_batches_match_small_batches(cuda, B, H, C, N):
    """The four phases of the fused upsample + conv run as ONE launch (phase = slowest digit of the tile index); at large batches
    the tile modes change (cta_group::2 pairs, two-tile units).  Checked against the same op on sub-batches of 5, whose small-batch
    path is pinned to the fp64 reference above; statistics against the output itself."""
    g = torch.Generator().manual_seed(B + H + C + N)
    x = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
    k = torch.randn(3, 3, C, N, generator=g) / math.sqrt(9 * C)
    bias = torch.randn(N, generator=g).to(cuda)
    w4 = ops.upconv_weights(k).bfloat16().to(cuda)
    out = ops.upconv_gemm(x, w4, bias=bias, want_stats=True)
    torch.cuda.synchronize()
    for i in (0, 5, B - 5):
        sub = ops.upconv_gemm(x[i:i + 5].contiguous(), w4, bias=bias)
        assert (out[i:i + 5].float() - sub.float()).abs().max().item() <= 1e-2 * sub.float().abs().max().item(), i
    if (H * H) % 128 == 0:
        st, n = out.gn_stats
        assert n == 4 * H * H // 128
        assert torch.allclose(st[:, :, 0].sum(1), out.float().sum(dim=(1, 2)), rtol=2e-2, atol=1.0)
        assert torch.allclose(st[:, :, 1].sum(1), (out.float() ** 2).sum(dim=(1, 2)), rtol=2e-2, atol=1.0)


def test_fused_groupnorm_is_deterministic_under_repetition(cuda):
    """The cluster exchange of the fused GroupNorm epilogue (DSMEM reads ordered by cluster-scope mbarriers, double-buffered by
    tile parity) and its register-parked pass: 30 repetitions of a multi-round launch give the same bits -- a missing barrier or a
    buffer reused too early shows up as run-to-run differences long before it shows up in a tolerance."""
    g = torch.Generator().manual_seed(7)
    outs = {}
    for (B, H, C, N) in ((592, 32, 128, 128), (700, 16, 256, 256)):
        x = _bf(torch.randn(B, H, H, C, generator=g)).to(cuda)
        w = _bf(torch.randn(N, 9 * C, generator=g) / math.sqrt(9 * C)).to(cuda)
        rb = torch.randn(B, N, generator=g).to(cuda)
        gam, bet = (1 + 0.2 * torch.randn(N, generator=g)).to(cuda), (0.1 * torch.randn(N, generator=g)).to(cuda)
        first = None
        for rep in range(30):
            out = ops.conv_gemm([(x, 9)], w, rowbias=rb, want_stats=True, gn=(gam, bet))
            assert out.gn_fused
            if first is None:
                first = out.clone()
            else:
                assert torch.equal(out, first), (B, H, rep)
        assert torch.isfinite(first.float()).all()


@pytest.mark.parametrize("env,select", [
    ({"SDB_GN_GX": "0"}, "test_conv_gemm_with_fused_groupnorm and (296-32 or 297-32) and False"),     # cluster-of-4 + DSMEM exchange
    ({"SDB_GN_GX": "0"}, "test_fused_groupnorm_is_deterministic_under_repetition"),
    ({"SDB_ATTN_V1": "1"}, "test_attention_core_fused"),                                              # the round-1 attention kernel
    ({"SDB_GEMM_PAIR128": "0"}, "test_conv_gemm_with_fused_groupnorm and 512-4-512 and False"),       # single-CTA 128-column tiles at 4x4
    ({"SDB_GEMM_SWAP_UP": "0", "SDB_GEMM_SLAB_UP": "0", "SDB_GEMM_SWAP_S2": "0", "SDB_GEMM_SWAP_MULTI": "0"},
     "test_upconv_gemm_large_batches or test_conv_gemm_stride2_tma or test_conv_gemm_swapped_units_spanning_images"),   # thread = pixel-row forms
    ({"SDB_GN_MULTI_SWAP_OFF": "1"}, "test_conv_gemm_with_fused_groupnorm and (512-8-512 or 64-8-256 or 37-8-256) and False"),   # 8x8 fused GroupNorm, pixel-row form
])
def test_fallback_kernels_behind_tuning_knobs(cuda, env, select):
    """The launch shapes that are no longer the default stay reachable (more than 8 streams asking for the global-memory
    GroupNorm exchange fall back to clusters; C != 256 attention takes the v1 kernel): the same parity tests in a child process
    with the knob set (the knobs are read once per process)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k", f"({select}) and not fallback",
                        "-p", "no:cacheprovider"], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, r.stdout[-500:]


def test_device_global_buffers_are_per_device(cuda):
    """The GroupNorm exchange area and the attention identity tiles are `__device__` variables (one instance per device): one
    process driving two GPUs must get the same bits on both.  Skipped on single-GPU boxes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    g = torch.Generator().manual_seed(5)
    B = 160
    x = _bf(torch.randn(B, 32, 32, 128, generator=g))
    w = _bf(torch.randn(128, 9 * 128, generator=g) / math.sqrt(9 * 128))
    rb = torch.randn(B, 128, generator=g)
    gam, bet = 1 + 0.2 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    q = _bf(torch.randn(8, 256, 256, generator=g))
    k = _bf(torch.randn(8, 256, 256, generator=g))
    vt = _bf(torch.randn(8, 256, 256, generator=g))
    res = _bf(torch.randn(8, 256, 256, generator=g))
    outs = []
    for d in (1, 0):                       # the second device first: a cached device-0 address would fault or miscompute here
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            o = ops.conv_gemm([(x.to(dev), 9)], w.to(dev), rowbias=rb.to(dev), want_stats=True, gn=(gam.to(dev), bet.to(dev)))
            a = ops.attention_core(q.to(dev), k.to(dev), vt.to(dev), 256 ** -0.5, block=256, residual=res.to(dev), C=256)
            torch.cuda.synchronize(dev)
            assert o.gn_fused
            outs.append((o.cpu(), a.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
