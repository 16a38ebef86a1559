"""GPU parity of the FP32-faithful arm (SD_PRECISION_FP32_FAITHFUL / SD_GEMM_SPLIT3: hi|lo bf16 operand pairs,
hi*hi + lo*hi + hi*lo on the tensor cores, fp32 accumulation, exact swish / softmax).

north_star: "final samples and log-density trajectories within rel 1e-3 in fp32".  The gate here is rel 1e-3 on the
score-net output against the reference's own ddpm.py output (tests/golden/ref_scorenet.npz) and the fp64 oracle; what the
arm actually achieves (2.3e-5 .. 2.7e-5) is asserted at 4e-5 (1.5x measured)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import scorenet as OS
from super_diffusion_b200 import ops
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils

pytestmark = pytest.mark.gpu

GATE = 1e-3        # the north_star's fp32 tolerance
ACHIEVED = 4e-5    # regression bar for this arm: 1.5x the worst measured rel-L2 (2.3e-5 .. 2.7e-5; max error 2.5e-5 .. 3.6e-5 of the
                   # output range; the fp32 CPU oracle itself sits at 3e-6 of fp64) -- profiles/r02_precision_report.json


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def _maxrel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_split_pair_roundtrip(cuda):
    x = torch.randn(1000, 64, generator=torch.Generator().manual_seed(0)) * 3
    assert _maxrel(ops.merge_pair(ops.split_pair(x)), x) < 2 ** -16


@pytest.mark.parametrize("B,H,C,N,taps", [(4, 32, 128, 128, 9), (16, 16, 256, 256, 9), (512, 4, 256, 256, 9), (8, 8, 512, 256, 9),
                                          (4, 32, 128, 128, 1), (300, 16, 256, 256, 9), (296, 32, 128, 128, 9), (5, 32, 64, 48, 9)])
def test_conv_gemm_split_matches_fp64(cuda, B, H, C, N, taps):
    """Every tile mode of the implicit GEMM (dual / swapped / pair / pair-swapped / slab / plain) with split operands."""
    g = torch.Generator().manual_seed(B + H + C)
    x = torch.randn(B, H, H, C, generator=g)
    w = torch.randn(N, taps, C, generator=g) / (taps * C) ** 0.5        # [N, tap, c]
    bias = torch.randn(N, generator=g)
    k = 3 if taps == 9 else 1
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().view(N, k, k, C).permute(0, 3, 1, 2), bias.double(),
                   padding=k // 2).permute(0, 2, 3, 1)
    out = ops.conv_gemm([(ops.split_pair(x).to(cuda), taps)], ops.split_pair(w.view(N, -1)).to(cuda), bias=bias.to(cuda),
                        want_stats=True, split=True)
    torch.cuda.synchronize()
    assert out.shape == (B, H, H, 2 * N)
    got = ops.merge_pair(out)
    assert _maxrel(got, ref) < 4e-5, _maxrel(got, ref)   # 2^-17 output pair + three dropped-term products
    if hasattr(out, "gn_stats"):
        st, n = out.gn_stats
        tot = st.double().sum(1).cpu()                                   # [B, 2, N]
        assert _maxrel(tot[:, 0], ref.reshape(B, -1, N).sum(1)) < 1e-4
        assert _maxrel(tot[:, 1], (ref ** 2).reshape(B, -1, N).sum(1)) < 1e-4
    # fp32 output + a 1-tap second source (NIN shortcut shape)
    x2 = torch.randn(B, H, H, 64, generator=g)
    w2 = torch.randn(N, 64, generator=g) / 8
    ref2 = ref + torch.einsum("bhwc,nc->bhwn", x2.double(), w2.double())
    wcat = torch.cat([w.view(N, -1), w2], 1)
    o2 = ops.conv_gemm([(ops.split_pair(x).to(cuda), taps), (ops.split_pair(x2).to(cuda), 1)], ops.split_pair(wcat).to(cuda),
                       bias=bias.to(cuda), out_f32=True, split=True)
    assert _maxrel(o2, ref2) < 2e-5, _maxrel(o2, ref2)


def test_strided_and_upsampled_conv_split(cuda):
    g = torch.Generator().manual_seed(5)
    B, H, C = 6, 32, 128
    x = torch.randn(B, H, H, C, generator=g)
    w = torch.randn(C, 3, 3, C, generator=g) / (9 * C) ** 0.5            # [N, kh, kw, c]
    bias = torch.randn(C, generator=g)
    xs = ops.split_pair(x).to(cuda)
    # stride-2 SAME conv, flax padding (0, 1)
    ref = F.conv2d(F.pad(x.double().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.double().permute(0, 3, 1, 2), bias.double(),
                   stride=2).permute(0, 2, 3, 1)
    out = ops.conv_gemm_s2(xs, ops.split_pair(w.reshape(C, -1)).to(cuda), bias=bias.to(cuda), want_stats=True, split=True)
    assert _maxrel(ops.merge_pair(out), ref) < 2e-5
    # nearest x2 upsample + 3x3 conv
    B, H, C = 6, 16, 256
    x = torch.randn(B, H, H, C, generator=g)
    k = torch.randn(3, 3, C, C, generator=g) / (9 * C) ** 0.5            # HWIO
    bias = torch.randn(C, generator=g)
    up = x.double().permute(0, 3, 1, 2).repeat_interleave(2, 2).repeat_interleave(2, 3)
    ref = F.conv2d(up, k.double().permute(3, 2, 0, 1), bias.double(), padding=1).permute(0, 2, 3, 1)
    out = ops.upconv_gemm(ops.split_pair(x).to(cuda), ops.split_pair(ops.upconv_weights(k)).to(cuda), bias=bias.to(cuda),
                          want_stats=True, split=True)
    assert out.shape == (B, 2 * H, 2 * H, 2 * C)
    assert _maxrel(ops.merge_pair(out), ref) < 2e-5, _maxrel(ops.merge_pair(out), ref)


def test_batched_gemm_split_and_softmax(cuda):
    g = torch.Generator().manual_seed(7)
    nb, S, C = 6, 256, 256
    q, k = torch.randn(nb, S, C, generator=g), torch.randn(nb, S, C, generator=g)
    sc = ops.batched_gemm(ops.split_pair(q).to(cuda), ops.split_pair(k).to(cuda), out_f32=True, split=True)
    ref = q.double() @ k.double().transpose(1, 2)
    assert _maxrel(sc, ref) < 2e-5
    for block in (256, 64):
        p = ops.softmax_rows_split(sc, C ** -0.5, block=block)
        mask = torch.block_diag(*[torch.ones(block, block)] * (S // block)).bool()
        refp = torch.softmax((ref * C ** -0.5).masked_fill(~mask, -float("inf")), -1)
        assert _maxrel(ops.merge_pair(p), refp) < 2e-5
    # P V + bias + residual with split everything, channel sums for the next GroupNorm
    vt, res, bias = torch.randn(nb, C, S, generator=g), torch.randn(nb, S, C, generator=g), torch.randn(C, generator=g)
    pm = torch.softmax(ref * C ** -0.5, -1)
    out = ops.batched_gemm(ops.split_pair(pm.float()).to(cuda), ops.split_pair(vt).to(cuda), bias=bias.to(cuda),
                           residual=ops.split_pair(res).to(cuda), want_stats=True, split=True)
    refo = pm.float().double() @ vt.double().transpose(1, 2) + bias.double() + res.double()
    assert _maxrel(ops.merge_pair(out), refo) < 2e-5
    assert hasattr(out, "gn_stats")


@pytest.mark.parametrize("B,H,C0,C1,swish", [(3, 32, 128, 0, True), (5, 16, 256, 128, True), (9, 8, 256, 256, True),
                                             (8, 4, 256, 0, False), (2, 32, 256, 128, True)])
def test_groupnorm_swish_split(cuda, B, H, C0, C1, swish):
    g = torch.Generator().manual_seed(B * H)
    C = C0 + C1
    x0 = torch.randn(B, H, H, C0, generator=g) * 2 + 0.5
    x1 = torch.randn(B, H, H, C1, generator=g) if C1 else None
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xx = torch.cat([x0, x1], -1).double() if C1 else x0.double()
    xg = xx.reshape(B, H * H, 32, C // 32)
    mean, var = xg.mean((1, 3), keepdim=True), xg.var((1, 3), unbiased=False, keepdim=True)
    y = ((xg - mean) / (var + 1e-6).sqrt()).reshape(B, H, H, C) * gamma.double() + beta.double()
    ref = y * torch.sigmoid(y) if swish else y
    out = ops.groupnorm_swish(ops.split_pair(x0).to(cuda), gamma.to(cuda), beta.to(cuda),
                              x1=ops.split_pair(x1).to(cuda) if C1 else None, swish=swish, split=True)
    assert out.shape == (B, H, H, 2 * C)
    assert _maxrel(ops.merge_pair(out), ref) < 2e-5, _maxrel(ops.merge_pair(out), ref)


def _setup(conditioned, seed):
    cfg = vpsde.get_config(conditioned=conditioned)
    model, params = mutils.init_model(seed, cfg, zero_init_scale=1.0)
    return cfg, model, mutils.perturb_params(params, torch.Generator().manual_seed(seed + 100))


@pytest.mark.parametrize("conditioned,B,t", [(False, 8, 0.73), (True, 16, 0.05), (False, 3, 1.0)])
def test_forward_fp32_faithful_matches_fp64_oracle(cuda, conditioned, B, t):
    cfg, model, params = _setup(conditioned, seed=3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 32, 32, 3, generator=g)
    y = torch.randint(0, 10, (B,), generator=g) if conditioned else None
    p64 = OS.params_to(params, dtype=torch.float64)
    with torch.no_grad():
        ref = OS.scorenet_apply(p64, cfg, torch.full((B, 1, 1, 1), t, dtype=torch.float64), x.double(), y)
    net = model.bind(params, cuda, precision="fp32")
    out = net(torch.full((B,), t), x.to(cuda), y.to(cuda) if y is not None else None)
    torch.cuda.synchronize()
    rel, mx = _rel(out, ref), _maxrel(out, ref)
    assert rel < GATE and mx < GATE, (rel, mx)
    assert rel < ACHIEVED and mx < 1.5 * ACHIEVED, (rel, mx)
    # the bf16 arm on the same inputs, for the stated-separately deviation
    bf = model.bind(params, cuda)(torch.full((B,), t), x.to(cuda), y.to(cuda) if y is not None else None)
    assert 5e-4 < _rel(bf, ref) < 2e-2


def test_forward_fp32_faithful_matches_reference_vectors(cuda):
    """Against the output of the reference's own cifar/models/ddpm.py (tests/golden/ref_scorenet.npz)."""
    from test_reference_vectors import _load, _our_params, _t
    for name, c in _load("ref_scorenet.npz").items():
        config, model, params = _our_params(c)
        net = model.bind(params, cuda, precision="fp32")
        out = net(_t(c["t"]).to(cuda), _t(c["x"]).to(cuda).contiguous(), _t(c["y"]).to(cuda))
        torch.cuda.synchronize()
        got, ref = out.double().cpu().numpy(), c["out"]
        rms = float(np.sqrt(((got - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))
        assert rms <= GATE, (name, rms)
        assert np.abs(got - ref).max() <= GATE * np.abs(ref).max(), (name, np.abs(got - ref).max() / np.abs(ref).max())
        assert rms <= ACHIEVED, (name, rms)


def test_forward_fp32_faithful_schedule_table_and_batch_independence(cuda):
    from super_diffusion_b200 import sde
    cfg, model, params = _setup(False, seed=5)
    net = model.bind(params, cuda, precision="fp32")
    x = torch.randn(11, 32, 32, 3, generator=torch.Generator().manual_seed(2)).to(cuda)
    full = net(0.31, x)
    part = net(0.31, x[3:8].contiguous())
    assert _maxrel(full[3:8], part) < 1e-4
    table = sde.schedule_table([0.9, 0.31], 1e-3, cuda)
    counter = torch.ones(1, dtype=torch.int32, device=cuda)
    viat = net(None, x, sched=table, step_counter=counter)
    assert _maxrel(viat, full) < 1e-4
