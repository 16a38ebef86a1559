"""Loop-level GPU parity: the public sampling entry points against free-running CPU oracle loops
on identical inputs and noise (north_star: samples and log-density trajectories within rel 1e-3
in fp32, kappa / mixing weights within 1e-4)."""
import math
import os

import pytest
import torch

from oracle import schedule as S
from oracle import scorenet as OS
from oracle import steps as O
from oracle import toy
from super_diffusion_b200 import dynamics, eval_utils, ops, sde
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils
from super_diffusion_b200.superposition import SuperDiffSampler, sd_superdiff, superdiff_and, superdiff_or

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / (1e-12 + b.double().abs().max())).item()


@pytest.mark.parametrize("mode", ["or", "and"])
def test_toy_config1_trajectories(cuda, mode):
    """BASELINE config 1 (2-D mixture toy, two score models, 1000 steps, float32 time accumulation) at batch 4096.

    Free-running fp32 (GPU) vs free-running fp64 (oracle): samples and log-density trajectories within rel 1e-3.
    kappa is a function of the accumulated fp32 state, so along a free-running 1000-step trajectory it is held to
    2e-3 (OR; the fp32 accumulation noise of ll ~ 50 alone moves softmax(ll) by ~1e-3) and, for AND (unclipped,
    superposition_edu.ipynb:904), to 1e-2 relative.  The 1e-4 kappa gate on
    *identical inputs* is the teacher-forced part below (and tests/test_step_gpu.py)."""
    B, n, dt = 4096, 1000, 1e-3
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 2, generator=g)
    noise = torch.randn(n, B, 2, generator=torch.Generator().manual_seed(3))
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    xr, llr, tr = toy.loop_toy(fns, x0.double(), noise, mode, n, dt, record=True)
    run = superdiff_or if mode == "or" else superdiff_and
    x, ll, w, traj = run(fns, x0.to(cuda), n_steps=n, dt=dt, noise=noise.to(cuda), record=True)
    torch.cuda.synchronize()
    kap = traj["kappa"][:, :, 0].cpu().double()
    kr = tr["kappa"]
    tx, tl = traj["x"].cpu().double(), traj["ll"].cpu().double()
    assert _rel(tx, tr["x"]) <= 1e-3 and _rel(tl, tr["ll"]) <= 1e-3
    assert torch.isfinite(tx).all() and torch.isfinite(tl).all()
    # calibration (CPU, same seeds, B=1024): the literal fp32 transcription of the notebook deviates from the fp64
    # truth by 7.9e-4 (OR kappa) and 4.2e-3 relative (AND kappa, unclipped, |kappa| up to 1.3e3 near t = 1)
    if mode == "or":
        assert (kap - kr).abs().max().item() <= 2e-3
    else:
        assert ((kap - kr).abs() / (1 + kr.abs())).max().item() <= 1e-2
    # teacher-forced: identical fp32 inputs (oracle state and scores rounded to fp32) -> kappa within 1e-4
    ts = S.time_grid(n, dt, "float32")
    for i in range(0, n, 97):
        t = float(ts[i])
        xi = tr["x"][i].float()
        lli = tr["ll"][i].float()
        tt = torch.full((B, 1), t)
        sc = torch.stack([f(tt, xi) for f in fns])
        a32, b32 = float(torch.tensor(S.dlog_alphadt(t)).float()), float(torch.tensor(S.beta(t)).float())
        t32, dt32 = float(torch.tensor(t).float()), float(torch.tensor(dt).float())
        md = O.MODE_OR if mode == "or" else O.MODE_AND
        xr1, llr1, wr1 = O.step_vpsde_gram(xi, noise[i], sc, lli, a32, b32, t32, dt32, md, O.DLOGQ_ITO,
                                           ito_const=4 * dt32 * a32)
        xo, lo, wo = ops.step_vpsde(xi.to(cuda), noise[i].to(cuda), [s_.contiguous() for s_ in sc.to(cuda)], lli.to(cuda).clone(),
                                    S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt, md, ops.DLOGQ_ITO, ito_scale=4.0)
        wellc = wr1.abs().max(dim=1).values < 1e4
        assert ((wo.cpu().double() - wr1).abs() / (1 + wr1.abs()))[wellc].max().item() <= 1e-4
        assert torch.allclose(lo.cpu().double()[wellc], llr1[wellc], rtol=1e-5, atol=1e-4 * (1 + llr1[wellc].abs().max().item()))


@pytest.mark.parametrize("mode", ["or", "and"])
def test_toy_config1_with_the_notebook_mlp(cuda, mode):
    """BASELINE config 1 as the reference runs it: two MLP score models (superposition_edu.ipynb:157-173, 3 -> 512 x 4 -> 2,
    swish) feeding the SuperDiff OR / AND loops.  The product ships the MLP as a PyTorch module (models/toy_mlp.py); here two
    lecun-normal-initialised instances run on the GPU in fp32 against the fp64 oracle MLP on the same parameters, free-running
    for 1000 steps (float32 time accumulation like the notebook)."""
    from super_diffusion_b200.models.toy_mlp import MLP, get_sscore
    B, n, dt = 1024, 1000, 1e-3
    x0 = torch.randn(B, 2, generator=torch.Generator().manual_seed(0))
    noise = torch.randn(n, B, 2, generator=torch.Generator().manual_seed(3))
    mlps = [MLP.init(1), MLP.init(2)]
    p64 = [OS.params_to(m.to_flax(), dtype=torch.float64) for m in mlps]
    with torch.no_grad():
        xr, llr, tr = toy.loop_toy([lambda t, x, p=p: OS.toy_mlp_apply(p, t, x) for p in p64], x0.double(), noise, mode, n, dt,
                                   record=True)
    fns = [get_sscore(m.to(cuda)) for m in mlps]
    run = superdiff_or if mode == "or" else superdiff_and
    x, ll, w, traj = run(fns, x0.to(cuda), n_steps=n, dt=dt, noise=noise.to(cuda), record=True)
    torch.cuda.synchronize()
    tx, tl = traj["x"].cpu().double(), traj["ll"].cpu().double()
    assert torch.isfinite(tx).all() and torch.isfinite(tl).all()
    # OR: rel 1e-3 along the whole free-running trajectory.  AND over two UNTRAINED networks is a harsher problem than the
    # notebook's (kappa is unclipped and |s_1 - s_2| is small where two random MLPs happen to agree, superposition_edu.ipynb:904):
    # fp32 rounding is amplified along 1000 free-running steps (measured 2.6e-2 / 4.0e-2), so the free run is held to 1e-1 and the
    # 1e-3 / 1e-4 gates are applied teacher-forced below, on the oracle's own states.
    tol = 1e-3 if mode == "or" else 1e-1
    assert _rel(tx, tr["x"]) <= tol and _rel(tl, tr["ll"]) <= tol, (_rel(tx, tr["x"]), _rel(tl, tr["ll"]))
    ts = S.time_grid(n, dt, "float32")
    md = O.MODE_OR if mode == "or" else O.MODE_AND
    for i in range(0, n, 97):
        t = float(ts[i])
        xi, lli = tr["x"][i].float(), tr["ll"][i].float()
        tt = torch.full((B, 1), t)
        sc = [f(tt.to(cuda), xi.to(cuda)).contiguous() for f in fns]                   # the product's MLPs, fp32 on the GPU
        sc64 = torch.stack([OS.toy_mlp_apply(p, tt.double(), xi.double()) for p in p64])
        assert _rel(torch.stack(sc).cpu(), sc64) <= 1e-5
        f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
        xr1, llr1, wr1 = O.step_vpsde_gram(xi, noise[i], torch.stack(sc).cpu(), lli, f32(S.dlog_alphadt(t)), f32(S.beta(t)), f32(t),
                                           f32(dt), md, O.DLOGQ_ITO, ito_const=4 * f32(dt) * f32(S.dlog_alphadt(t)))
        xo, lo, wo = ops.step_vpsde(xi.to(cuda), noise[i].to(cuda), sc, lli.to(cuda).clone(), S.dlog_alphadt(t), S.beta(t),
                                    S.sigma(t), dt, md, ops.DLOGQ_ITO, ito_scale=4.0)
        wellc = wr1.abs().max(dim=1).values < 1e4
        assert ((wo.cpu().double() - wr1).abs() / (1 + wr1.abs()))[wellc].max().item() <= 1e-4
        assert _rel(xo.cpu()[wellc], xr1[wellc]) <= 1e-3 and _rel(lo.cpu()[wellc], llr1[wellc]) <= 1e-3


def _gauss_sscore(mu, var):
    """Exact sigma_t * grad log q_t of N(mu, var I) data under the forward process (any dimension)."""
    def fn(t, x):
        tt = t.to(x.dtype).reshape(-1, 1)
        alpha = torch.exp(S.log_alpha(tt))
        m = torch.tensor(mu, dtype=x.dtype, device=x.device)
        return -tt * (x - alpha * m) / (alpha ** 2 * var + tt ** 2)
    return fn


def test_general_m_and_or_small_d(cuda):
    """M = 3 models in D = 16 (small-D kernel): general AND solve (not in the reference) and OR with
    temperature and bias, 200 free-running steps against the fp64 Gram oracle."""
    B, D, n, dt = 512, 16, 200, 5e-3
    x0 = torch.randn(B, D, generator=torch.Generator().manual_seed(1))
    noise = torch.randn(n, B, D, generator=torch.Generator().manual_seed(2))
    gm = torch.Generator().manual_seed(4)
    fns = [_gauss_sscore((2.0 * torch.randn(D, generator=gm)).tolist(), v) for v in (0.16, 0.5, 1.0)]
    ts = S.time_grid(n, dt, "float32")
    for mode, kw in ((O.MODE_AND, {}), (O.MODE_OR, dict(temperature=2.0, logp_bias=[0.1, 0.0, -0.1]))):
        x = x0.double().clone()
        ll = torch.zeros(B, 3, dtype=torch.float64)
        for i in range(n):
            t = float(ts[i])
            s = torch.stack([f(torch.full((B, 1), t, dtype=torch.float64), x) for f in fns])
            x, ll, wr = O.step_vpsde_gram(x, noise[i], s, ll, S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt, mode, O.DLOGQ_ITO,
                                          ito_const=D * D * dt * S.dlog_alphadt(t), **kw)
        if mode == O.MODE_AND:
            from super_diffusion_b200.superposition import _vpsde_loop
            xg, llg, wg, _ = _vpsde_loop(fns, x0.to(cuda), ops.MODE_AND, ops.DLOGQ_ITO, n, dt, noise.to(cuda), 0, 1.0, None,
                                         float(D * D), torch.zeros(B, 3), "float32", False)
        else:
            xg, llg, wg, _ = superdiff_or(fns, x0.to(cuda), n_steps=n, dt=dt, noise=noise.to(cuda), temperature=2.0,
                                          logp_bias=[0.1, 0.0, -0.1])
        assert _rel(xg.cpu(), x) <= 1e-3 and _rel(llg.cpu(), ll) <= 1e-3, mode
        assert ((wg.cpu().double() - wr).abs() / (1 + wr.abs())).max().item() <= 2e-3, mode


def _two_models(cuda, seeds=(10, 11)):
    cfg = vpsde.get_config()
    models, states, params = [], [], []
    for s in seeds:
        model, p = mutils.init_model(s, cfg, zero_init_scale=1.0)
        models.append(model); params.append(p)
        states.append(mutils.State(params_ema=p, model_params=p))
    return cfg, models, states, params


def test_cifar_sampler_graph_equals_eager_and_generator(cuda):
    """The CUDA-graph sampler, its eager twin and the reference-shaped get_generator / joint_vf closures
    produce the same samples, log-densities and weights (same kernels, same noise)."""
    cfg, models, states, _ = _two_models(cuda)
    B, n = 8, 5
    nets = [m.bind(s.params_ema, cuda) for m, s in zip(models, states)]
    g = torch.Generator(device=cuda).manual_seed(5)
    x0 = torch.randn(B, 32, 32, 3, generator=g, device=cuda)
    noise = torch.randn(n, B, 32, 32, 3, generator=g, device=cuda)
    out = {}
    for use_graph in (True, False):
        smp = SuperDiffSampler(nets, B, mode="or", n_steps=n, dt=5e-3, temperature=1e6, device=cuda, use_graph=use_graph)
        x, lq, w = smp.sample(x0=x0, noise=noise)
        out[use_graph] = (x.clone(), lq.clone(), w.clone())
        assert smp.launches_per_step > 150          # ~240 at this batch: every op is its own kernel launch (no library fallback)
    for a, b in zip(out[True], out[False]):
        assert torch.equal(a, b)
    # closure API (cifar/dynamics.py:115 signature): increments, caller adds them
    vf = dynamics.get_joint_stoch_vf(0, models, states)
    x, logq, t, dt = x0.clone(), torch.zeros(B, 2, device=cuda), 1.0, 5e-3
    for i in range(n):
        dx, dlogq = vf(t, (x, logq), {"key": 0, "labels": None, "dt": dt, "noise": noise[i]})
        x = x + dx; logq = logq + dlogq; t += -dt
    assert torch.allclose(x, out[True][0], rtol=1e-5, atol=1e-5)
    assert torch.allclose(logq, out[True][1], rtol=1e-4, atol=1e-3)
    # get_generator: reference loop shape (x0 ~ N(0,I) from the key, logq0 = 0, n = int(1/dt))
    cfg.eval.batch_size = 4
    gen = eval_utils.get_generator(models, cfg, vf, dt=0.25, device=cuda, return_logq=True)
    xa, na, lqa = gen(7, None)
    xb, nb, lqb = gen(7, None)
    assert na == 4 and xa.shape == (4, 32, 32, 3) and torch.equal(xa, xb) and torch.equal(lqa, lqb)
    assert (lqa.max(dim=1).values == 0).all()
    # the generator ran as a replayed CUDA graph (vector_field.sampler_spec); the per-launch path gives the same bits
    assert getattr(vf, "sampler_spec", None) is not None
    xe, ne, lqe = gen(7, None, eager=True)
    assert torch.equal(xa, xe) and torch.equal(lqa, lqe)
    # ... also with caller-supplied noise and a log-density trace (pinned host memory, one D2H per step)
    nz = torch.randn(4, 4, 32, 32, 3, generator=torch.Generator().manual_seed(3)).pin_memory()
    tr_g, tr_e = torch.empty(4, 4, 2).pin_memory(), torch.empty(4, 4, 2).pin_memory()
    xg2, _, lqg2 = gen(7, None, noise=nz, logq_trace=tr_g)
    xe2, _, lqe2 = gen(7, None, noise=nz, logq_trace=tr_e, eager=True)
    torch.cuda.synchronize()
    assert torch.equal(xg2, xe2) and torch.equal(lqg2, lqe2) and torch.equal(tr_g, tr_e) and torch.equal(tr_g[-1], lqg2.cpu())


def test_sampler_lazy_capture_keeps_state_and_guards_the_schedule(cuda):
    """step() without an explicit capture(): the warm-up inside capture() must not advance the state (the first timestep was
    applied twice before), and stepping past the last schedule row raises instead of reading beyond the table."""
    cfg, models, states, _ = _two_models(cuda)
    B, n = 8, 3
    nets = [m.bind(s.params_ema, cuda) for m, s in zip(models, states)]
    g = torch.Generator(device=cuda).manual_seed(11)
    x0 = torch.randn(B, 32, 32, 3, generator=g, device=cuda)
    noise = torch.randn(n, B, 32, 32, 3, generator=g, device=cuda)
    ref = SuperDiffSampler(nets, B, mode="or", n_steps=n, dt=5e-3, device=cuda)
    ref.capture()
    ref.reset(x0)
    lazy = SuperDiffSampler(nets, B, mode="or", n_steps=n, dt=5e-3, device=cuda)
    lazy.reset(x0)                       # state set BEFORE the lazy capture
    for i in range(n):
        ref.step(noise[i]); lazy.step(noise[i])
    torch.cuda.synchronize()
    assert torch.equal(ref.x, lazy.x) and torch.equal(ref.logq, lazy.logq)
    assert int(lazy.counter.item()) == n - 1          # saturating device counter: never past the last row
    with pytest.raises(RuntimeError, match="reset"):
        lazy.step(noise[0])
    lazy.reset(x0)
    lazy.step(noise[0])


def test_cifar_or_steps_against_cpu_oracle(cuda):
    """Three free-running SuperDiff-OR steps: B200 path (bf16 score-net GEMMs) vs the CPU oracle (fp32 score-net,
    literal cifar/dynamics.py:123-136).  bf16-denoiser tolerance, stated separately from the fp32 1e-3 gate:
    samples rel 3e-3 (dx is O(dt) so denoiser error enters scaled by dt*2b), log-densities rel 3e-2."""
    cfg, models, states, params = _two_models(cuda)
    B, n, dt = 4, 3, 5e-3
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(B, 32, 32, 3, generator=g)
    noise = torch.randn(n, B, 32, 32, 3, generator=g)
    x, logq, t = x0.clone(), torch.zeros(B, 2), 1.0
    with torch.no_grad():
        for i in range(n):
            tt = torch.full((B, 1, 1, 1), t)
            s = torch.stack([OS.scorenet_apply(p, cfg, tt, x, None) for p in params])
            dx, dlogq, w_ref = O.or_step_cifar_literal(x, logq, s, noise[i], t, dt)
            x = x + dx; logq = logq + dlogq; t += -dt
    nets = [m.bind(s_.params_ema, cuda) for m, s_ in zip(models, states)]
    smp = SuperDiffSampler(nets, B, mode="or", n_steps=200, dt=dt, temperature=1e6, device=cuda)
    smp.capture()
    smp.reset(x0.to(cuda))
    for i in range(n):
        smp.step(noise[i].to(cuda))
    torch.cuda.synchronize()
    assert _rel(smp.x.cpu(), x) <= 3e-3
    gap = (logq[:, 0] - logq[:, 1]).abs()
    assert _rel(smp.logq.cpu(), logq) <= 3e-2
    clear = gap > 0.1 * gap.max()
    assert torch.equal(smp.weights.cpu()[clear].argmax(1), w_ref[clear].argmax(1))


def test_cifar_or_200_steps_both_arms_against_the_fp64_oracle(cuda):
    """The reference-default loop (cifar/eval_utils.py:75-77: 200 Euler-Maruyama steps, dt = 5e-3) with the REAL U-Net -- two
    random-init score-nets, batch 64, SuperDiff-OR at the reference's T = 1e6 -- on both precision arms against the fp64 CPU
    oracle trajectory committed as tests/golden/cifar_or_200step_oracle_float64.npz (tools/deviation_study.py --oracle; identical
    x0 and noise from seeded CPU generators).  north_star: "final samples and log-density trajectories within rel 1e-3 in fp32
    (stated separately for bf16 denoiser GEMMs)".  Measured (profiles/r02_deviation.md):
        FP32-faithful arm: final samples 9.0e-6 (median) / 1.0e-5 (max) rel L2, log-density gap 1.1e-5 median / 2.5e-5 max
        bf16 arm:          final samples 4.1e-3 / 4.4e-3,                         log-density gap 1.3e-3 median / 1.3e-2 max (step 1)
    every sample follows the oracle's OR-winner sequence on both arms.  Gates: 1e-3 for the fp32 arm (the north_star's), and
    <= 1.5x the measured values for the bf16 arm."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import deviation_study as dev
    ref = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cifar_or_200step_oracle_float64.npz")))
    B, n = int(ref["B"]), int(ref["n"])
    x0, noise = dev.inputs(B, n)
    cfg, mods, params = dev.models()
    gates = {"fp32": dict(x=1e-3, gap_med=1e-3, gap_max=1e-3), "bf16": dict(x=6.6e-3, gap_med=2e-3, gap_max=1.9e-2)}
    for prec, gate in gates.items():
        nets = [m.bind(p, cuda, precision=prec) for m, p in zip(mods, params)]
        smp = SuperDiffSampler(nets, B, mode="or", n_steps=n, dt=1.0 / n, temperature=1e6, device=cuda)
        smp.capture()
        smp.reset(x0.to(cuda))
        lq, w = [], []
        for i in range(n):
            smp.step(noise[i].to(cuda))
            lq.append(smp.logq.double().cpu().numpy()); w.append(smp.weights.double().cpu().numpy())
        torch.cuda.synchronize()
        row = dev.compare(prec, np.stack(lq), np.stack(w), smp.x.double().cpu().numpy(), ref)
        assert row["winner_sequence_match_frac"] == 1.0, row
        assert row["final_sample_rel_l2_max"] <= gate["x"], row
        assert row["logq_gap_rel_err_median"] <= gate["gap_med"] and row["logq_gap_rel_err_max"] <= gate["gap_max"], row
        del smp, nets


def test_sd_latent_loop_against_cpu_oracle(cuda):
    """SD-style loop (clip_eval.py:348-415) with a small caller-supplied velocity network standing in for the UNet."""
    torch.manual_seed(0)
    B, C, H = 3, 4, 16
    N = 12
    ws = {k: 0.3 * torch.randn(C, C, 3, 3) for k in ("obj", "bg", "uncond")}

    def make_vel(dtype, device):
        wd = {k: v.to(dtype=dtype, device=device) for k, v in ws.items()}

        def get_vel(t, sigma, latents, which):
            h = latents / ((sigma ** 2 + 1) ** 0.5)                       # clip_eval.py:92
            return torch.tanh(torch.nn.functional.conv2d(h, wd[which], padding=1)) * (1.0 + 0.001 * t)
        return get_vel

    lat0 = torch.randn(B, C, H, H, generator=torch.Generator().manual_seed(1))
    z = torch.randn(N, B, C, H, H, generator=torch.Generator().manual_seed(2))
    sig, ts, init = S.edm_sigmas(N)
    for method in ("and", "or", "avg"):
        gv = make_vel(torch.float64, "cpu")
        x = lat0.double() * init
        ll = torch.ones(B, 2, dtype=torch.float64)
        kr = []
        for i in range(N):
            sigma, dsigma = float(sig[i]), float(sig[i + 1] - sig[i])
            vo, vu, vb = gv(float(ts[i]), sigma, x, "obj"), gv(float(ts[i]), sigma, x, "uncond"), gv(float(ts[i]), sigma, x, "bg")
            dx, ll, kappa = O.sd_step_literal(x, z[i].double(), vo, vb, vu, ll, sigma, dsigma, method,
                                              guidance_scale=7.5, lift=0.0, num_inference_steps=N, T=1.0, logp=0.0)
            x = x + dx
            kr.append(kappa)
        xg, llg, kg, traj = sd_superdiff(make_vel(torch.float32, cuda), lat0.to(cuda), method=method,
                                         num_inference_steps=N, noise=z.to(cuda), record=True)
        assert _rel(xg.cpu(), x) <= 1e-3, method
        assert _rel(llg.cpu(), ll) <= 1e-3, method
        krs = torch.stack(kr)
        assert (traj["kappa"][1:].cpu().double() - krs).abs().max().item() <= 1e-4 * (1 + krs.abs().max().item()), method


def test_cifar_sampler_and_mode_and_three_models(cuda):
    """BASELINE config 3 shape of work at a tiny batch: SuperDiff-AND on CIFAR tensors through the graph sampler
    (log-density increments equalised across models, weights sum to 1), and an M = 3 OR run."""
    cfg, models, states, _ = _two_models(cuda)
    nets = [m.bind(s.params_ema, cuda) for m, s in zip(models, states)]
    B, n = 8, 4
    g = torch.Generator(device=cuda).manual_seed(3)
    x0 = torch.randn(B, 32, 32, 3, generator=g, device=cuda)
    noise = torch.randn(n, B, 32, 32, 3, generator=g, device=cuda)
    smp = SuperDiffSampler(nets, B, mode="and", n_steps=n, dt=5e-3, device=cuda)
    x, lq, w = smp.sample(x0=x0, noise=noise)
    assert torch.isfinite(x).all() and torch.isfinite(lq).all()
    assert torch.allclose(w.sum(1), torch.ones(B, device=cuda), atol=1e-5)
    assert (lq[:, 0] - lq[:, 1]).abs().max().item() <= 1e-3 * (1 + lq.abs().max().item())
    # pinned host noise: the pipelined path (copy stream + staging buffers) gives bit-identical results
    x_dev, lq_dev = x.clone(), lq.clone()
    x_h, lq_h, _ = smp.sample(x0=x0, noise=noise.cpu().pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(x_h, x_dev) and torch.equal(lq_h, lq_dev)
    model3, p3 = mutils.init_model(12, cfg, zero_init_scale=1.0)
    nets3 = nets + [model3.bind(p3, cuda)]
    smp3 = SuperDiffSampler(nets3, B, mode="or", n_steps=n, dt=5e-3, temperature=1e6, device=cuda)
    x3, lq3, w3 = smp3.sample(x0=x0, noise=noise)
    assert w3.shape == (B, 3) and torch.allclose(w3.sum(1), torch.ones(B, device=cuda), atol=1e-5)
    assert (lq3.max(dim=1).values == 0).all()


def test_sample_driver_writes_reference_npz_format(cuda, tmp_path):
    """run_lib.evaluate_joint_samples: samples_{i}.npz with uint8 `samples` [B,32,32,3] and `num_steps` (cifar/run_lib.py:244-251)."""
    import numpy as np
    from super_diffusion_b200 import run_lib
    cfg = vpsde.get_config()
    cfg.eval.batch_size = 4
    cfg.eval.num_samples = 8
    params = [mutils.init_model(s, cfg, zero_init_scale=1.0)[1] for s in (1, 2)]
    d = run_lib.evaluate_joint_samples(cfg, str(tmp_path), "eval", params, stoch=True, dt=0.25, device=cuda)
    files = sorted(os.listdir(d))
    assert files == ["samples_0.npz", "samples_1.npz"]
    z = np.load(os.path.join(d, files[0]))
    assert z["samples"].dtype == np.uint8 and z["samples"].shape == (4, 32, 32, 3) and int(z["num_steps"]) == 4


def test_cli_joint_eval_from_exported_checkpoints(cuda, tmp_path):
    """python -m super_diffusion_b200.main --mode eval_joint_fid_stoch --chkpts a.npz,b.msgpack (cifar/main.py:33-35 ->
    run_lib.evaluate_joint_fid) and --mode eval_fid (deterministic single-model path, run_lib.evaluate_fid :129-167)."""
    import numpy as np
    from super_diffusion_b200 import checkpoint, main as cli
    cfg = vpsde.get_config()
    pa = mutils.init_model(1, cfg, zero_init_scale=1.0)[1]
    pb = mutils.init_model(2, cfg, zero_init_scale=1.0)[1]
    checkpoint.save_npz(tmp_path / "a.npz", pa)
    (tmp_path / "b.msgpack").write_bytes(checkpoint.to_msgpack_bytes(pb))
    d = cli.launch(["--config", "vpsde", "--workdir", str(tmp_path), "--mode", "eval_joint_fid_stoch", "--chkpts",
                    f"{tmp_path / 'a.npz'}, {tmp_path / 'b.msgpack'}", "--batch_size", "4", "--num_batches", "1", "--dt", "0.25"])
    assert d.endswith(os.path.join("eval", "samples_stoch"))
    z = np.load(os.path.join(d, "samples_0.npz"))
    assert z["samples"].dtype == np.uint8 and z["samples"].shape == (4, 32, 32, 3) and int(z["num_steps"]) == 4
    checkpoint.save_npz(tmp_path / "params_ema.npz", pa)
    d2 = cli.launch(["--config", "vpsde", "--workdir", str(tmp_path), "--mode", "eval_fid", "--batch_size", "4", "--num_batches", "1",
                     "--dt", "0.25"])
    assert d2.endswith(os.path.join("eval", "samples")) and os.path.exists(os.path.join(d2, "samples_0.npz"))


def test_cli_deterministic_joint_eval_at_a_batch_that_does_not_fill_attention_tiles(cuda, tmp_path):
    """--mode eval_joint_fid (deterministic SuperDiff-OR: get_joint_vf, Hutchinson divergence through the score-net JVP) with the
    reference's kind of batch -- eval.batch_size = 100 there, 12 per GPU on 8 GPUs: not a multiple of the 8 images a 4x4 attention
    tile packs.  The JVP pads the batch internally (it raised NotImplementedError before); --precision fp32 runs the stochastic
    mode on the FP32-faithful arm."""
    import numpy as np
    from super_diffusion_b200 import checkpoint, main as cli
    cfg = vpsde.get_config()
    for i, seed in enumerate((1, 2)):
        checkpoint.save_npz(tmp_path / f"m{i}.npz", mutils.init_model(seed, cfg, zero_init_scale=1.0)[1])
    chk = f"{tmp_path / 'm0.npz'},{tmp_path / 'm1.npz'}"
    d = cli.launch(["--config", "vpsde", "--workdir", str(tmp_path), "--mode", "eval_joint_fid", "--chkpts", chk, "--batch_size", "12",
                    "--num_batches", "1", "--dt", "0.5"])
    z = np.load(os.path.join(d, "samples_0.npz"))
    assert z["samples"].shape == (12, 32, 32, 3) and int(z["num_steps"]) == 2
    d = cli.launch(["--config", "vpsde", "--workdir", str(tmp_path), "--mode", "eval_joint_fid_stoch", "--chkpts", chk, "--batch_size", "12",
                    "--num_batches", "1", "--dt", "0.5", "--precision", "fp32", "--eval_folder", "eval32"])
    z = np.load(os.path.join(d, "samples_0.npz"))
    assert z["samples"].shape == (12, 32, 32, 3)
