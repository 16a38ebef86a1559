"""Parity against outputs of the REFERENCE'S OWN SOURCE executed in the build container
(tests/golden/ref_*.npz, produced by tests/golden/make_reference_vectors.py: the reference's unmodified
cifar/dynamics.py, cifar/eval_utils.py, cifar/models/*.py, notebook cells and applications/images/clip_eval.py
run over the NumPy / CPU-PyTorch API shims in tests/golden/ref_shims/).

CPU (-m "not gpu"): the oracle reproduces them (this is what pins the oracle).
GPU (-m gpu): the CUDA path, through the C ABI and the reference-named Python entry points, reproduces them
within the north_star tolerances (fp32: samples / log-densities rel 1e-3, kappa / weights 1e-4; bf16
score-net GEMMs stated separately).
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import schedule as S
from oracle import scorenet as OS
from oracle import steps as O
from oracle import toy

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(fname):
    z = np.load(os.path.join(HERE, fname))
    cases = {}
    for k in z.files:
        c, f = k.split("/")
        cases.setdefault(c, {})[f] = z[k]
    return cases


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t if dtype is None else t.to(dtype)


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (1e-300 + np.abs(b).max()))


def ref_normal(key, shape):
    """The draw the shim's jax.random.normal(key, shape) made (tests/golden/ref_shims/jax/random.py)."""
    return np.random.default_rng(int(key)).standard_normal(tuple(shape)).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# model stand-ins shared by the generator script and these tests
# ---------------------------------------------------------------------------------------------

def gauss_score(mu, t, x, y):
    """GaussModel of make_reference_vectors.py in torch: sigma_t * grad log N(alpha_t (mu + 0.1 y), alpha_t^2/4 + t^2)."""
    t = float(t)
    m = mu + 0.1 * y.to(x.dtype)[:, None, None, None]
    alpha = math.exp(-0.5 * t * 0.1 - 0.25 * t ** 2 * 19.9)
    var = alpha ** 2 * 0.25 + t ** 2
    return -t * (x - alpha * m) / var


class GaussModelTorch:
    """Model-seam stand-in on the product side: get_model_fn calls model.bind(params)(t, x, y)."""

    def bind(self, params, device=None):
        def fn(t, x, y):
            return gauss_score(params["mu"].to(x.device, x.dtype), float(torch.as_tensor(t).reshape(-1)[0]), x,
                               y.to(x.device)).contiguous()
        return fn


class TableModelTorch:
    def bind(self, params, device=None):
        return lambda t, x, y: params["table"]


def nonlin_score(mu, phi, t, x):
    """NonlinModel of make_reference_vectors.py in torch."""
    t = float(t)
    alpha = math.exp(-0.5 * t * 0.1 - 0.25 * t ** 2 * 19.9)
    var = alpha ** 2 * 0.25 + t ** 2
    return -t * (x - alpha * mu) / var + 0.3 * torch.sin(1.7 * x + phi) + 0.2 * x * torch.roll(x, 1, dims=2)


class NonlinModelTorch:
    def bind(self, params, device=None):
        return lambda t, x, y: nonlin_score(params["mu"].to(x.device, x.dtype), params["phi"], float(torch.as_tensor(t).reshape(-1)[0]), x)


def sd_unet_stub(x_in, t, phase):
    return 0.8 * x_in * math.cos(phase) + torch.sin(1.3 * x_in + phase + 1e-3 * float(t))


# ---------------------------------------------------------------------------------------------
# CPU: oracle == reference outputs
# ---------------------------------------------------------------------------------------------

def test_oracle_matches_reference_cifar_steps():
    for name, c in _load("ref_cifar_steps.npz").items():
        kind, t, dt = str(c["kind"]), float(c["t"]), float(c["dt"])
        x, eps, s, logq = (_t(c[k], torch.float64) for k in ("x", "eps", "scores", "logq"))
        if kind == "stoch":
            dx, dlogq, w = O.or_step_cifar_literal(x, logq, s, eps, t, dt)
            assert _rel(dx, c["dx"]) < 1e-13 and np.abs(dlogq.numpy() - c["dlogq"]).max() < 1e-9 * (1 + np.abs(c["dlogq"]).max()), name
            # ... and the Gram form the kernels implement
            xo, lq, _ = O.step_vpsde_gram(x, eps, s, logq, S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt, O.MODE_OR,
                                          O.DLOGQ_CIFAR_MAXSUB, temperature=1e6)
            assert _rel(xo - x, c["dx"]) < 1e-12, name
            assert np.abs((lq - logq).numpy() - c["dlogq"]).max() < 1e-8 * (1 + np.abs(c["dlogq"]).max()), name
        elif kind in ("avg1", "avg0"):
            dx, dlogq = O.avg_step_cifar_literal(x, s, eps, t, dt, stoch=(kind == "avg1"))
            assert _rel(dx, c["dx"]) < 1e-13 and np.array_equal(dlogq.numpy(), c["dlogq"]), name
        else:
            dx, dlogq = O.single_ode_step_cifar_literal(x, s[0], t, dt)
            assert _rel(dx, c["dx"]) < 1e-13 and np.array_equal(dlogq.numpy(), c["dlogq"]), name


def test_oracle_matches_reference_cifar_ode():
    """get_joint_vf (cifar/dynamics.py:59-97): scores and J eps from torch.func.jvp of the same stand-in models."""
    for name, c in _load("ref_cifar_ode.npz").items():
        t, dt = float(c["t"]), float(c["dt"])
        x, logq = _t(c["x"], torch.float64), _t(c["logq"], torch.float64)
        M = c["mus"].shape[0]
        ss, js = [], []
        for i in range(M):
            f = lambda _x, i=i: nonlin_score(_t(c["mus"][i], torch.float64), float(c["phis"][i]), t, _x)
            s_i, j_i = torch.func.jvp(f, (x,), (_t(c["probes"][i], torch.float64),))
            ss.append(s_i); js.append(j_i)
        dx, dlogq, _ = O.ode_step_cifar_literal(x, logq, torch.stack(ss), torch.stack(js), _t(c["probes"], torch.float64), t, dt)
        assert _rel(dx, c["dx"]) < 1e-12, name
        # the fixture's jax.jvp stand-in is a central difference (h = 1e-6): ~1e-9 relative on the divergence term
        assert np.abs(dlogq.numpy() - c["dlogq"]).max() < 1e-7 * (1 + np.abs(c["dlogq"]).max()), name


def _oracle_cifar_loop(c):
    mus = _t(c["mus"], torch.float64)
    y = _t(c["labels"])
    x = _t(c["x0"], torch.float64)
    M = mus.shape[0]
    logq = torch.zeros(x.shape[0], M, dtype=torch.float64)
    n, dt = int(c["n"]), float(c["dt"])
    ts = S.time_grid(n, dt, "float64")
    traj = []
    for i in range(n):
        t = float(ts[i])
        s = torch.stack([gauss_score(mus[m], t, x, y) for m in range(M)])
        dx, dlogq, _ = O.or_step_cifar_literal(x, logq, s, _t(c["noise"][i], torch.float64), t, dt)
        x, logq = x + dx, logq + dlogq
        traj.append(logq.clone())
    return x, torch.stack(traj)


def test_oracle_matches_reference_cifar_loop():
    for name, c in _load("ref_cifar_loop.npz").items():
        assert np.allclose(S.time_grid(int(c["n"]), float(c["dt"]), "float64"), c["t"], rtol=0, atol=0), name
        x, traj = _oracle_cifar_loop(c)
        assert _rel(x, c["x"]) < 1e-10, name
        assert np.abs(traj[9::10].numpy() - c["logq"]).max() < 1e-7 * (1 + np.abs(c["logq"]).max()), name


def _our_params(c):
    from super_diffusion_b200.configs import vpsde
    from super_diffusion_b200.models import ddpm  # noqa: F401  (registers 'score-net')
    from super_diffusion_b200.models import utils as mutils
    config = vpsde.get_config(conditioned=(str(c["cfgfile"]) == "vpsdeA"))
    gen = torch.Generator().manual_seed(int(c["seed"]))
    model, params = mutils.init_model(gen, config, zero_init_scale=float(c["zero_init_scale"]))
    params = mutils.perturb_params(params, gen)
    return config, model, params


def _tree_items(p, prefix=""):
    for k, v in p.items():
        if isinstance(v, dict):
            yield from _tree_items(v, prefix + k + "/")
        else:
            yield prefix + k, v


def test_config_and_param_tree_match_reference():
    """cifar/configs/sm/cifar/vpsde*.py values and the Flax parameter names / shapes the reference's own init produces."""
    for name, c in _load("ref_scorenet.npz").items():
        config, _, params = _our_params(c)
        ref_cfg = json.loads(str(c["config_json"]))
        for sec, d in ref_cfg.items():
            if isinstance(d, dict):
                for k, v in d.items():
                    ours = config[sec][k]
                    if str(c["cfgfile"]) != "vpsde" and (sec, k) == ("data", "train_split"):
                        continue   # dataset split of the A/B variants: training-side only
                    assert (list(ours) if isinstance(ours, tuple) else ours) == v, (name, sec, k)
            else:
                assert config[sec] == d, (name, sec)
        ref_shapes = json.loads(str(c["shapes_json"]))
        ours = {k: list(v.shape) for k, v in _tree_items(params)}
        assert ours == ref_shapes, name
        assert sum(int(np.prod(s)) for s in ours.values()) == int(c["n_params"])
        chk = float(sum(v.double().abs().sum() for _, v in _tree_items(params)))
        assert abs(chk - float(c["param_checksum"])) <= 1e-9 * chk, "parameter generator drifted: regenerate the fixture"


def test_oracle_matches_reference_scorenet():
    for name, c in _load("ref_scorenet.npz").items():
        config, _, params = _our_params(c)
        p64 = OS.params_to(params, dtype=torch.float64)
        out = OS.scorenet_apply(p64, config, _t(c["t"], torch.float64), _t(c["x"], torch.float64), _t(c["y"]))
        assert _rel(out, c["out"]) < 1e-6, (name, _rel(out, c["out"]))   # fp32 frequency table in both (layers.py:455)


def test_oracle_matches_reference_toy():
    z = _load("ref_toy.npz")
    c = z["toy_calls"]
    x, dx, s1, s2, eps = (_t(c[k], torch.float64) for k in ("x", "dx", "s1", "s2", "eps"))
    t, dt = float(c["t"]), float(c["dt"])
    assert _rel(O.stoch_dll_toy_literal(t, dt, x, dx, s1, ndim=2), c["stoch_dll"]) < 1e-13
    assert _rel(O.select_kappa_toy_literal(t, dt, x, s1, s2, eps), c["select_kappa"]) < 1e-12
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    for mode in ("or", "and"):
        for prec, acc, tol in (("f64", "float64", 1e-9), ("f32", "float32", 2e-3)):
            c = z[f"toy_{mode}_{prec}"]
            n, dt, B = int(c["n"]), float(c["dt"]), int(c["bs"])
            x0 = _t(ref_normal(c["x0_key"], (B, 2)), torch.float64)
            assert np.array_equal(x0.numpy(), c["x0"].astype(np.float64))
            noise = torch.stack([_t(ref_normal(k, (B, 2))) for k in c["step_keys"]])
            xf, ll, tr = toy.loop_toy(fns, x0, noise, mode, n, dt, accumulate=acc, record=True)
            assert _rel(tr["x"][::250].permute(1, 0, 2), c["x_quarters"]) < tol, (mode, prec)
            assert _rel(tr["ll"][::50].permute(1, 0, 2), c["ll"]) < tol, (mode, prec)
            if mode == "and":
                k_ref = c["kappa"]
                err = np.abs(tr["kappa"][::50].numpy() - k_ref) / (1 + np.abs(k_ref))
                assert err.max() < (1e-9 if prec == "f64" else 2e-2), (mode, prec, err.max())
    # F10: the notebook's float32 `t += -dt` really ends 0.93 % high
    assert abs(float(z["toy_or_f32"]["t_final"]) - (S.time_grid(1001, 1e-3, "float32")[-1])) < 1e-9


def ref_probe(key, shape):
    """The +-1 probe behind the shim's jax.random.randint(key, shape, 0, 2).astype(float)*2 - 1."""
    return (np.random.default_rng(int(key)).integers(0, 2, tuple(shape)).astype(np.float32) * 2 - 1)


def test_oracle_matches_reference_toy_ode():
    """The notebook's deterministic cells (vector_field / get_dll / get_kappa and the AND, kappa = 1/2 and OR loops)."""
    z = _load("ref_toy.npz")
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    for mode in ("and", "avg", "or"):
        c = z[f"toy_ode_{mode}"]
        n, dt, B = int(c["n"]), float(c["dt"]), int(c["bs"])
        x0 = _t(c["x0"], torch.float64)
        probes = torch.stack([_t(ref_probe(k, (B, 2))) for k in c["step_keys"]])
        xf, ll, tr = toy.loop_toy_ode(fns, x0, probes, mode, n, dt, accumulate="float64", record=True)
        # the fixture's jax.jvp stand-in is a central difference (h = 1e-6)
        assert _rel(tr["x"][::250].permute(1, 0, 2), c["x_quarters"]) < 1e-6, (mode, _rel(tr["x"][::250].permute(1, 0, 2), c["x_quarters"]))
        assert _rel(tr["ll"][::50].permute(1, 0, 2), c["ll"]) < 1e-6, (mode, _rel(tr["ll"][::50].permute(1, 0, 2), c["ll"]))


def test_oracle_matches_reference_sd_and_ode():
    """clip_eval.py:377-391 (method "and_ode"): teacher-forced on the recorded velocities / divergences, and closed-loop with
    torch.func.jvp through the same UNet stand-in."""
    c = _load("ref_sd.npz")["sd_and_ode"]
    N = int(c["N"])
    sig, ts, _ = S.edm_sigmas(N)
    lat = _t(c["latents0"])
    ll = torch.ones(lat.shape[0], 2, dtype=torch.float64)
    ph = c["emb_phase"]
    for i in range(N):
        sigma, dsigma = float(sig[i]), float(sig[i + 1]) - float(sig[i])
        eps = _t(c["probes"][i], torch.float64)
        f = lambda _x, p: sd_unet_stub(_x / ((sigma ** 2 + 1) ** 0.5), ts[i], p)
        vo, jo = torch.func.jvp(lambda _x: f(_x, ph[0]), (lat,), (eps,))
        vb, jb = torch.func.jvp(lambda _x: f(_x, ph[2]), (lat,), (eps,))
        vu = f(lat, ph[1])
        dlo, dlb = -(eps * jo).sum((1, 2, 3)), -(eps * jb).sum((1, 2, 3))
        assert _rel(vo, c["v_obj"][i]) < 1e-10 and _rel(dlo, c["dlog_obj"][i]) < 1e-10 and _rel(dlb, c["dlog_bg"][i]) < 1e-10
        dx, ll, kappa = O.sd_and_ode_step_literal(lat, vo, vb, vu, dlo, dlb, ll, sigma, dsigma, guidance_scale=float(c["guidance"]),
                                                  num_inference_steps=N)
        lat = lat + dx
        assert np.abs(kappa.numpy() - c["kappa"][i + 1]).max() < 1e-10 * (1 + np.abs(c["kappa"][i + 1]).max()), i
        assert np.abs(ll[:, 0].numpy() - c["ll_obj"][i + 1]).max() < 1e-9 * (1 + np.abs(c["ll_obj"][i + 1]).max()), i
        assert np.abs(ll[:, 1].numpy() - c["ll_bg"][i + 1]).max() < 1e-9 * (1 + np.abs(c["ll_bg"][i + 1]).max()), i
    assert _rel(lat, c["latents"]) < 1e-10


def test_oracle_matches_reference_sd():
    for name, c in _load("ref_sd.npz").items():
        if name == "sd_and_ode":
            continue
        method, N = str(c["method"]), int(c["N"])
        sig, ts, init = S.edm_sigmas(N)
        assert np.array_equal(sig.astype(np.float64), c["sigmas"])
        lat = _t(c["latents0"])
        ll = torch.ones(lat.shape[0], 2, dtype=torch.float64)
        for i in range(N):
            sigma, dsigma = float(sig[i]), float(sig[i + 1]) - float(sig[i])   # fp64 difference, like the fp64 fixture run
            vo, vb, vu = (_t(c[k][i]) for k in ("v_obj", "v_bg", "v_unc"))
            # velocities are functions of the running latents: regenerate them to prove the loop is closed
            ph = c["emb_phase"]
            x_in = lat / ((sigma ** 2 + 1) ** 0.5)
            assert _rel(sd_unet_stub(x_in, ts[i], ph[0]), vo) < 1e-9 and _rel(sd_unet_stub(x_in, ts[i], ph[2]), vb) < 1e-9
            dx, ll, kappa = O.sd_step_literal(lat, _t(c["z"][i], torch.float64), vo, vb, vu, ll, sigma, dsigma, method,
                                              guidance_scale=float(c["guidance"]), lift=0.0, num_inference_steps=N,
                                              T=float(c["T"]), logp=float(c["logp"]))
            lat = lat + dx
            assert np.abs(ll[:, 0].numpy() - c["ll_obj"][i + 1]).max() < 1e-9 * (1 + np.abs(c["ll_obj"][i + 1]).max()), (name, i)
            assert np.abs(ll[:, 1].numpy() - c["ll_bg"][i + 1]).max() < 1e-9 * (1 + np.abs(c["ll_bg"][i + 1]).max()), (name, i)
            assert np.abs(kappa.numpy() - c["kappa"][i + 1]).max() < 1e-10 * (1 + np.abs(c["kappa"][i + 1]).max()), (name, i)
        assert _rel(lat, c["latents"]) < 1e-10, name


# ---------------------------------------------------------------------------------------------
# GPU: CUDA path == reference outputs
# ---------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_cuda_matches_reference_cifar_steps(cuda):
    """cifar/dynamics.py closures (same names / signatures) over the fused step kernel, fed the reference's draw."""
    from super_diffusion_b200 import dynamics
    from super_diffusion_b200.models.utils import State
    for name, c in _load("ref_cifar_steps.npz").items():
        kind, t, dt = str(c["kind"]), float(c["t"]), float(c["dt"])
        x, eps, logq = (_t(c[k]).to(cuda) for k in ("x", "eps", "logq"))
        tables = [_t(s).to(cuda).contiguous() for s in c["scores"]]
        models = [TableModelTorch() for _ in tables]
        states = [State(params_ema={"table": tb}, model_params={"table": tb}) for tb in tables]
        args = {"key": 7, "labels": _t(c["labels"]).to(cuda), "dt": dt, "noise": eps}
        if kind == "stoch":
            vf = dynamics.get_joint_stoch_vf(0, models, states)
        elif kind in ("avg1", "avg0"):
            vf = dynamics.get_avg_vf(0, models, states, stoch=(kind == "avg1"))
        else:
            cfg = type("C", (), {"data": type("D", (), {"t_0": 0.0, "t_1": 1.0})})
            vf = dynamics.get_vpsde(cfg, models[0], train=False)[2]
            args["state"] = states[0]
        dx, dlogq = vf(t, (x, logq), args)
        # dx is formed as (x + dx) - x in fp32: absolute error ~ ulp(x)
        assert np.abs(dx.cpu().numpy() - c["dx"]).max() <= 1e-3 * np.abs(c["dx"]).max() + 4e-6, name
        assert dlogq.shape == c["dlogq"].shape, name
        assert np.abs(dlogq.cpu().numpy() - c["dlogq"]).max() <= 1e-3 * np.abs(c["dlogq"]).max() + 1e-30, name
        if kind == "stoch":
            # mixing weights on identical inputs within 1e-4 (reference: softmax(1e6 * logq), dynamics.py:124)
            from super_diffusion_b200 import ops, sde
            _, _, w = ops.step_vpsde(x, eps, tables, logq.clone(), sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t), dt,
                                     ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB, temperature=1e6)
            w_ref = torch.softmax(1e6 * _t(c["logq"], torch.float64), dim=-1).numpy()
            assert np.abs(w.cpu().numpy() - w_ref).max() <= 1e-4, name


@pytest.mark.gpu
def test_cuda_matches_reference_cifar_ode(cuda):
    """dynamics.get_joint_vf over sd_rowdot + sd_step_vpsde_ode, torch.func.jvp for the caller-supplied models."""
    from super_diffusion_b200 import dynamics
    from super_diffusion_b200.models.utils import State
    for name, c in _load("ref_cifar_ode.npz").items():
        t, dt = float(c["t"]), float(c["dt"])
        x, logq = _t(c["x"]).to(cuda), _t(c["logq"]).to(cuda)
        M = c["mus"].shape[0]
        models = [NonlinModelTorch() for _ in range(M)]
        states = [State(params_ema={"mu": _t(c["mus"][i]).to(cuda), "phi": float(c["phis"][i])}) for i in range(M)]
        vf = dynamics.get_joint_vf(0, models, states)
        dx, dlogq = vf(t, (x, logq), {"key": 3, "labels": None, "dt": dt, "probes": [_t(p).to(cuda) for p in c["probes"]]})
        assert np.abs(dx.cpu().numpy() - c["dx"]).max() <= 1e-3 * np.abs(c["dx"]).max() + 4e-6, name
        assert np.abs(dlogq.cpu().numpy() - c["dlogq"]).max() <= 1e-3 * np.abs(c["dlogq"]).max(), name


@pytest.mark.gpu
def test_cuda_matches_reference_cifar_loop(cuda):
    """cifar/eval_utils.py:72-86 artifact_generator over get_joint_stoch_vf: 200 steps, reference draws."""
    from super_diffusion_b200 import dynamics, eval_utils
    from super_diffusion_b200.config_dict import ConfigDict
    from super_diffusion_b200.models.utils import State
    for name, c in _load("ref_cifar_loop.npz").items():
        config = ConfigDict()
        config.eval, config.data = ConfigDict(), ConfigDict()
        config.eval.batch_size, config.data.image_size, config.data.num_channels = 4, 8, 3
        models = [GaussModelTorch() for _ in c["mus"]]
        states = [State(params_ema={"mu": _t(mu).to(cuda)}) for mu in c["mus"]]
        vf = dynamics.get_joint_stoch_vf(0, models, states)
        gen = eval_utils.get_generator(models, config, vf, device=cuda, return_logq=True)
        x, n, logq = gen(0, _t(c["labels"]).to(cuda), x0=_t(c["x0"]), noise=_t(c["noise"]))
        assert n == int(c["n"])
        assert _rel(x.cpu().numpy(), c["x"]) <= 1e-3, (name, _rel(x.cpu().numpy(), c["x"]))
        ref = c["logq"][-1]
        assert np.abs(logq.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max(), name


@pytest.mark.gpu
def test_cuda_matches_reference_scorenet(cuda):
    """The reference ScoreNet's output (fp64, reference source over the shims) vs the tcgen05 forward.
    bf16 operands / fp32 accumulation: stated separately from the fp32 gate (north_star) — rel-RMS <= 1.8e-2,
    max error <= 2.4e-2 of the output range (1.5x measured) (the oracle-vs-CUDA tests in test_scorenet_gpu.py use the same gate)."""
    from super_diffusion_b200.models import utils as mutils
    for name, c in _load("ref_scorenet.npz").items():
        config, model, params = _our_params(c)
        fn = mutils.get_model_fn(model, params)
        out = fn(_t(c["t"]).to(cuda), _t(c["x"]).to(cuda).contiguous(), _t(c["y"]).to(cuda))
        torch.cuda.synchronize()
        got, ref = out.double().cpu().numpy(), c["out"]
        rms = float(np.sqrt(((got - ref) ** 2).mean()) / np.sqrt((ref ** 2).mean()))
        assert rms <= 1.8e-2, (name, rms)          # measured 1.20e-2 / 1.06e-2 (profiles/r02_precision_report.json)
        assert np.abs(got - ref).max() <= 2.4e-2 * np.abs(ref).max(), (name, np.abs(got - ref).max())   # measured 1.56e-2 / 1.50e-2


@pytest.mark.gpu
def test_cuda_matches_reference_toy(cuda):
    """The notebook's own OR / AND cells (bs 512, 1000 steps, float32 time accumulation) vs superdiff_or / superdiff_and."""
    from super_diffusion_b200.superposition import superdiff_and, superdiff_or
    z = _load("ref_toy.npz")
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    for mode in ("or", "and"):
        c = z[f"toy_{mode}_f32"]
        n, dt, B = int(c["n"]), float(c["dt"]), int(c["bs"])
        x0 = _t(c["x0"].astype(np.float32)).to(cuda)
        noise = torch.stack([_t(ref_normal(k, (B, 2))) for k in c["step_keys"]]).to(cuda)
        run = superdiff_or if mode == "or" else superdiff_and
        x, ll, w, traj = run(fns, x0, n_steps=n, dt=dt, noise=noise, record=True)
        torch.cuda.synchronize()
        # the fixture itself is an fp32 run (NumPy evaluation order); 2e-3 = fp32-vs-fp32 along 1000 free-running steps
        assert _rel(traj["x"][::250].permute(1, 0, 2).cpu().numpy(), c["x_quarters"]) <= 2e-3, mode
        assert _rel(traj["ll"][::50].permute(1, 0, 2).cpu().numpy(), c["ll"]) <= 2e-3, mode
        # and against the fp64 run of the same cells (time accumulated in fp64 there: F10 drift shows up at ~1e-2)
        c64 = z[f"toy_{mode}_f64"]
        assert _rel(x.cpu().numpy(), c64["x_final"]) <= 5e-2, mode


@pytest.mark.gpu
def test_cuda_matches_reference_toy_ode(cuda):
    """superposition.superdiff_ode (torch.func.jvp + sd_rowdot + sd_step_vpsde_ode) vs the notebook's ODE cells, 1000 steps."""
    from super_diffusion_b200.superposition import superdiff_ode
    z = _load("ref_toy.npz")
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    for mode in ("and", "avg", "or"):
        c = z[f"toy_ode_{mode}"]
        n, dt, B = int(c["n"]), float(c["dt"]), int(c["bs"])
        probes = torch.stack([_t(ref_probe(k, (B, 2))) for k in c["step_keys"]]).to(cuda)
        x, ll, kappa, traj = superdiff_ode(fns, _t(c["x0"]).float().to(cuda), mode=mode, n_steps=n, dt=dt, probes=probes,
                                           accumulate="float64", record=True)
        torch.cuda.synchronize()
        ex = _rel(traj["x"][::250].permute(1, 0, 2).cpu().numpy(), c["x_quarters"])
        el = _rel(traj["ll"][::50].permute(1, 0, 2).cpu().numpy(), c["ll"])
        assert ex <= 1e-3 and el <= 1e-3, (mode, ex, el)


@pytest.mark.gpu
def test_cuda_matches_reference_sd(cuda):
    """clip_eval.py run(args) and / or / avg: teacher-forced per step (identical inputs) and free-running."""
    from super_diffusion_b200 import ops
    from super_diffusion_b200.superposition import sd_superdiff
    for name, c in _load("ref_sd.npz").items():
        if name == "sd_and_ode":
            continue
        method, N = str(c["method"]), int(c["N"])
        mode = {"and": ops.MODE_AND, "or": ops.MODE_OR, "avg": ops.MODE_AVG}[method]
        sig = c["sigmas"]
        # teacher-forced: rebuild the reference's latents_i from its own recorded quantities
        lat = _t(c["latents0"], torch.float64)
        for i in range(N):
            sigma, dsigma = float(sig[i]), float(sig[i + 1]) - float(sig[i])   # fp64 difference, like the fp64 fixture run
            vo, vb, vu = (_t(c[k][i], torch.float64) for k in ("v_obj", "v_bg", "v_unc"))
            ll_in = torch.stack([_t(c["ll_obj"][i]), _t(c["ll_bg"][i])], 1)
            f = lambda a: a.float().to(cuda).contiguous()
            lo, ll, kap = ops.step_edm_cfg(f(lat), f(_t(c["z"][i])), f(vo), f(vb), f(vu), f(ll_in), sigma, dsigma, mode,
                                           guidance=float(c["guidance"]), lift_term=0.0, temperature=float(c["T"]),
                                           logp=float(c["logp"]), kappa_fixed=0.5)
            dxr, llr, kr = O.sd_step_literal(lat, _t(c["z"][i], torch.float64), vo, vb, vu, ll_in, sigma, dsigma, method,
                                             guidance_scale=float(c["guidance"]), num_inference_steps=N, T=float(c["T"]),
                                             logp=float(c["logp"]))
            lat = lat + dxr
            assert np.abs(kap.cpu().numpy() - c["kappa"][i + 1]).max() <= 1e-4 * (1 + np.abs(c["kappa"][i + 1]).max()), (name, i)
            ref_ll = np.stack([c["ll_obj"][i + 1], c["ll_bg"][i + 1]], 1)
            assert np.abs(ll.cpu().numpy() - ref_ll).max() <= 1e-3 * np.abs(ref_ll).max(), (name, i)
            assert _rel(lo.cpu().numpy(), lat.numpy()) <= 1e-3, (name, i)
        assert _rel(lat.numpy(), c["latents"]) < 1e-10
        # free-running through the public loop with the same UNet stand-in
        ph = {"obj": float(c["emb_phase"][0]), "uncond": float(c["emb_phase"][1]), "bg": float(c["emb_phase"][2])}

        def get_vel(t, sigma, latents, which):
            return sd_unet_stub(latents / ((sigma ** 2 + 1) ** 0.5), t, ph[which])

        init = float(np.max(sig))
        x, ll, kappa, traj = sd_superdiff(get_vel, (_t(c["latents0"]) / init).float().to(cuda), method=method,
                                          num_inference_steps=N, guidance_scale=float(c["guidance"]), T=float(c["T"]),
                                          logp=float(c["logp"]), noise=_t(c["z"]).to(cuda), record=True)
        torch.cuda.synchronize()
        assert _rel(x.cpu().numpy(), c["latents"]) <= 2e-3, (name, _rel(x.cpu().numpy(), c["latents"]))
        ref_ll = np.stack([c["ll_obj"][-1], c["ll_bg"][-1]], 1)
        assert np.abs(ll.cpu().numpy() - ref_ll).max() <= 2e-3 * np.abs(ref_ll).max(), name


@pytest.mark.gpu
def test_cuda_matches_reference_sd_and_ode(cuda):
    """sd_step_edm_ode teacher-forced on the reference's recorded tensors, then sd_superdiff(method="and_ode") closed loop."""
    from super_diffusion_b200 import ops
    from super_diffusion_b200.superposition import sd_superdiff
    c = _load("ref_sd.npz")["sd_and_ode"]
    N = int(c["N"])
    sig = c["sigmas"]
    f = lambda a: _t(np.asarray(a)).float().to(cuda).contiguous()
    lat = _t(c["latents0"], torch.float64)
    for i in range(N):
        sigma, dsigma = float(sig[i]), float(sig[i + 1] - sig[i])
        vo, vb, vu = (_t(c[k][i], torch.float64) for k in ("v_obj", "v_bg", "v_unc"))
        dlo, dlb = _t(c["dlog_obj"][i], torch.float64), _t(c["dlog_bg"][i], torch.float64)
        ll_in = torch.stack([_t(c["ll_obj"][i]), _t(c["ll_bg"][i])], 1)
        lo, ll, kap = ops.step_edm_ode(f(lat), f(vo), f(vb), f(vu), f(torch.stack([dlo, dlb], 1)), f(ll_in), sigma, dsigma,
                                       guidance=float(c["guidance"]), lift_term=0.0)
        dxr, _, _ = O.sd_and_ode_step_literal(lat, vo, vb, vu, dlo, dlb, ll_in, sigma, dsigma, guidance_scale=float(c["guidance"]),
                                              num_inference_steps=N)
        lat = lat + dxr
        assert np.abs(kap.cpu().numpy() - c["kappa"][i + 1]).max() <= 1e-4 * (1 + np.abs(c["kappa"][i + 1]).max()), i
        ref_ll = np.stack([c["ll_obj"][i + 1], c["ll_bg"][i + 1]], 1)
        assert np.abs(ll.cpu().numpy() - ref_ll).max() <= 1e-3 * np.abs(ref_ll).max(), i
        assert _rel(lo.cpu().numpy(), lat.numpy()) <= 1e-3, i
    ph = {"obj": float(c["emb_phase"][0]), "uncond": float(c["emb_phase"][1]), "bg": float(c["emb_phase"][2])}

    def get_vel(t, sigma, latents, which):
        return sd_unet_stub(latents / ((sigma ** 2 + 1) ** 0.5), t, ph[which])

    init = float(np.max(sig))
    x, ll, kappa, _ = sd_superdiff(get_vel, (_t(c["latents0"]) / init).float().to(cuda), method="and_ode", num_inference_steps=N,
                                   guidance_scale=float(c["guidance"]), noise=_t(c["probes"]).to(cuda))
    torch.cuda.synchronize()
    assert _rel(x.cpu().numpy(), c["latents"]) <= 2e-3, _rel(x.cpu().numpy(), c["latents"])
    ref_ll = np.stack([c["ll_obj"][-1], c["ll_bg"][-1]], 1)
    assert np.abs(ll.cpu().numpy() - ref_ll).max() <= 2e-3 * np.abs(ref_ll).max()


# ---------------------------------------------------------------------------------------------------------------------
# applications/proteins/superdiff/composition.py: two-component (translations / rotations) AND / OR mixing.
# Fixture = the reference's own kappa_AND / kappa_OR / compute_kappas / compute_stoch_dll method bodies executed on a stub
# object (tests/golden/make_protein_vectors.py).
# ---------------------------------------------------------------------------------------------------------------------
def _protein_cases():
    z = np.load(os.path.join(HERE, "ref_protein.npz"))
    for case, op in (("and", "AND"), ("or", "OR"), ("and_lift", "AND")):
        yield case, op, (lambda k, case=case: z[f"{case}/{k}"])


def test_oracle_matches_reference_protein_mixing():
    for case, op, g in _protein_cases():
        n, dt = int(g("n_steps")), float(g("dt"))
        ll = [0.0, 0.0, 0.0, 0.0]
        for i in range(n):
            sc = {k: _t(g("s_" + k)[i], torch.float64) for k in ("pt", "ft", "pr", "fr")}
            dxt, dxr, kt, kr, ll = O.protein_step_literal(
                _t(g("x")[i], torch.float64), sc, _t(g("eps")[i], torch.float64), ll, float(g("a_trans")[i]), float(g("beta_trans")[i]),
                float(g("beta_rots")[i]), dt, op, T=float(g("T")), logp=float(g("logp")), lift_trans=float(g("lift_trans")[i]),
                lift_rots=float(g("lift_rots")[i]))
            tol = 1e-10 if op == "AND" else 2e-7       # OR: the reference's kappa is a float32 softmax of float32 accumulators (:178-181,434)
            assert _rel(dxt, g("dx_trans")[i]) < tol and _rel(dxr, g("dx_rots")[i]) < tol, (case, i)
            assert abs(float(kt) - float(g("kappa_trans")[i].ravel()[0])) < 1e-7, (case, i)     # OR: softmax of float32 accumulators
            assert abs(float(kr) - float(g("kappa_rots")[i].ravel()[0])) < 1e-7, (case, i)
            ref_ll = g("ll")[i]
            assert max(abs(float(a) - b) / (1 + abs(b)) for a, b in zip(ll, ref_ll)) < 1e-6, (case, i)   # reference accumulates ll in float32
            ll = [float(v) for v in ref_ll]


@pytest.mark.gpu
def test_cuda_matches_reference_protein_mixing(cuda):
    """superposition.protein_superdiff_step (two launches of the fused VP-SDE step kernel per timestep, sigma = 1) teacher-forced on
    the reference's recorded states: dx and kappa within 1e-4, log-likelihood increments within 1e-5 of their magnitude."""
    from super_diffusion_b200.superposition import protein_superdiff_step
    for case, op, g in _protein_cases():
        n, dt = int(g("n_steps")), float(g("dt"))
        prev = np.zeros(4)
        for i in range(n):
            f = lambda a: _t(a.astype(np.float32)).to(cuda).contiguous()
            sc = {k: f(g("s_" + k)[i]) for k in ("pt", "ft", "pr", "fr")}
            ll = _t(prev.astype(np.float32)).to(cuda).reshape(1, 4).contiguous()
            dxt, dxr, kt, kr, ll = protein_superdiff_step(
                f(g("x")[i]), sc, f(g("eps")[i]), ll, float(g("a_trans")[i]), float(g("beta_trans")[i]), float(g("beta_rots")[i]), dt,
                operator=op, T=float(g("T")), logp=float(g("logp")), lift_trans=float(g("lift_trans")[i]), lift_rots=float(g("lift_rots")[i]))
            torch.cuda.synchronize()
            assert _rel(dxt.cpu(), g("dx_trans")[i]) <= 1e-4 and _rel(dxr.cpu(), g("dx_rots")[i]) <= 1e-4, (case, i)
            assert abs(float(kt) - float(g("kappa_trans")[i].ravel()[0])) <= 1e-4, (case, i)
            assert abs(float(kr) - float(g("kappa_rots")[i].ravel()[0])) <= 1e-4, (case, i)
            ref_ll = g("ll")[i]
            inc, ref_inc = ll.cpu().numpy().ravel().astype(np.float64) - prev.astype(np.float32), ref_ll - prev
            assert np.abs(inc - ref_inc).max() <= 1e-5 * (1 + np.abs(ref_inc).max()) + 1e-7 * np.abs(ref_ll).max(), (case, i)
            prev = ref_ll
