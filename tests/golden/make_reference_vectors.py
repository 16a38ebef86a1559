"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN SOURCE FILES in this container.

    python tests/golden/make_reference_vectors.py      # needs /root/reference; rewrites tests/golden/ref_*.npz

The reference cannot be installed here (JAX / Flax / diffusers / ml_collections absent, no network),
but its sampling path is plain array code, so the third-party packages are replaced by the NumPy /
CPU-PyTorch API shims of tests/golden/ref_shims/ (README there) and the reference files are imported
*unmodified from where they lie* under /root/reference:

    cifar/dynamics.py            get_joint_stoch_vf, get_avg_vf, get_vpsde          -> ref_cifar_steps.npz
    cifar/dynamics.py            get_joint_vf (ODE + Hutchinson probes)             -> ref_cifar_ode.npz
    cifar/eval_utils.py          get_generator.artifact_generator (the loop)        -> ref_cifar_loop.npz
    cifar/models/{ddpm,layers,normalization,utils}.py + configs/sm/cifar/vpsde*.py   -> ref_scorenet.npz
    notebooks/superposition_edu.ipynb  code cells (get_stoch_dll, select_kappa, the OR / AND loops)
                                                                                    -> ref_toy.npz
    applications/images/clip_eval.py   run(args) for method and / or / avg          -> ref_sd.npz

Nothing is copied from the reference: cells / modules are read and executed at generation time only.
What the shims restate instead of pinning (third-party primitives, random streams) is listed in
ref_shims/README.md.  The fixtures travel; /root/reference and the shims are never needed at test
time.
"""
import importlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SD_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(1, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def f32r(a):
    """Round to float32-representable values (so fp32 kernels consume the inputs exactly)."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


# ---------------------------------------------------------------------------------------------
# cifar/
# ---------------------------------------------------------------------------------------------

def import_cifar():
    for n in ("optax", "diffrax", "orbax"):
        _stub(n)
    sys.path.insert(0, os.path.join(REF, "cifar"))
    import jax
    dynamics = importlib.import_module("dynamics")
    eval_utils = importlib.import_module("eval_utils")
    mutils = importlib.import_module("models.utils")
    ddpm = importlib.import_module("models.ddpm")
    assert dynamics.__file__.startswith(REF) and ddpm.__file__.startswith(REF)
    return jax, dynamics, eval_utils, mutils, ddpm


class TableModel:
    """Stand-in for a flax module at the model seam (cifar/models/utils.py:86-96 calls
    model.apply(variables, t, x, y, train=False, mutable=False)): returns a fixed score tensor."""

    def apply(self, variables, t, x, y, train=False, mutable=False):
        assert t.shape == (x.shape[0], 1, 1, 1)
        return variables["params"]["table"]


class GaussModel:
    """Analytic score model: data ~ N(mu + 0.1*y, 0.25 I) under q_t = N(alpha_t x_1, t^2):
    returns sigma_t * grad log q_t(x) (what the reference's networks approximate)."""

    def apply(self, variables, t, x, y, train=False, mutable=False):
        mu = variables["params"]["mu"] + 0.1 * np.asarray(y, dtype=x.dtype)[:, None, None, None]
        alpha = np.exp(-0.5 * t * 0.1 - 0.25 * t ** 2 * 19.9)
        var = alpha ** 2 * 0.25 + t ** 2
        return -t * (x - alpha * mu) / var


class St:
    def __init__(self, params):
        self.params_ema = params
        self.model_params = params


def cifar_steps(jax, dynamics):
    jax.config.update("jax_enable_x64", True)
    from jax import random as jr
    out = {}
    rng = np.random.default_rng(20261018)
    cases = [("or_m2", "stoch", 2, 2, 0.62, 5e-3), ("or_m3_last", "stoch", 3, 2, 5e-3, 5e-3),
             ("or_m2_t1", "stoch", 2, 3, 1.0, 1e-3), ("avg_stoch", "avg1", 2, 2, 0.4, 5e-3),
             ("avg_ode", "avg0", 2, 2, 0.4, 5e-3), ("single_ode", "vpsde", 1, 2, 0.7, 1e-2)]
    for name, kind, M, B, t, dt in cases:
        shape = (B, 32, 32, 3)
        x = f32r(rng.standard_normal(shape))
        tables = [f32r(rng.standard_normal(shape)) for _ in range(M)]
        logq = f32r(1e-6 * rng.standard_normal((B, M)))
        labels = np.arange(B) % 10
        models = [TableModel() for _ in range(M)]
        states = [St({"table": tb}) for tb in tables]
        jr.DRAWS.clear()
        args = {"key": 7, "labels": labels, "dt": dt}
        if kind == "stoch":
            vf = dynamics.get_joint_stoch_vf(0, models, states)
        elif kind == "avg1":
            vf = dynamics.get_avg_vf(0, models, states, stoch=True)
        elif kind == "avg0":
            vf = dynamics.get_avg_vf(0, models, states, stoch=False)
        else:
            cfg = types.SimpleNamespace(data=types.SimpleNamespace(t_0=0.0, t_1=1.0))
            vf = dynamics.get_vpsde(cfg, models[0], train=False)[2]
            states[0].model_params = {"table": tables[0]}
            args["state"] = states[0]
        dx, dlogq = vf(t, (x, logq), args)
        eps = jr.DRAWS[0] if jr.DRAWS else np.zeros(shape)
        assert len(jr.DRAWS) <= 1
        f4 = lambda a: np.asarray(a, dtype=np.float32)        # inputs are fp32-representable: store them as fp32
        assert np.array_equal(f4(eps).astype(np.float64), eps)
        out[name] = dict(kind=kind, t=t, dt=dt, x=f4(x), logq=f4(logq), labels=labels, scores=f4(np.stack(tables)), eps=f4(eps),
                         dx=np.asarray(dx), dlogq=np.asarray(dlogq))
    return out


class NonlinModel:
    """Smooth non-linear score stand-in with a non-diagonal Jacobian (tests re-implement it in torch):
    s = -t (x - alpha mu)/var + 0.3 sin(1.7 x + phi) + 0.2 x * roll(x, 1, axis=2)."""

    def apply(self, variables, t, x, y, train=False, mutable=False):
        p = variables["params"]
        alpha = np.exp(-0.5 * t * 0.1 - 0.25 * t ** 2 * 19.9)
        var = alpha ** 2 * 0.25 + t ** 2
        return -t * (x - alpha * p["mu"]) / var + 0.3 * np.sin(1.7 * x + p["phi"]) + 0.2 * x * np.roll(x, 1, axis=2)


def cifar_ode_steps(jax, dynamics):
    """cifar/dynamics.py:59-97 get_joint_vf (deterministic SuperDiff-OR with Hutchinson probes)."""
    jax.config.update("jax_enable_x64", True)
    from jax import random as jr
    out = {}
    rng = np.random.default_rng(77)
    for name, M, B, t, dt in (("ode_m2", 2, 3, 0.55, 5e-3), ("ode_m3", 3, 2, 0.08, 1e-2)):
        shape = (B, 8, 8, 3)
        x = f32r(rng.standard_normal(shape))
        logq = f32r(1e-6 * rng.standard_normal((B, M)))
        mus = [f32r(0.6 * rng.standard_normal((1, 8, 8, 3))) for _ in range(M)]
        phis = [float(np.float32(rng.uniform(0, 3))) for _ in range(M)]
        models = [NonlinModel() for _ in range(M)]
        states = [St({"mu": mus[i], "phi": phis[i]}) for i in range(M)]
        vf = dynamics.get_joint_vf(0, models, states)
        jr.DRAWS.clear(); jr.LOG.clear()
        dx, dlogq = vf(t, (x, logq), {"key": 3, "labels": np.arange(B) % 10, "dt": dt})
        assert [k for k, *_ in jr.LOG] == ["randint"] * M
        probes = np.stack([d.astype(np.float32) * 2 - 1 for d in jr.DRAWS])
        out[name] = dict(t=t, dt=dt, x=x.astype(np.float32), logq=logq.astype(np.float32), mus=np.stack(mus).astype(np.float32),
                         phis=np.asarray(phis), probes=probes, dx=np.asarray(dx), dlogq=np.asarray(dlogq))
    return out


def cifar_loop(jax, dynamics, eval_utils):
    jax.config.update("jax_enable_x64", True)
    from jax import random as jr
    import ml_collections
    out = {}
    for name, M in (("gen_or_m2", 2), ("gen_or_m3", 3)):
        config = ml_collections.ConfigDict()
        config.eval = ml_collections.ConfigDict()
        config.data = ml_collections.ConfigDict()
        config.eval.batch_size, config.data.image_size, config.data.num_channels = 4, 8, 3
        rng = np.random.default_rng(100 + M)
        mus = [f32r(0.8 * rng.standard_normal((1, 8, 8, 3))) for _ in range(M)]
        models = [GaussModel() for _ in range(M)]
        states = [St({"mu": mu}) for mu in mus]
        inner = dynamics.get_joint_stoch_vf(0, models, states)
        rec = {"dlogq": [], "t": []}

        def vf(t, data, args, inner=inner, rec=rec):
            dx, dlogq = inner(t, data, args)
            rec["dlogq"].append(np.asarray(dlogq).copy())
            rec["t"].append(t)
            return dx, dlogq

        gen = eval_utils.get_generator(models, config, vf)
        labels = np.array([0, 3, 5, 9])
        jr.DRAWS.clear()
        x, n = gen(11 + M, labels)
        draws = list(jr.DRAWS)
        assert len(draws) == n + 1 and n == 200
        out[name] = dict(mus=np.stack(mus), labels=labels, x0=draws[0].astype(np.float32),
                         noise=np.stack(draws[1:]).astype(np.float32), n=n, dt=5e-3, t=np.asarray(rec["t"]),
                         x=np.asarray(x), logq=np.cumsum(np.stack(rec["dlogq"]), 0)[9::10])
    return out


def tree_to_numpy(p, dtype=np.float64):
    if isinstance(p, dict):
        return {k: tree_to_numpy(v, dtype) for k, v in p.items()}
    return p.detach().cpu().numpy().astype(dtype)


def tree_shapes(p, prefix=""):
    out = {}
    for k, v in p.items():
        if isinstance(v, dict):
            out.update(tree_shapes(v, prefix + k + "/"))
        else:
            out[prefix + k] = tuple(int(s) for s in np.shape(v))
    return out


def tree_checksum(p):
    return float(sum(np.abs(v).sum() for v in _leaves(p)))


def _leaves(p):
    for v in p.values():
        if isinstance(v, dict):
            yield from _leaves(v)
        else:
            yield np.asarray(v)


def scorenet(jax, mutils_ref):
    """The reference ScoreNet (ddpm.py) applied, through the reference's get_model_fn, to parameter
    trees produced by THIS repo's init (Flax names / layouts): pins module wiring + naming."""
    jax.config.update("jax_enable_x64", True)
    from super_diffusion_b200.models import utils as our_mutils
    from super_diffusion_b200.models import ddpm as _our_ddpm  # noqa: F401  (registers 'score-net' on our side)
    out = {}
    for name, cfgfile, seed, B in (("vpsde", "vpsde", 1, 2), ("vpsdeA_conditioned", "vpsdeA", 2, 2)):
        spec = importlib.util.spec_from_file_location("refcfg_" + cfgfile, os.path.join(REF, "cifar/configs/sm/cifar", cfgfile + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        config = mod.get_config()
        # 1. parameter-tree structure of the reference's own init (names + shapes)
        model, ref_params = mutils_ref.init_model(5, config)
        ref_shapes = tree_shapes(ref_params)
        # 2. our init (same names expected), non-degenerate variant, fp32-representable values
        gen = torch.Generator().manual_seed(seed)
        _, ours = our_mutils.init_model(gen, config, zero_init_scale=1.0)
        ours = our_mutils.perturb_params(ours, gen)
        our_shapes = tree_shapes(tree_to_numpy(ours))
        assert ref_shapes == our_shapes, (set(ref_shapes.items()) ^ set(our_shapes.items()))
        params = tree_to_numpy(ours)
        rng = np.random.default_rng(seed)
        x = f32r(rng.standard_normal((B, 32, 32, 3)))
        t = f32r(np.array([0.3, 0.9])[:B]).reshape(B, 1, 1, 1)
        y = np.array([3, 7][:B])
        model_fn = mutils_ref.get_model_fn(model, params, train=False)
        score = np.asarray(model_fn(t, x, y))
        out[name] = dict(cfgfile=cfgfile, seed=seed, zero_init_scale=1.0, x=x.astype(np.float32), t=t.astype(np.float32), y=y,
                         out=score, param_checksum=tree_checksum(params), n_params=sum(int(np.prod(s)) for s in ref_shapes.values()),
                         shapes_json=json.dumps({k: list(v) for k, v in sorted(ref_shapes.items())}),
                         config_json=json.dumps({k: dict(v) if isinstance(v, dict) else v for k, v in config.items()}, default=list))
        print(name, "params", out[name]["n_params"], "out rms", float(np.sqrt((score ** 2).mean())))
    return out


# ---------------------------------------------------------------------------------------------
# notebooks/superposition_edu.ipynb
# ---------------------------------------------------------------------------------------------

def toy(jax):
    from jax import random as jr
    nb = json.load(open(os.path.join(REF, "notebooks/superposition_edu.ipynb")))
    code = [("".join(c["source"]), i) for i, c in enumerate(nb["cells"]) if c["cell_type"] == "code"]

    def cell(marker, nth=0):
        hits = [s for s, _ in code if marker in s]
        return hits[nth]

    defs = cell("def sample_data")                # schedule lambdas, ndim, q_t
    sdll = cell("def get_sscore")                 # get_sscore + get_stoch_dll
    or_loop = cell("x_gen_or = jnp.copy(x_gen)")
    kap = cell("def select_kappa")
    and_loop = cell("x_gen_and = jnp.copy(x_gen)")

    MEANS = {"up": [(-1.5, 1.5), (1.5, 1.5)], "down": [(-1.5, -1.5), (1.5, -1.5)]}

    def mixture_state(which):
        mu = np.array(MEANS[which])

        def apply_fn(params, t, x):
            dt_ = x.dtype
            t = np.asarray(t, dtype=dt_)
            alpha = np.exp(-0.5 * t * 0.1 - 0.25 * t ** 2 * 19.9).astype(dt_)
            var = ((alpha * 0.4) ** 2 + t ** 2).astype(dt_)
            diff = x[:, None, :] - alpha[:, None, :] * mu[None].astype(dt_)
            logw = -0.5 * (diff ** 2).sum(-1) / var
            w = np.exp(logw - logw.max(1, keepdims=True))
            w = w / w.sum(1, keepdims=True)
            grad = -(w[:, :, None] * diff).sum(1) / var
            return (t * grad).astype(dt_).view(type(x)) if hasattr(x, "at") else (t * grad).astype(dt_)
        return types.SimpleNamespace(apply_fn=apply_fn, params=None)

    out = {}
    for mode in ("f32", "f64"):
        jax.config.update("jax_enable_x64", mode == "f64")
        import jax.numpy as jnp
        from tqdm import trange
        from functools import partial
        ns = dict(jax=jax, jnp=jnp, np=np, random=jr, trange=lambda n: range(n), partial=partial,
                  state_up=mixture_state("up"), state_down=mixture_state("down"))
        exec(defs, ns)
        exec(sdll, ns)
        for loop_name, src in (("or", or_loop), ("and", and_loop)):
            if loop_name == "and":
                exec(kap, ns)
                kappas = []
                inner = ns["select_kappa"]

                def rec_kappa(*a, inner=inner, kappas=kappas):
                    k = inner(*a)
                    kappas.append(np.asarray(k).copy())
                    return k
                ns["select_kappa"] = rec_kappa
            ns["key"] = jr.PRNGKey(1234 if loop_name == "or" else 4321)
            ns["x_t"] = jnp.zeros((512, 2))       # the notebook reads x_t.shape[1] left over from an earlier cell
            jr.LOG.clear(); jr.DRAWS.clear()
            exec(src, ns)
            keys = [k for kind, k, shp in jr.LOG]
            n = ns["n"]
            # draw 0 = x0; OR: one draw per step; AND: select_kappa and dx each draw from the same ikey
            per_step = 1 if loop_name == "or" else 2
            assert len(keys) == 1 + per_step * n, (len(keys), n)
            step_keys = keys[1::per_step]
            if per_step == 2:
                assert keys[1::2] == keys[2::2]
            x_gen = np.asarray(ns["x_gen"])
            ll1, ll2 = np.asarray(ns["ll_1"]), np.asarray(ns["ll_2"])
            d = dict(x0_key=keys[0], step_keys=np.asarray(step_keys, dtype=np.int64), n=n, dt=float(ns["dt"]), bs=int(ns["bs"]),
                     x0=x_gen[:, 0, :], x_final=x_gen[:, -1, :], x_quarters=x_gen[:, ::250, :],
                     ll=np.stack([ll1[:, ::50], ll2[:, ::50]], -1), t_final=np.asarray(ns["t"])[0, 0], x_dtype=str(x_gen.dtype))
            if loop_name == "and":
                d["kappa"] = np.stack(kappas)[::50]
            out[f"toy_{loop_name}_{mode}"] = d
            print(f"toy_{loop_name}_{mode}", "x rms", float(np.sqrt((x_gen[:, -1] ** 2).mean())), "t_final", d["t_final"])
    # ---- deterministic (ODE) cells: vector_field (jvp + Rademacher probe), get_dll / get_kappa, the three loops ----
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    from functools import partial
    vf_cell = cell("def vector_field(key,t,x,state)")
    dll_cell = cell("def get_dll")
    loops = {"and": cell("kappa = get_kappa("), "avg": cell("x_gen_avg = jnp.copy(x_gen)"), "or": cell("max_ll = jnp.maximum")}
    for mode, src in loops.items():
        ns = dict(jax=jax, jnp=jnp, np=np, random=jr, trange=lambda n: range(n), partial=partial,
                  state_up=mixture_state("up"), state_down=mixture_state("down"))
        exec(defs, ns); exec(vf_cell, ns); exec(dll_cell, ns)
        ns["key"] = jr.PRNGKey({"and": 11, "avg": 12, "or": 13}[mode])
        ns["x_t"] = jnp.zeros((512, 2))
        jr.LOG.clear(); jr.DRAWS.clear()
        exec(src, ns)
        n = ns["n"]
        kinds = [k for k, *_ in jr.LOG]
        # draw 0: x0 (normal); per step the same ikey feeds vector_field twice (one randint each)
        assert kinds == ["normal"] + ["randint"] * (2 * n), (kinds[:4], len(kinds))
        keys = [k for _, k, _ in jr.LOG]
        assert keys[1::2] == keys[2::2]
        x_gen = np.asarray(ns["x_gen"])
        ll1, ll2 = np.asarray(ns["ll_1"]), np.asarray(ns["ll_2"])
        out[f"toy_ode_{mode}"] = dict(x0_key=keys[0], step_keys=np.asarray(keys[1::2], dtype=np.int64), n=n, dt=float(ns["dt"]),
                                      bs=int(ns["bs"]), x0=x_gen[:, 0, :], x_final=x_gen[:, -1, :], x_quarters=x_gen[:, ::250, :],
                                      ll=np.stack([ll1[:, ::50], ll2[:, ::50]], -1))
        print(f"toy_ode_{mode}", "x rms", float(np.sqrt((x_gen[:, -1] ** 2).mean())), "ll range", float(ll1.min()), float(ll1.max()))
    # single calls of the two estimators on random fp64 inputs
    jax.config.update("jax_enable_x64", True)
    import jax.numpy as jnp
    ns = dict(jax=jax, jnp=jnp, np=np, random=jr, partial=None)
    exec(defs, ns); exec(sdll, ns); exec(kap, ns)
    rng = np.random.default_rng(5)
    B = 16
    ns["bs"] = B
    t = 0.37 * jnp.ones((B, 1))
    x, dx, s1, s2 = (f32r(rng.standard_normal((B, 2))) for _ in range(4))
    jr.DRAWS.clear()
    kappa = ns["select_kappa"](99, t, 1e-3, x, s1, s2)
    out["toy_calls"] = dict(t=0.37, dt=1e-3, x=x, dx=dx, s1=s1, s2=s2, eps=jr.DRAWS[0],
                            stoch_dll=np.asarray(ns["get_stoch_dll"](t, 1e-3, x, dx, s1)), select_kappa=np.asarray(kappa))
    return out


# ---------------------------------------------------------------------------------------------
# applications/images/clip_eval.py
# ---------------------------------------------------------------------------------------------

class _Stop(Exception):
    pass


def sd_unet_stub(x_in, t, emb):
    """Deterministic, smooth stand-in for UNet2DConditionModel (caller-supplied module in the product;
    tests/test_reference_vectors.py re-implements exactly this function in torch)."""
    phase = emb.mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
    return 0.8 * x_in * torch.cos(phase) + torch.sin(1.3 * x_in + phase + 1e-3 * float(t))


def sd(oracle_sigmas):
    torch.set_default_dtype(torch.float64)

    class _FromPretrained:
        @classmethod
        def from_pretrained(cls, *a, **kw):
            return cls()

        def to(self, *a, **kw):
            return self

    class VAE(_FromPretrained):
        config = types.SimpleNamespace(scaling_factor=0.18215)

        def decode(self, *a, **kw):
            raise _Stop()

    class Tok(_FromPretrained):
        model_max_length = 8

        def __call__(self, prompt, **kw):
            ids = torch.tensor([[sum(map(ord, p)) % 97 + j for j in range(8)] for p in prompt])
            return types.SimpleNamespace(input_ids=ids)

    class TextEnc(_FromPretrained):
        def __call__(self, ids):
            e = torch.sin(ids.double()[:, :, None] * torch.arange(1, 5).double()[None, None, :] * 0.37)
            return (e,)

    class UNet(_FromPretrained):
        config = types.SimpleNamespace(in_channels=4)

        def __call__(self, x, t, encoder_hidden_states=None):
            return types.SimpleNamespace(sample=sd_unet_stub(x, t, encoder_hidden_states))

    class Sched(_FromPretrained):
        def set_timesteps(self, n):
            sig, ts, init = oracle_sigmas(n)
            self.sigmas = torch.tensor(sig, dtype=torch.float64)
            self.timesteps = torch.tensor(ts, dtype=torch.float64)
            self.init_noise_sigma = init

    _stub("diffusers", AutoencoderKL=VAE, UNet2DConditionModel=UNet, EulerDiscreteScheduler=Sched)
    _stub("transformers", CLIPTextModel=TextEnc, CLIPTokenizer=Tok, CLIPProcessor=_FromPretrained, CLIPModel=_FromPretrained)
    _stub("ImageReward")
    _stub("wandb", log=lambda *a, **k: None, init=lambda *a, **k: None)
    _stub("matplotlib"); _stub("matplotlib.pyplot")
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("ref_clip_eval", os.path.join(REF, "applications/images/clip_eval.py"))
    ce = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ce)
    ce.torch_device = torch.device("cpu")
    ce.tqdm = lambda it, **kw: it
    real_manual_seed = torch.cuda.manual_seed
    out = {}
    for method, T, logp in (("and", 1.0, 0.0), ("or", 2.0, 0.1), ("avg", 1.0, 0.0), ("and_ode", 1.0, 0.0)):
        N, B = 12, 3
        args = types.SimpleNamespace(method=method, batch_size=B, num_inference_steps=N, seed=1, height=64, width=64,
                                     T=T, logp=logp, guidance_scale=7.5, obj="a cat", bg="a dog")
        zs, vels, divs, probes = [], [], [], []
        real_randn_like, real_get_vel, real_randint_like = torch.randn_like, ce.get_vel, torch.randint_like
        g = torch.Generator().manual_seed(77)

        def randn_like(x, g=g, zs=zs):
            z = torch.randn(x.shape, generator=g, dtype=torch.float32).to(x.dtype)   # fp32-representable draws
            zs.append(z.clone())
            return z

        def get_vel(t, sigma, latents, embeddings, *a, vels=vels, divs=divs, **kw):
            v, d = real_get_vel(t, sigma, latents, embeddings, *a, **kw)
            vels.append(v.clone())
            divs.append(d.clone())
            return v, d

        def randint_like(x, high, dtype=None, g=g, probes=probes):
            r = torch.randint(0, high, x.shape, generator=g).to(dtype or x.dtype)
            probes.append((r * 2 - 1).clone())
            return r

        lat_hist = []
        torch.randn_like, ce.get_vel, torch.randint_like = randn_like, get_vel, randint_like
        torch.cuda.manual_seed = lambda s: torch.Generator().manual_seed(s)
        try:
            try:
                ce.run(args)
            except _Stop as e:
                tb = e.__traceback__
                loc = None
                while tb is not None:
                    if tb.tb_frame.f_code.co_name == "run":
                        loc = tb.tb_frame.f_locals
                    tb = tb.tb_next
        finally:
            torch.randn_like, ce.get_vel, torch.randint_like = real_randn_like, real_get_vel, real_randint_like
            torch.cuda.manual_seed = real_manual_seed
        assert loc is not None
        kappa = loc["kappa"]
        kappa = kappa if torch.is_tensor(kappa) else torch.full((N + 1, B), float(kappa))
        sig = loc["scheduler"].sigmas if "scheduler" in loc else ce.scheduler.sigmas
        if method == "and_ode":
            # call order: vel_obj, vel_uncond (:354-355), then vel_obj+div, vel_bg+div, vel_uncond (:380-382)
            vel = torch.stack(vels).reshape(N, 5, B, 4, 8, 8)
            dv = torch.stack(divs).reshape(N, 5, B)
            out["sd_and_ode"] = dict(method=method, guidance=7.5, N=N, sigmas=sig.numpy(), timesteps=ce.scheduler.timesteps.numpy(),
                                     latents0=(loc["latents_og"] * ce.scheduler.init_noise_sigma).numpy(),
                                     probes=torch.stack(probes).numpy().astype(np.float32), v_obj=vel[:, 2].numpy(), v_bg=vel[:, 3].numpy(),
                                     v_unc=vel[:, 4].numpy(), dlog_obj=dv[:, 2].numpy(), dlog_bg=dv[:, 3].numpy(),
                                     ll_obj=loc["ll_obj"].numpy(), ll_bg=loc["ll_bg"].numpy(), kappa=kappa.numpy(),
                                     latents=loc["latents"].numpy(),
                                     emb_phase=np.array([float(e.mean()) for e in (loc["obj_embeddings"][:1], loc["uncond_embeddings"][:1],
                                                                                     loc["bg_embeddings"][:1])]))
            print("sd and_ode latents rms", float(loc["latents"].pow(2).mean().sqrt()), "kappa[-1]", kappa[-1].numpy())
            continue
        # call order inside the loop: vel_obj, vel_uncond, vel_bg  (clip_eval.py:354-355,394)
        vel = torch.stack(vels).reshape(N, 3, B, 4, 8, 8)
        out[f"sd_{method}"] = dict(method=method, T=T, logp=logp, guidance=7.5, N=N, sigmas=sig.numpy(),
                                   timesteps=ce.scheduler.timesteps.numpy(),
                                   latents0=(loc["latents_og"] * ce.scheduler.init_noise_sigma).numpy(),
                                   z=torch.stack(zs).numpy().astype(np.float32), v_obj=vel[:, 0].numpy(), v_unc=vel[:, 1].numpy(),
                                   v_bg=vel[:, 2].numpy(), ll_obj=loc["ll_obj"].numpy(), ll_bg=loc["ll_bg"].numpy(),
                                   kappa=kappa.numpy(), latents=loc["latents"].numpy(),
                                   emb_phase=np.array([float(e.mean()) for e in (loc["obj_embeddings"][:1], loc["uncond_embeddings"][:1],
                                                                                   loc["bg_embeddings"][:1])]))
        print("sd", method, "latents rms", float(loc["latents"].pow(2).mean().sqrt()), "kappa[-1]", kappa[-1].numpy())
    torch.set_default_dtype(torch.float32)
    return out


def save(fname, cases):
    flat = {}
    for cname, d in cases.items():
        for k, v in d.items():
            flat[f"{cname}/{k}"] = np.asarray(v)
    path = os.path.join(HERE, fname)
    np.savez_compressed(path, **flat)
    print(fname, os.path.getsize(path) // 1024, "KiB")


def main():
    which = set(sys.argv[1:]) or {"cifar", "scorenet", "toy", "sd"}
    jax, dynamics, eval_utils, mutils_ref, _ = import_cifar()
    if "cifar" in which:
        save("ref_cifar_steps.npz", cifar_steps(jax, dynamics))
        save("ref_cifar_loop.npz", cifar_loop(jax, dynamics, eval_utils))
        save("ref_cifar_ode.npz", cifar_ode_steps(jax, dynamics))
    if "scorenet" in which:
        save("ref_scorenet.npz", scorenet(jax, mutils_ref))
    if "toy" in which:
        save("ref_toy.npz", toy(jax))
    if "sd" in which:
        from oracle.schedule import edm_sigmas
        save("ref_sd.npz", sd(edm_sigmas))


if __name__ == "__main__":
    main()
