"""Model registry / State / init_model / get_model_fn.

Host-side mirror of the reference's cifar/models/utils.py: ``State`` (:30-39),
``register_model`` / ``get_model`` (:42-65), ``init_model`` (:68-83) and
``get_model_fn`` (:86-96) keep their names, argument meaning and error
behaviour.  Parameters are a nested dict of torch tensors that uses the names
Flax's auto-naming gives the reference's modules and Flax kernel layouts
(conv HWIO, dense (in, out)), so a converted ``params_ema`` pytree drops in.
"""
import dataclasses
import math
from typing import Any

import torch


@dataclasses.dataclass
class State:
    """cifar/models/utils.py:30-39 (flax.struct.dataclass in the reference)."""
    step: int = 0
    opt_state: Any = None
    model_params: Any = None
    ema_rate: float = 0.9999
    params_ema: Any = None
    key: Any = None
    sampler_state: Any = None
    wandbid: Any = None

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)


_MODELS = {}


def register_model(cls=None, *, name=None):
    """A decorator for registering model classes (cifar/models/utils.py:45-61)."""

    def _register(cls):
        local_name = cls.__name__ if name is None else name
        if local_name in _MODELS:
            raise ValueError(f"Already registered model with name: {local_name}")
        _MODELS[local_name] = cls
        return cls

    if cls is None:
        return _register
    return _register(cls)


def get_model(name):
    return _MODELS[name]


# ---------------------------------------------------------------------------
# Parameter initialisation (cifar/models/layers.py:60-63 default_init,
# flax defaults for GroupNorm / Dense bias / Embed)
# ---------------------------------------------------------------------------

def _variance_scaling_uniform(gen, shape, fan_in, fan_out, scale):
    # jax.nn.initializers.variance_scaling(scale, 'fan_avg', 'uniform')
    scale = 1e-10 if scale == 0 else scale
    limit = math.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * limit


def _conv(gen, cin, cout, scale=1.0):
    return {"kernel": _variance_scaling_uniform(gen, (3, 3, cin, cout), 9 * cin, 9 * cout, scale),
            "bias": torch.zeros(cout)}


def _dense(gen, cin, cout, scale=1.0):
    return {"kernel": _variance_scaling_uniform(gen, (cin, cout), cin, cout, scale),
            "bias": torch.zeros(cout)}


def _nin(gen, cin, cout, scale=0.1):
    return {"W": _variance_scaling_uniform(gen, (cin, cout), cin, cout, scale),
            "b": torch.zeros(cout)}


def _gn(c):
    return {"scale": torch.ones(c), "bias": torch.zeros(c)}


def init_scorenet_params(gen, config, zero_init_scale=0.0):
    """Parameter tree of the 'score-net' (cifar/models/ddpm.py:47-101), faithful
    init.  ``zero_init_scale`` replaces the reference's ``init_scale=0.`` (-> 1e-10,
    layers.py:62) on conv_out / ResBlock conv2 / attention out-proj; 0.0 is the
    faithful value, a positive value gives the non-degenerate variant used by the
    parity fixtures."""
    m = config.model
    nf, ch_mult, nrb = m.nf, tuple(m.ch_mult), m.num_res_blocks
    attn_res = tuple(m.attn_resolutions)
    nres = len(ch_mult)
    C_img = config.data.num_channels
    p = {}
    p["Dense_0"] = _dense(gen, nf, nf * 4)
    p["Dense_1"] = _dense(gen, nf * 4, nf * 4)
    if m.conditioned:
        p["Embed_0"] = {"embedding": torch.randn(config.data.num_classes, nf * 4, generator=gen)
                        / math.sqrt(nf * 4)}
    p["Conv_0"] = _conv(gen, C_img, nf)
    counters = {"res": 0, "attn": 0, "down": 0, "up": 0}

    def res(cin, cout):
        blk = {"GroupNorm_0": _gn(cin), "Conv_0": _conv(gen, cin, cout),
               "Dense_0": _dense(gen, nf * 4, cout), "GroupNorm_1": _gn(cout),
               "Conv_1": _conv(gen, cout, cout, zero_init_scale)}
        if cin != cout:
            blk["NIN_0"] = _nin(gen, cin, cout)
        p[f"ResnetBlockDDPM_{counters['res']}"] = blk
        counters["res"] += 1

    def attn(c):
        p[f"AttnBlock_{counters['attn']}"] = {
            "GroupNorm_0": _gn(c), "NIN_0": _nin(gen, c, c), "NIN_1": _nin(gen, c, c),
            "NIN_2": _nin(gen, c, c), "NIN_3": _nin(gen, c, c, zero_init_scale)}
        counters["attn"] += 1

    size = config.data.image_size
    chans = [nf]
    c = nf
    for lvl in range(nres):
        for _ in range(nrb):
            res(c, nf * ch_mult[lvl])
            c = nf * ch_mult[lvl]
            if size in attn_res:
                attn(c)
            chans.append(c)
        if lvl != nres - 1:
            p[f"Downsample_{counters['down']}"] = {"Conv_0": _conv(gen, c, c)}
            counters["down"] += 1
            size //= 2
            chans.append(c)
    res(c, c)
    attn(c)
    res(c, c)
    for lvl in reversed(range(nres)):
        for _ in range(nrb + 1):
            res(c + chans.pop(), nf * ch_mult[lvl])
            c = nf * ch_mult[lvl]
        if size in attn_res:
            attn(c)
        if lvl != 0:
            p[f"Upsample_{counters['up']}"] = {"Conv_0": _conv(gen, c, c)}
            counters["up"] += 1
            size *= 2
    assert not chans
    p["GroupNorm_0"] = _gn(c)
    p["Conv_1"] = _conv(gen, c, C_img, zero_init_scale)
    return p


def perturb_params(params, gen, bias_std=0.05, gn_std=0.1):
    """Make every bias / GroupNorm affine non-trivial (test fixtures: a faithful
    init has all-zero biases and unit scales, which would hide indexing bugs)."""
    out = {}
    for k, v in params.items():
        if isinstance(v, dict):
            out[k] = perturb_params(v, gen, bias_std, gn_std)
        elif k in ("bias", "b"):
            out[k] = v + bias_std * torch.randn(v.shape, generator=gen)
        elif k == "scale":
            out[k] = v + gn_std * torch.randn(v.shape, generator=gen)
        else:
            out[k] = v
    return out


def count_params(params):
    if isinstance(params, dict):
        return sum(count_params(v) for v in params.values())
    return params.numel()


def init_model(rng, config, zero_init_scale=0.0):
    """Initialise a model (cifar/models/utils.py:68-83).  ``rng`` is an int seed
    or a torch.Generator (the reference takes a jax PRNGKey).  Returns
    ``(model, initial_params)`` like the reference."""
    model_name = config.model.name
    model = get_model(model_name)(config=config)
    gen = rng if isinstance(rng, torch.Generator) else torch.Generator().manual_seed(int(rng))
    initial_params = init_scorenet_params(gen, config, zero_init_scale)
    return model, initial_params


def get_model_fn(model, params, train=False):
    """cifar/models/utils.py:86-96.  Returns ``model_fn(t, x, y, rng=None)`` with
    t (B,1,1,1), x (B,H,W,C) NHWC, y (B,) int labels."""
    if train:
        raise NotImplementedError(
            "train=True (dropout) is outside the sampling path; the reference only "
            "uses it in the DSM loss (cifar/dynamics.py:36)")
    bound = model.bound_for(params) if hasattr(model, "bound_for") else model.bind(params)

    def model_fn(t, x, y, rng=None):
        return bound(t, x, y)

    model_fn.bound = bound      # lets get_generator drive repo-native nets through the CUDA-graph sampler
    return model_fn


def get_model_jvp_fn(model, params):
    """``jvp_fn(t, x, y, v) -> (model_fn(t, x, y), d/dh model_fn(t, x + h v, y)|_0)`` -- the pair jax.jvp returns at
    cifar/dynamics.py:84.  The B200 score-net has tangent kernels (``_Bound.jvp``); any other bound model that is a
    differentiable PyTorch callable goes through ``torch.func.jvp``."""
    bound = model.bound_for(params) if hasattr(model, "bound_for") else model.bind(params)
    if hasattr(bound, "jvp"):
        return lambda t, x, y, v: bound.jvp(t, x, y, v)

    def jvp_fn(t, x, y, v):
        return torch.func.jvp(lambda _x: bound(t, _x, y), (x,), (v,))
    return jvp_fn
