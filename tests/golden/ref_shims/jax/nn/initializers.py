"""jax.nn.initializers stand-ins (shapes/scales per the JAX docs; NumPy streams)."""
import math

import numpy as _np

from .._core import asjarr, default_float


def zeros(key, shape, dtype=None):
    return asjarr(_np.zeros(shape, dtype or default_float()))


def ones(key, shape, dtype=None):
    return asjarr(_np.ones(shape, dtype or default_float()))


def normal(stddev=1e-2):
    def init(key, shape, dtype=None):
        return asjarr((stddev * _np.random.default_rng(int(key)).standard_normal(shape)).astype(dtype or default_float()))
    return init


def variance_scaling(scale, mode, distribution, in_axis=-2, out_axis=-1):
    def init(key, shape, dtype=None):
        rf = int(_np.prod(shape)) // (shape[in_axis] * shape[out_axis])
        fan_in, fan_out = shape[in_axis] * rf, shape[out_axis] * rf
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        var = scale / denom
        rng = _np.random.default_rng(int(key))
        if distribution == "uniform":
            lim = math.sqrt(3 * var)
            out = rng.uniform(-lim, lim, shape)
        elif distribution == "normal":
            out = math.sqrt(var) * rng.standard_normal(shape)
        else:  # truncated_normal
            out = math.sqrt(var) / 0.87962566103423978 * _np.clip(rng.standard_normal(shape), -2, 2)
        return asjarr(out.astype(dtype or default_float()))
    return init


def lecun_normal():
    return variance_scaling(1.0, "fan_in", "truncated_normal")
