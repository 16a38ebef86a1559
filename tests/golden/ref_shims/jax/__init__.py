"""Minimal stand-in for the parts of JAX the reference's sampling path touches
(see ../README.md).  NumPy arrays, eager execution; `jit` is the identity,
`grad` / `jvp` go through CPU PyTorch autograd."""
import numpy as _np
import torch as _torch

from . import numpy  # noqa: F401  (jax.numpy)
from . import random  # noqa: F401
from . import nn  # noqa: F401
from . import image  # noqa: F401
from . import lax  # noqa: F401
from ._core import JArr, asjarr, default_float, config  # noqa: F401


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def vmap(fn, *a, **kw):
    raise NotImplementedError("vmap is outside the sampling path exercised by the fixtures")


def device_count():
    return 1


def local_device_count():
    return 1


def process_index():
    return 0


def value_and_grad(*a, **kw):
    raise NotImplementedError


def grad(fn):
    """d fn / d arg0 for scalar-valued fn built from arithmetic (torch autograd).
    A Python-float argument gives a Python float back (weakly typed, like JAX)."""

    def dfn(t):
        scalar = not isinstance(t, _np.ndarray)
        dt = _torch.float64 if scalar or t.dtype == _np.float64 else _torch.float32
        tt = _torch.tensor(_np.asarray(t), dtype=dt, requires_grad=True)
        out = fn(tt)
        (g,) = _torch.autograd.grad(out, tt)
        if scalar:
            return float(g)
        return asjarr(g.numpy().astype(t.dtype))

    return dfn


def jvp(fn, primals, tangents):
    """(fn(x), d/dh fn(x + h v)|_0) by central differences in float64 (h = 1e-6): exact to ~1e-10 for the smooth
    stand-in score models of the fixtures; jax.jvp itself is third-party (restated, not pinned)."""
    (x,), (v,) = primals, tangents
    x64, v64 = _np.asarray(x, dtype=_np.float64), _np.asarray(v, dtype=_np.float64)
    h = 1e-6
    out = fn(asjarr(_np.asarray(x)))
    d = (_np.asarray(fn(asjarr(x64 + h * v64)), dtype=_np.float64) - _np.asarray(fn(asjarr(x64 - h * v64)), dtype=_np.float64)) / (2 * h)
    return out, asjarr(d.astype(_np.asarray(out).dtype))
