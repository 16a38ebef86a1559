"""jax.random stand-in: keys are Python ints, streams are NumPy generators
(NOT Threefry -- fixtures record the draws instead of relying on the stream)."""
import numpy as _np

from ._core import asjarr, default_float

LOG = []          # (kind, key, shape) of every draw, in call order
DRAWS = []        # the arrays themselves (cleared by the generator script between cases)


def _derive(key, *data):
    ss = _np.random.SeedSequence([int(key) & 0xFFFFFFFF, *[int(d) & 0xFFFFFFFF for d in data]])
    return int(ss.generate_state(1, dtype=_np.uint32)[0])


def PRNGKey(seed):
    return int(seed)


def split(key, num=2):
    return [_derive(key, 0x5711, i) for i in range(num)]


def fold_in(key, data):
    return _derive(key, 0xF01D, int(round(float(data))))


def normal(key, shape=(), dtype=None):
    # draws are float32-representable in either precision mode, so fp32 kernels can consume them exactly
    out = _np.random.default_rng(int(key)).standard_normal(tuple(shape)).astype(_np.float32).astype(dtype or default_float())
    LOG.append(("normal", int(key), tuple(shape)))
    DRAWS.append(out.copy())
    return asjarr(out)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    out = _np.random.default_rng(int(key)).uniform(minval, maxval, tuple(shape)).astype(dtype or default_float())
    return asjarr(out)


def randint(key, shape, minval, maxval, dtype=_np.int32):
    out = _np.random.default_rng(int(key)).integers(minval, maxval, tuple(shape)).astype(dtype)
    LOG.append(("randint", int(key), tuple(shape)))
    DRAWS.append(out.copy())
    return asjarr(out)
