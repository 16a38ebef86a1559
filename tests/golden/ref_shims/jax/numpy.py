"""jax.numpy stand-in: NumPy with JAX's default dtype (float32 unless
jax.config.update('jax_enable_x64', True)) and weakly-typed scalar results."""
import numpy as _np

from ._core import JArr, asjarr, default_float

ndarray = _np.ndarray
float32, float64, int32, int64, uint8 = _np.float32, _np.float64, _np.int32, _np.int64, _np.uint8
pi = _np.pi
newaxis = None


def _default(dtype):
    return default_float() if dtype is None else dtype


def zeros(shape, dtype=None):
    return asjarr(_np.zeros(shape, _default(dtype)))


def ones(shape, dtype=None):
    return asjarr(_np.ones(shape, _default(dtype)))


def zeros_like(x, dtype=None):
    return asjarr(_np.zeros_like(x, dtype=dtype))


def ones_like(x, dtype=None):
    return asjarr(_np.ones_like(x, dtype=dtype))


def arange(*a, dtype=None):
    out = _np.arange(*a, dtype=dtype)
    if dtype is None and out.dtype == _np.float64:
        out = out.astype(default_float())
    if dtype is None and out.dtype == _np.int64:
        out = out.astype(_np.int32)
    return asjarr(out)


def array(x, dtype=None):
    out = _np.array(x, dtype=dtype)
    if dtype is None and out.dtype == _np.float64 and not isinstance(x, _np.ndarray):
        out = out.astype(default_float())
    return asjarr(out)


asarray = array


def copy(x):
    return asjarr(_np.array(x, copy=True))


def expand_dims(x, axis):
    return asjarr(_np.expand_dims(x, axis))


def __getattr__(name):
    fn = getattr(_np, name)
    if not callable(fn):
        return fn

    def wrapped(*a, **kw):
        return asjarr(fn(*a, **kw))

    wrapped.__name__ = name
    return wrapped
