import numpy as _np

from .._core import asjarr
from . import initializers  # noqa: F401


def softmax(x, axis=-1):
    # jax.nn.softmax: exp(x - max) / sum(exp(x - max))
    x = _np.asarray(x)
    e = _np.exp(x - _np.max(x, axis=axis, keepdims=True))
    return asjarr(e / _np.sum(e, axis=axis, keepdims=True))


def sigmoid(x):
    return asjarr(1.0 / (1.0 + _np.exp(-_np.asarray(x))))


def swish(x):
    return asjarr(x * sigmoid(x))


silu = swish


def relu(x):
    return asjarr(_np.maximum(x, 0))


def elu(x, alpha=1.0):
    return asjarr(_np.where(x > 0, x, alpha * _np.expm1(_np.minimum(x, 0))))


def leaky_relu(x, negative_slope=0.01):
    return asjarr(_np.where(x >= 0, x, negative_slope * x))
