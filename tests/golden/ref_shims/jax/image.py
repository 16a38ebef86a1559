import numpy as _np

from ._core import asjarr


def resize(x, shape, method):
    """jax.image.resize for the one use in the sampling path: 'nearest' integer
    upscaling (output pixel i samples input floor((i + 0.5) * in / out))."""
    if method != "nearest":
        raise NotImplementedError(method)
    x = _np.asarray(x)
    out = x
    for ax, (n_in, n_out) in enumerate(zip(x.shape, shape)):
        if n_in != n_out:
            idx = _np.floor((_np.arange(n_out) + 0.5) * n_in / n_out).astype(_np.int64)
            out = _np.take(out, idx, axis=ax)
    return asjarr(out)
