"""flax.linen stand-in.

Reproduces what the reference relies on:
  * modules are dataclasses built from class annotations (`NIN(C, init_scale=0.)`);
  * a submodule constructed inside a parent's `@nn.compact` method is named
    `<ClassName>_<k>` with one counter per class name and parent (Flax auto-naming), even
    when it is constructed by a free function called from that method (`ddpm_conv3x3`);
  * `self.param(name, init, shape)` creates (init) or looks up (apply) a leaf;
  * `model.init(rngs, *args)` -> {'params': tree}; `model.apply({'params': tree}, *args)`.

Layer semantics restated from the Flax 0.9 documentation (not pinned by the reference):
  Conv: NHWC activations, HWIO kernel, padding 'SAME' = lax SAME (total = max((ceil(n/s)-1)*s +
  (k-1)*d + 1 - n, 0), low = total // 2); Dense: x @ kernel + bias; GroupNorm: num_groups 32,
  epsilon 1e-6, statistics over (H, W, C/G) with var = max(E[x^2] - E[x]^2, 0), per-channel
  scale and bias; Embed: table lookup; Dropout(deterministic=True): identity.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from jax._core import asjarr, default_float
from jax.nn import swish, silu, relu, elu, leaky_relu, sigmoid, softmax  # noqa: F401
from jax.nn import initializers as _init

_STACK = []       # modules whose compact method is executing
_MODE = {"init": False, "key": 0}


def compact(fn):
    def wrapper(self, *a, **kw):
        self._counters = {}
        _STACK.append(self)
        try:
            return fn(self, *a, **kw)
        finally:
            _STACK.pop()
    wrapper.__name__ = fn.__name__
    return wrapper


def _fields(cls):
    names, defaults = [], {}
    for klass in reversed(cls.__mro__):
        if klass in (object, Module):
            continue
        for n in klass.__dict__.get("__annotations__", {}):
            if n not in names:
                names.append(n)
            if n in klass.__dict__:
                defaults[n] = klass.__dict__[n]
    return names, defaults


class Module:
    def __init__(self, *args, **kw):
        names, defaults = _fields(type(self))
        if len(args) > len(names):
            raise TypeError(f"{type(self).__name__}: too many positional arguments")
        vals = dict(defaults)
        vals.update(dict(zip(names, args)))
        name = kw.pop("name", None)
        kw.pop("parent", None)
        for k, v in kw.items():
            if k not in names:
                raise TypeError(f"{type(self).__name__}: unexpected field {k}")
            vals[k] = v
        for n in names:
            if n not in vals:
                raise TypeError(f"{type(self).__name__}: missing field {n}")
            object.__setattr__(self, n, vals[n])
        self._counters = {}
        self._params = None
        if _STACK:
            parent = _STACK[-1]
            cname = type(self).__name__
            k = parent._counters.get(cname, 0)
            parent._counters[cname] = k + 1
            self.name = name or f"{cname}_{k}"
            if _MODE["init"]:
                self._params = parent._params.setdefault(self.name, {})
            else:
                # parameter-free modules (Dropout) have no entry in a real pytree
                self._params = parent._params.get(self.name, {})
        else:
            self.name = name

    def param(self, name, init_fn, *init_args):
        if _MODE["init"]:
            if name not in self._params:
                _MODE["key"] += 1
                self._params[name] = np.asarray(init_fn(_MODE["key"], *init_args))
        elif name not in self._params:
            raise KeyError(f"missing parameter {self.name}/{name}")
        p = np.asarray(self._params[name])
        if init_args and tuple(p.shape) != tuple(init_args[0]):
            raise ValueError(f"{self.name}/{name}: shape {p.shape} != {tuple(init_args[0])}")
        return asjarr(p)

    def init(self, rngs, *args, **kw):
        _MODE["init"], _MODE["key"] = True, 1000
        self._params = {}
        try:
            self(*args, **kw)
        finally:
            _MODE["init"] = False
        return _Variables({"params": self._params})

    def apply(self, variables, *args, mutable=False, rngs=None, **kw):
        self._params = variables["params"]
        return self(*args, **kw)


class _Variables(dict):
    def pop(self, k):          # flax FrozenDict.pop returns (rest, value); models/utils.py:79 keeps the 2-tuple
        v = dict.pop(self, k)
        return v


def _same_pad(n, k, s, d):
    eff = (k - 1) * d + 1
    total = max((math.ceil(n / s) - 1) * s + eff - n, 0)
    return total // 2, total - total // 2


class Conv(Module):
    features: int
    kernel_size: tuple
    strides: tuple = (1, 1)
    padding: str = "SAME"
    use_bias: bool = True
    kernel_dilation: tuple = (1, 1)
    kernel_init: object = None
    bias_init: object = None

    @compact
    def __call__(self, x):
        x = np.asarray(x)
        kh, kw = self.kernel_size
        cin = x.shape[-1]
        kernel = self.param("kernel", self.kernel_init or _init.lecun_normal(), (kh, kw, cin, self.features))
        if self.padding != "SAME":
            raise NotImplementedError(self.padding)
        sh, sw = self.strides
        dh, dw = self.kernel_dilation
        ph, pw = _same_pad(x.shape[1], kh, sh, dh), _same_pad(x.shape[2], kw, sw, dw)
        xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        xt = F.pad(xt, (pw[0], pw[1], ph[0], ph[1]))
        wt = torch.from_numpy(np.ascontiguousarray(np.asarray(kernel))).to(xt.dtype).permute(3, 2, 0, 1)
        y = F.conv2d(xt, wt, None, stride=(sh, sw), dilation=(dh, dw)).permute(0, 2, 3, 1).numpy()
        if self.use_bias:
            y = y + np.asarray(self.param("bias", self.bias_init or _init.zeros, (self.features,))).astype(y.dtype)
        return asjarr(y)


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None

    @compact
    def __call__(self, x):
        x = np.asarray(x)
        kernel = self.param("kernel", self.kernel_init or _init.lecun_normal(), (x.shape[-1], self.features))
        y = x @ np.asarray(kernel).astype(x.dtype)
        if self.use_bias:
            y = y + np.asarray(self.param("bias", self.bias_init or _init.zeros, (self.features,))).astype(x.dtype)
        return asjarr(y)


class Embed(Module):
    num_embeddings: int
    features: int
    embedding_init: object = None

    @compact
    def __call__(self, ids):
        table = self.param("embedding", self.embedding_init or _init.variance_scaling(1.0, "fan_in", "normal", out_axis=0),
                           (self.num_embeddings, self.features))
        return asjarr(np.asarray(table)[np.asarray(ids)])


class GroupNorm(Module):
    num_groups: int = 32
    epsilon: float = 1e-6
    use_bias: bool = True
    use_scale: bool = True

    @compact
    def __call__(self, x):
        x = np.asarray(x)
        C = x.shape[-1]
        G = self.num_groups
        g = x.reshape(x.shape[0], -1, G, C // G)
        mean = g.mean(axis=(1, 3), keepdims=True)
        mean2 = (g * g).mean(axis=(1, 3), keepdims=True)
        var = np.maximum(mean2 - mean * mean, 0.0)
        y = ((g - mean) / np.sqrt(var + self.epsilon)).reshape(x.shape)
        if self.use_scale:
            y = y * np.asarray(self.param("scale", _init.ones, (C,))).astype(x.dtype)
        if self.use_bias:
            y = y + np.asarray(self.param("bias", _init.zeros, (C,))).astype(x.dtype)
        return asjarr(y)


class Dropout(Module):
    rate: float

    @compact
    def __call__(self, x, deterministic=True):
        if not deterministic and self.rate > 0 and not _MODE["init"]:
            raise NotImplementedError("dropout sampling is outside the sampling path (train=False)")
        return x


def avg_pool(x, window_shape, strides=None, padding="VALID"):
    x = np.asarray(x)
    xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
    return asjarr(F.avg_pool2d(xt, window_shape, strides or window_shape).permute(0, 2, 3, 1).numpy())


def max_pool(x, window_shape, strides=None, padding="VALID"):
    raise NotImplementedError
