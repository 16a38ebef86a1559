"""The notebook's toy score model (notebooks/superposition_edu.ipynb:157-173): ``MLP(num_hid=512, num_out=ndim)`` applied to
``hstack([t, x])`` -- four ``Dense(512) + swish`` layers and a linear head (0.79 M parameters for ndim = 2), wrapped as the
``get_sscore(state, t, x)`` callable (:773-774) the SuperDiff loops consume.

SURVEY.md §8 row a6: the score model stays a caller-supplied PyTorch module (1.58 MFLOP / sample: launch-bound, not a kernel
target); the per-step superposition math runs in the fused sm_100a step kernel.  Parameters use the Flax tree of the notebook
(``params['Dense_i']['kernel' | 'bias']``, kernel [in, out]) so a tree exported from the reference's ``train_model`` loads with
``MLP.from_flax``; ``init`` draws Flax's defaults (lecun-normal kernels, zero biases, :180-183).
"""
import math

import torch
from torch import nn

from . import utils


@utils.register_model(name="toy-mlp")
class MLP(nn.Module):
    def __init__(self, num_hid=512, num_out=2, ndim=2, config=None):
        super().__init__()
        if config is not None:                       # registry call shape: get_model(name)(config=config)
            num_hid = getattr(config.model, "num_hid", num_hid)
            num_out = ndim = getattr(config.data, "ndim", ndim)
        self.num_hid, self.num_out, self.ndim = num_hid, num_out, ndim
        dims = [ndim + 1] + [num_hid] * 4 + [num_out]
        self.layers = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(5)])

    def forward(self, t, x):
        """t: (B, 1) (or anything broadcastable to it), x: (B, ndim) -> (B, num_out) = sigma_t * grad log q_t(x)."""
        if not torch.is_tensor(t):
            t = torch.full((x.shape[0], 1), float(t), device=x.device, dtype=x.dtype)
        t = t.to(x.dtype).reshape(-1, 1).expand(x.shape[0], 1)
        h = torch.cat([t, x], dim=1)                 # jnp.hstack([t, x]) (:163)
        for lin in self.layers[:4]:
            h = nn.functional.silu(lin(h))           # nn.swish (:165-171)
        return self.layers[4](h)

    # -- Flax parameter tree <-> module ------------------------------------------------------------------------
    @classmethod
    def init(cls, rng, num_hid=512, num_out=2, ndim=2):
        """Flax nn.Dense defaults: kernel ~ lecun_normal (truncated normal, std = sqrt(1/fan_in) / .8796), bias = 0."""
        gen = rng if isinstance(rng, torch.Generator) else torch.Generator().manual_seed(int(rng))
        m = cls(num_hid, num_out, ndim)
        with torch.no_grad():
            for lin in m.layers:
                fan_in = lin.in_features
                std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
                w = torch.empty(lin.out_features, fan_in)
                nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
                lin.weight.copy_(w)
                lin.bias.zero_()
        return m

    @classmethod
    def from_flax(cls, params):
        """params: {'Dense_i': {'kernel': [in, out], 'bias': [out]}} (optionally nested under 'params')."""
        params = params.get("params", params)
        k0, k4 = torch.as_tensor(params["Dense_0"]["kernel"]), torch.as_tensor(params["Dense_4"]["kernel"])
        m = cls(num_hid=k0.shape[1], num_out=k4.shape[1], ndim=k0.shape[0] - 1)
        with torch.no_grad():
            for i, lin in enumerate(m.layers):
                p = params[f"Dense_{i}"]
                lin.weight.copy_(torch.as_tensor(p["kernel"]).T)
                lin.bias.copy_(torch.as_tensor(p["bias"]))
        return m

    def to_flax(self):
        return {f"Dense_{i}": {"kernel": lin.weight.detach().T.contiguous().clone(), "bias": lin.bias.detach().clone()}
                for i, lin in enumerate(self.layers)}


def get_sscore(model):
    """The notebook's ``get_sscore(state, t, x)`` (:773-774) with the state bound: score_fn(t, x), no autograd graph."""
    def score_fn(t, x):
        with torch.no_grad():
            return model(t, x)
    return score_fn
