"""Toy 2-D Gaussian-mixture score models and reference loops (test infrastructure).

The notebook trains two MLPs on the 'up' / 'down' mixtures of
notebooks/superposition_edu.ipynb:60-70 (component means 3*(k-0.5), std 0.4).
Training is outside the sampling path, so the fixtures use the *exact* scores of
those mixtures under the notebook's forward process q_t = N(alpha_t x_1, sigma_t^2)
(:95-98) as caller-supplied score models — smooth, deterministic and identical
in every dtype/device.  ``loop_*`` restate the sampling cells :797-822 (OR) and
:922-949 (AND) with the notebook's float32 time accumulation.
"""
import math

import torch

from . import schedule as S
from . import steps as O

MEANS = {
    "up": [(-1.5, 1.5), (1.5, 1.5)],      # randint([0,1],[2,2]) -> x in {0,1}, y = 1
    "down": [(-1.5, -1.5), (1.5, -1.5)],  # randint([0,0],[2,1]) -> x in {0,1}, y = 0
}
DATA_STD = 0.4


def mixture_sscore(which):
    """Returns score_fn(t, x) = sigma_t * grad_x log q_t(x) for the given mixture (t: (B,1) or scalar)."""
    means = MEANS[which]

    def fn(t, x):
        mu = torch.tensor(means, dtype=x.dtype, device=x.device)          # (K, 2)
        tt = t if torch.is_tensor(t) else torch.tensor(float(t), dtype=x.dtype, device=x.device)
        tt = tt.to(x.dtype).reshape(-1, 1) if tt.dim() > 0 else tt.reshape(1, 1)
        alpha = torch.exp(S.log_alpha(tt))                                 # (B|1, 1)
        var = (alpha * DATA_STD) ** 2 + tt ** 2                             # sigma_t = t
        diff = x[:, None, :] - alpha[:, None, :] * mu[None]                # (B, K, 2)
        logw = -0.5 * (diff ** 2).sum(-1) / var                            # (B, K)
        w = torch.softmax(logw, dim=1)
        grad = -(w[:, :, None] * diff).sum(1) / var                        # grad log q_t
        return tt * grad
    return fn


def loop_toy(score_fns, x0, noise, mode, n_steps=1000, dt=1e-3, accumulate="float32", record=False):
    """Free-running reference loop in the dtype of x0.  mode 'or' -> :797-822, 'and' -> :922-949."""
    ts = S.time_grid(n_steps, dt, accumulate)
    x = x0.clone()
    B = x.shape[0]
    if mode == "or":
        ll = torch.zeros(B, 2, dtype=x.dtype)
    else:
        ll = O.ll0_toy(x0)[:, None].expand(-1, 2).clone()
    traj = {"ll": [ll.clone()], "kappa": [], "x": [x.clone()]}
    for i in range(n_steps):
        t = float(ts[i])
        tt = torch.full((B, 1), t, dtype=x.dtype)
        s1, s2 = score_fns[0](tt, x), score_fns[1](tt, x)
        if mode == "or":
            dx, dll, kappa = O.or_step_toy_literal(x, ll, s1, s2, noise[i].to(x.dtype), t, dt, ndim=2)
        else:
            dx, dll, kappa = O.and_step_toy_literal(x, s1, s2, noise[i].to(x.dtype), t, dt, ndim=2)
        x = x + dx
        ll = ll + dll
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(kappa.clone()); traj["x"].append(x.clone())
    if record:
        return x, ll, {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, None


def loop_toy_ode(score_fns, x0, probes, mode, n_steps=1000, dt=1e-3, accumulate="float32", record=False):
    """notebooks/superposition_edu.ipynb ODE cells, restated: ``vector_field`` (:242-247: probe eps, (s, <J eps, eps>) by
    forward-mode differentiation), ``get_dll`` / ``get_kappa`` (:397-409) and the loops :426-447 (mode 'and'),
    :520-541 ('avg', kappa = 0.5), :633-656 ('or', kappa from the running log-densities).  ``probes[i]`` is the
    +-1 draw of step i, shared by both models (same ikey).  Dtype follows x0."""
    ts = S.time_grid(n_steps, dt, accumulate)
    x = x0.clone()
    B, ndim = x.shape
    ll = torch.zeros(B, 2, dtype=x.dtype)
    traj = {"ll": [ll.clone()], "kappa": [], "x": [x.clone()]}
    for i in range(n_steps):
        t = float(ts[i])
        a, b, sig = S.dlog_alphadt(t), S.beta(t), S.sigma(t)
        tt = torch.full((B, 1), t, dtype=x.dtype)
        eps = probes[i].to(x.dtype)
        out = [torch.func.jvp(lambda _x, f=f: f(tt, _x), (x,), (eps,)) for f in score_fns]
        s1, s2 = out[0][0], out[1][0]
        div1, div2 = (out[0][1] * eps).sum(1, keepdim=True), (out[1][1] * eps).sum(1, keepdim=True)
        if mode == "and":
            kappa = sig * (div1 - div2) + (s1 * (s1 - s2)).sum(1, keepdim=True)           # get_kappa
            kappa = kappa / ((s1 - s2) ** 2).sum(1, keepdim=True)
        elif mode == "or":
            mx = torch.maximum(ll[:, 0], ll[:, 1])
            e1, e2 = torch.exp(ll[:, 0] - mx), torch.exp(ll[:, 1] - mx)
            kappa = (e1 / (e1 + e2))[:, None]
        else:
            kappa = torch.full((B, 1), 0.5, dtype=x.dtype)
        dxdt = a * x - b * (s2 + kappa * (s1 - s2))

        def get_dll(s, div):
            v = a * x - b * s
            dlldt = -a * ndim + b * div
            return dlldt - ((s / sig) * (v - dxdt)).sum(1, keepdim=True)
        ll = torch.stack([ll[:, 0] - dt * get_dll(s1, div1).squeeze(1), ll[:, 1] - dt * get_dll(s2, div2).squeeze(1)], dim=1)
        x = x - dt * dxdt
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(kappa[:, 0].clone()); traj["x"].append(x.clone())
    if record:
        return x, ll, {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, None
