"""Vector fields of the SuperDiff sampling path — host-side mirror of the
reference's cifar/dynamics.py over the fused sm_100a step kernel.

``get_joint_stoch_vf`` (:100-137), ``get_avg_vf`` (:140-173) and ``get_vpsde``'s
vector field (:48-54) keep their names and the closure signature

    joint_vf(t, data=(x, logq), args={'key', 'labels', 'dt'[, 'noise']}) -> (dx, dlogq)

(increments, the caller adds them — SURVEY.md Appendix C.14).  Each closure also
carries ``.step(t, x, logq, args) -> (x_next, logq_next, weights)``, the fast
path ``get_generator`` uses: the kernel writes x + dx directly, so no increment
tensors are materialised.

Noise: the reference draws it from ``fold_in(key, t*10000)`` (:117,126).  Here
``args['noise']`` (caller supplied, same shape as x) is used when present;
otherwise it is drawn with ``torch.randn`` from a generator seeded by
``(key, round(t*10000))`` — the same "pure function of (key, t)" contract.
"""
import torch

from . import ops, sde
from .models import utils as mutils


def _seed_of(key):
    if isinstance(key, torch.Generator):
        return key.initial_seed()
    if torch.is_tensor(key):
        return int(key.reshape(-1)[0].item())
    return int(key)


def _noise_for(args, t, x):
    noise = args.get("noise") if isinstance(args, dict) else None
    if noise is not None:
        return noise
    g = torch.Generator(device=x.device)
    g.manual_seed((_seed_of(args["key"]) * 1_000_003 + int(round(float(t) * 10_000))) % (2 ** 63 - 1))
    return torch.randn(x.shape, generator=g, device=x.device, dtype=torch.float32)


def _nets(models, states):
    return [mutils.get_model_fn(models[i], states[i].params_ema, train=False) for i in range(len(models))]


def _scores(nets, t, x, labels):
    tt = torch.full((1,), float(t), device=x.device, dtype=torch.float32)
    return [net(tt, x, labels) for net in nets]


def _make(nets, mode, dlogq_mode, temperature, n_models, ode=False):
    def step(t, x, logq, args, x_out=None, weights=None):
        dt = float(args["dt"])
        scores = _scores(nets, t, x, args.get("labels"))
        if ode:
            # probability-flow form -dt*(a x - b s) (cifar/dynamics.py:52,165): the noise-free kernel entry
            return ops.step_vpsde_ode(x, scores, logq, sde.dlog_alphadt(t), sde.beta(t), float(t) + 1e-3, dt, mode,
                                      dlogq_mode, temperature=temperature, x_out=x_out, weights=weights)
        noise = _noise_for(args, t, x)
        return ops.step_vpsde(x, noise, scores, logq, sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t), dt,
                              mode, dlogq_mode, temperature=temperature, x_out=x_out, weights=weights)

    def joint_vf(t, data, args):
        x, logq = data
        if logq is None or logq.shape[-1] != n_models:
            logq = torch.zeros(x.shape[0], n_models, device=x.device, dtype=torch.float32)
        lq = logq.clone()
        x_next, lq, _ = step(t, x, lq, args)
        return x_next - x, lq - logq

    joint_vf.step = step
    joint_vf.num_models = n_models
    # what eval_utils.get_generator needs to run this vector field as a replayed CUDA graph (SuperDiffSampler) instead of
    # ~280 eager launches per timestep: the bound repo-native score-nets and the step configuration.  Only the stochastic
    # fields qualify (the ODE fields draw Hutchinson probes / run JVPs per step), and only over repo-native nets.
    bound = [getattr(n, "bound", None) for n in nets]
    if not ode and all(b is not None and hasattr(b, "plan") for b in bound):
        joint_vf.sampler_spec = {"nets": bound, "mode": {ops.MODE_OR: "or", ops.MODE_AND: "and", ops.MODE_AVG: "avg"}[mode],
                                 "temperature": temperature, "noise_for": _noise_for}
    return joint_vf


def get_joint_stoch_vf(key, models, states, temperature=1e6):
    """SuperDiff-OR, stochastic (cifar/dynamics.py:100-137).  ``temperature`` is the
    reference's hard-coded 1e6 (:124)."""
    return _make(_nets(models, states), ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB, temperature, len(models))


def _probe_for(args, t, x, i):
    probes = args.get("probes") if isinstance(args, dict) else None
    if probes is not None:
        return probes[i]
    g = torch.Generator(device=x.device)
    g.manual_seed(((_seed_of(args["key"]) * 1_000_003 + int(round(float(t) * 10_000))) * 31 + i + 1) % (2 ** 63 - 1))
    return (torch.randint(0, 2, x.shape, generator=g, device=x.device, dtype=torch.int32) * 2 - 1).to(torch.float32)


def get_joint_vf(key, models, states, temperature=1e6):
    """SuperDiff-OR along the probability-flow ODE with Hutchinson divergence estimates (cifar/dynamics.py:59-97).

    Per model: a Rademacher probe eps_i (:83; ``args['probes']`` -- a list of M tensors -- overrides the draw made from
    (key, t, i)), (s_i, J_i eps_i) by forward-mode differentiation of the score model (:84), div_i = -b <J_i eps_i, eps_i>
    (:86); then one fused kernel: w = softmax(1e6 logq) (:88), dx = -dt sum_i w_i (a x - b s_i) (:89) and
    dlogq_i = dt div_i + sum s_i/(t + 1e-3) * (dx + dt (a x - b s_i)), minus the row max (:90-95)."""
    jvps = [mutils.get_model_jvp_fn(models[i], states[i].params_ema) for i in range(len(models))]
    M = len(models)

    def step(t, x, logq, args, x_out=None, weights=None):
        dt = float(args["dt"])
        tt = torch.full((1,), float(t), device=x.device, dtype=torch.float32)
        add = torch.empty(x.shape[0], M, device=x.device, dtype=torch.float32)
        scores = []
        for i, f in enumerate(jvps):
            eps = _probe_for(args, t, x, i).contiguous()
            s, js = f(tt, x, args.get("labels"), eps)
            scores.append(s.contiguous())
            ops.rowdot(js.contiguous(), eps, scale=-dt * sde.beta(t), out=add, column=i)      # dt * div_i
        return ops.step_vpsde_ode(x, scores, logq, sde.dlog_alphadt(t), sde.beta(t), float(t) + 1e-3, dt, ops.MODE_OR,
                                  ops.DLOGQ_CIFAR_MAXSUB, temperature=temperature, dlogq_add=add, x_out=x_out, weights=weights)

    def joint_vf(t, data, args):
        x, logq = data
        if logq is None or logq.shape[-1] != M:
            logq = torch.zeros(x.shape[0], M, device=x.device, dtype=torch.float32)
        lq = logq.clone()
        x_next, lq, _ = step(t, x, lq, args)
        return x_next - x, lq - logq

    joint_vf.step = step
    joint_vf.num_models = M
    return joint_vf


def get_joint_and_vf(key, models, states):
    """SuperDiff-AND on image tensors: the notebook's kappa (superposition_edu.ipynb:899-905)
    applied to the CIFAR state (BASELINE config 3; no counterpart in cifar/dynamics.py).
    logq accumulates the Ito increments *without* the D^2*dt*a constant (ito_scale = 0)."""
    nets = _nets(models, states)
    return _make(nets, ops.MODE_AND, ops.DLOGQ_ITO, 1.0, len(models))


def get_avg_vf(key, models, states, stoch=True):
    """Averaged vector field (cifar/dynamics.py:140-173); with one model this is the plain
    reverse SDE used by evaluate_fid (cifar/run_lib.py:145).  dlogq = 0 (:171)."""
    return _make(_nets(models, states), ops.MODE_AVG, ops.DLOGQ_NONE, 1.0, len(models), ode=not stoch)


def get_vpsde(config, model, train):
    """cifar/dynamics.py:15-56.  Only the schedule and the sampling-time pieces are provided;
    q_t / loss belong to training (out of scope)."""
    if train:
        raise NotImplementedError("training (DSM loss) is outside the sampling path")

    def q_t(key, data, t):
        g = key if isinstance(key, torch.Generator) else torch.Generator(device=data.device).manual_seed(int(key))
        eps = torch.randn(data.shape, generator=g, device=data.device, dtype=data.dtype)
        tt = torch.as_tensor(t, device=data.device, dtype=data.dtype)
        x_t = torch.exp(sde.log_alpha(tt)) * data + tt * eps
        return eps, x_t

    def loss(*a, **k):
        raise NotImplementedError("training (DSM loss) is outside the sampling path")

    vf_cache = {}

    def vector_field(t, data, args):
        # cifar/dynamics.py:48-54: dx = -dt*(a x - b s) with the raw (non-EMA) parameters, dlogq = zeros (B, 1)
        # the bound net is cached per parameter-tree object (ScoreNet.bound_for): an Euler loop over this field uploads the
        # weights once, not once per step
        params = args["state"].model_params
        if vf_cache.get("params") is not params:
            net = mutils.get_model_fn(model, params, train=False)
            vf_cache["params"], vf_cache["vf"] = params, _make([net], ops.MODE_AVG, ops.DLOGQ_NONE, 1.0, 1, ode=True)
        return vf_cache["vf"](t, (data[0], None), args)

    return q_t, loss, vector_field
