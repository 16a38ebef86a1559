"""SuperDiff OR / AND sampling loops over the fused sm_100a step kernels.

Entry points mirror the reference's three superposition loops:

* ``superdiff_or`` / ``superdiff_and`` — notebooks/superposition_edu.ipynb:797-822
  (stochastic OR) and :922-949 (stochastic AND): caller-supplied score models
  ``score_fn(t, x) -> sigma_t * grad log q_t(x)``, VP-SDE schedule, ``t`` accumulated
  in float32 like the notebook, log-density trajectories returned.
* ``sd_superdiff`` — applications/images/clip_eval.py:348-415 (methods and / or /
  avg): caller-supplied velocity function (the UNet stays PyTorch), EDM sigma table.
* ``SuperDiffSampler`` — the CIFAR loop (cifar/eval_utils.py:72-86 over
  cifar/dynamics.py:115-136) with the whole timestep (M score-net forwards + fused
  step) captured once into a CUDA graph and replayed per step; per-step scalars come
  from a device-side schedule table, so no host value is baked into the graph.
"""
import math

import torch

from . import _lib, ops, sde


def _noise_iter(noise, n, shape, device, seed):
    """noise: None (draw with torch.randn from `seed`), a [n, *shape] tensor (device or pinned host),
    or a callable i -> tensor."""
    if noise is None:
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        return lambda i: torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    if callable(noise):
        return noise
    if noise.shape[0] < n:
        raise ValueError("noise tensor has fewer steps than n_steps")
    return lambda i: noise[i].to(device, non_blocking=True)


def _vpsde_loop(score_fns, x0, mode, dlogq_mode, n_steps, dt, noise, seed, temperature, logp_bias, ito_scale,
                ll0, accumulate, record):
    _lib.require_device()
    M = len(score_fns)
    x = x0.clone().contiguous()
    B = x.shape[0]
    dev = x.device
    ll = torch.zeros(B, M, device=dev, dtype=torch.float32) if ll0 is None else ll0.to(dev, torch.float32).clone().contiguous()
    w = torch.empty(B, M, device=dev, dtype=torch.float32)
    ts = sde.time_grid(n_steps, dt, accumulate)
    nz = _noise_iter(noise, n_steps, x.shape, dev, seed)
    bias = None if logp_bias is None else torch.as_tensor(logp_bias, dtype=torch.float32, device=dev)
    traj = {"ll": [ll.clone()], "kappa": [], "x": [x.clone()]} if record else None
    for i in range(n_steps):
        t = float(ts[i])
        tt = torch.full((B, 1), t, device=dev, dtype=torch.float32)
        scores = [f(tt, x).contiguous() for f in score_fns]
        ops.step_vpsde(x, nz(i), scores, ll, sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t), dt, mode, dlogq_mode,
                       temperature=temperature, logp_bias=bias, ito_scale=ito_scale, x_out=x, weights=w)
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(w.clone()); traj["x"].append(x.clone())
    if record:
        traj = {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, w, traj


def superdiff_or(score_fns, x0, n_steps=1000, dt=1e-3, noise=None, seed=0, temperature=1.0, logp_bias=None,
                 ito_scale=None, accumulate="float32", record=False):
    """Stochastic SuperDiff-OR (superposition_edu.ipynb:797-822): kappa = softmax(T * ll) (:813, T = 1),
    ll_k(0) = 0 (:806-807), ll_k += get_stoch_dll (:818-819).  ``ito_scale`` defaults to the notebook's
    ndim * D (the ``ndim*dt*a`` term broadcast over D elements, :778)."""
    D = x0[0].numel()
    ito = float(D * D) if ito_scale is None else float(ito_scale)
    return _vpsde_loop(score_fns, x0, ops.MODE_OR, ops.DLOGQ_ITO, n_steps, dt, noise, seed, temperature, logp_bias,
                       ito, None, accumulate, record)


def superdiff_and(score_fns, x0, n_steps=1000, dt=1e-3, noise=None, seed=0, ito_scale=None, accumulate="float32",
                  record=False):
    """Stochastic SuperDiff-AND (superposition_edu.ipynb:922-949): kappa from select_kappa (:899-905; the
    general-M linear solve of SURVEY.md A.3 for M > 2, which the reference does not have),
    ll_k(0) = -|x0|^2/2 - ndim*log(2 pi) (:932), same noise inside kappa and dx (:900,943)."""
    D = x0[0].numel()
    ito = float(D * D) if ito_scale is None else float(ito_scale)
    M = len(score_fns)
    ll0 = (-0.5 * (x0.reshape(x0.shape[0], -1) ** 2).sum(1) - D * math.log(2 * math.pi))[:, None].expand(-1, M)
    return _vpsde_loop(score_fns, x0, ops.MODE_AND, ops.DLOGQ_ITO, n_steps, dt, noise, seed, 1.0, None, ito, ll0,
                       accumulate, record)


def superdiff_ode(score_fns, x0, mode="and", n_steps=1000, dt=1e-3, probes=None, seed=0, accumulate="float32", record=False):
    """Deterministic SuperDiff on the probability-flow ODE with Hutchinson divergence estimates -- the notebook's ODE cells
    (superposition_edu.ipynb: ``vector_field`` :242-247, ``get_dll`` / ``get_kappa`` :397-409, loops :426-447 (AND),
    :520-541 (kappa = 0.5), :633-656 (OR)), two caller-supplied differentiable score models.

    Per step: one Rademacher probe eps shared by both models (the notebook passes the same ``ikey`` to both calls);
    (s_k, J_k eps) by ``torch.func.jvp``; div_k = <J_k eps, eps>;
      and: kappa = [sigma (div_1 - div_2) + <s_1, s_1 - s_2>] / |s_1 - s_2|^2          (get_kappa)
      or:  kappa = softmax(ll)[0];      avg: kappa = 1/2
    dx = -dt (a x - b (s_2 + kappa (s_1 - s_2)));  ll_k += dt a ndim - dt b div_k + sum s_k/sigma (dx + dt (a x - b s_k))  (get_dll).
    ``probes``: [n_steps, *x0.shape] tensor of +-1 or None (drawn from ``seed``).  Returns (x, ll (B,2), kappa, traj)."""
    _lib.require_device()
    if len(score_fns) != 2:
        raise ValueError("the notebook's ODE superposition is defined for two models")
    x = x0.clone().contiguous()
    B, dev = x.shape[0], x.device
    D = x[0].numel()
    ll = torch.zeros(B, 2, device=dev, dtype=torch.float32)
    w = torch.full((B, 2), 0.5, device=dev, dtype=torch.float32)
    add = torch.empty(B, 2, device=dev, dtype=torch.float32)
    ts = sde.time_grid(n_steps, dt, accumulate)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    kmode = {"and": ops.MODE_FIXED, "or": ops.MODE_OR, "avg": ops.MODE_AVG}[mode]
    traj = {"ll": [ll.clone()], "kappa": [], "x": [x.clone()]} if record else None
    for i in range(n_steps):
        t = float(ts[i])
        a, b, sig = sde.dlog_alphadt(t), sde.beta(t), sde.sigma(t)
        eps = (probes[i].to(dev, torch.float32) if probes is not None else
               (torch.randint(0, 2, x.shape, generator=g, device=dev, dtype=torch.int32) * 2 - 1).to(torch.float32)).contiguous()
        tt = torch.full((B, 1), t, device=dev, dtype=torch.float32)
        scores, divs = [], []
        for k, f in enumerate(score_fns):
            s_k, j_k = torch.func.jvp(lambda _x: f(tt, _x), (x,), (eps,))
            scores.append(s_k.contiguous())
            divs.append(ops.rowdot(j_k.contiguous(), eps))
            add[:, k] = dt * a * D - dt * b * divs[k]
        if mode == "and":
            d = (scores[0] - scores[1]).contiguous()
            kappa = (sig * (divs[0] - divs[1]) + ops.rowdot(scores[0], d)) / ops.rowdot(d, d)
            w[:, 0] = kappa
            w[:, 1] = 1.0 - kappa
        ops.step_vpsde_ode(x, scores, ll, a, b, sig, dt, kmode, ops.DLOGQ_ITO, temperature=1.0, dlogq_add=add, x_out=x, weights=w)
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(w[:, 0].clone()); traj["x"].append(x.clone())
    if record:
        traj = {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, w[:, 0].clone(), traj


def sd_superdiff(get_vel, latents0, method="and", num_inference_steps=50, guidance_scale=7.5, lift=0.0, T=1.0,
                 logp=0.0, kappa_avg=0.5, noise=None, seed=1, record=False):
    """Stable-Diffusion latent SuperDiff (clip_eval.py:348-415).  ``get_vel(t, sigma, latents, which)`` with
    which in {'obj', 'bg', 'uncond'} returns the velocity tensor (the reference's get_vel over its UNet,
    :89-105); ``latents0`` is the unit-variance draw of :329-333 and is scaled by init_noise_sigma (:340).
    Returns (latents, ll (B,2), kappa trajectory or last kappa, traj)."""
    _lib.require_device()
    if method == "and_ode":
        return _sd_and_ode(get_vel, latents0, num_inference_steps, guidance_scale, lift, noise, seed, record)
    mode = {"and": ops.MODE_AND, "or": ops.MODE_OR, "avg": ops.MODE_AVG}[method]
    sigmas, timesteps, init_sigma = sde.edm_sigmas(num_inference_steps)
    x = (latents0 * init_sigma).contiguous()
    B, dev = x.shape[0], x.device
    ll = torch.ones(B, 2, device=dev, dtype=torch.float32)                       # :348-349
    kappa = torch.full((B,), 0.5, device=dev, dtype=torch.float32)               # :308
    nz = _noise_iter(noise, num_inference_steps, x.shape, dev, seed)
    traj = {"ll": [ll.clone()], "kappa": [kappa.clone()]} if record else None
    for i in range(num_inference_steps):
        sigma, dsigma = float(sigmas[i]), float(sigmas[i + 1] - sigmas[i])       # :352-353
        t = float(timesteps[i])
        v_obj = get_vel(t, sigma, x, "obj").contiguous()
        v_unc = get_vel(t, sigma, x, "uncond").contiguous()
        v_bg = get_vel(t, sigma, x, "bg").contiguous()
        ops.step_edm_cfg(x, nz(i), v_obj, v_bg, v_unc, ll, sigma, dsigma, mode, guidance=guidance_scale,
                         lift_term=sigma * lift / num_inference_steps, temperature=T, logp=logp,
                         kappa_fixed=kappa_avg, latents_out=x, kappa_out=kappa)
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(kappa.clone())
    if record:
        traj = {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, kappa, traj


def _sd_and_ode(get_vel, latents0, num_inference_steps, guidance_scale, lift, probes, seed, record):
    """clip_eval.py:377-391 (method "and_ode"): Rademacher probe per step (:379), (vel_k, dlog_k) by forward-mode
    differentiation of the caller's velocity function (:97-103), fused kappa / latent / log-likelihood update."""
    sigmas, timesteps, init_sigma = sde.edm_sigmas(num_inference_steps)
    x = (latents0 * init_sigma).contiguous()
    B, dev = x.shape[0], x.device
    ll = torch.ones(B, 2, device=dev, dtype=torch.float32)
    kappa = torch.full((B,), 0.5, device=dev, dtype=torch.float32)
    dlog = torch.empty(B, 2, device=dev, dtype=torch.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    traj = {"ll": [ll.clone()], "kappa": [kappa.clone()]} if record else None
    for i in range(num_inference_steps):
        sigma, dsigma = float(sigmas[i]), float(sigmas[i + 1] - sigmas[i])
        t = float(timesteps[i])
        eps = (probes[i].to(dev, torch.float32) if probes is not None else
               (torch.randint(0, 2, x.shape, generator=g, device=dev, dtype=torch.int32) * 2 - 1).to(torch.float32)).contiguous()
        vels = []
        for k, which in enumerate(("obj", "bg")):
            v, jv = torch.func.jvp(lambda _x, which=which: get_vel(t, sigma, _x, which), (x,), (eps,))
            vels.append(v.contiguous())
            ops.rowdot(jv.contiguous(), eps, scale=-1.0, out=dlog, column=k)          # div = -(eps * jvp).sum  (:103)
        v_unc = get_vel(t, sigma, x, "uncond").contiguous()
        ops.step_edm_ode(x, vels[0], vels[1], v_unc, dlog, ll, sigma, dsigma, guidance=guidance_scale,
                         lift_term=lift / dsigma * sigma / num_inference_steps, latents_out=x, kappa_out=kappa)
        if record:
            traj["ll"].append(ll.clone()); traj["kappa"].append(kappa.clone())
    if record:
        traj = {k: torch.stack(v) for k, v in traj.items()}
    return x, ll, kappa, traj


def protein_superdiff_step(x_trans, scores, eps, ll, a_trans, beta_trans, beta_rots, dt, operator="AND", T=1.0, logp=0.0,
                           lift_trans=0.0, lift_rots=0.0):
    """Two-component (translations / rotations) SuperDiff mixing of two SE(3) diffusion models -- one timestep of the
    'composition' branch of CompositionDiffusion.latent_mixing (applications/proteins/superdiff/composition.py:483-531 with
    kappa_AND :378-420, kappa_OR :422-434 and compute_stoch_dll :333-358), up to the SE(3) update that consumes the result.

    Each component is the fused VP-SDE step kernel with sigma = 1 (the protein code works with true scores, not sigma * score):
      translations: drift f = a_trans * x (FrameDiff R3 diffuser: a = -b_t / 2), b = beta_trans, Ito constant ndim^2 * dt * a (:346)
      rotations:    no drift (a = 0; the state does not enter, dx is applied on SO(3) by the caller), b = beta_rots (:352-353)
    and the SAME noise ``eps`` in both (:483,516,519).  AND: kappa equalises the two models' density increments (the kernel's
    M = 2 closed form, identical to :405-415); a non-zero ``lift`` (= logp * normalised sigma_t / num_inference_steps, :417)
    is added as lift / (2 dt b |s_1 - s_2|^2) and the step re-run with that kappa.  OR: kappa = softmax([T (ll_1 + logp), T ll_2])[0].

    x_trans, eps and the four score tensors (dict keys 'pt', 'ft', 'pr', 'fr' = proteus / framediff x trans / rots) are
    [B, L, 3] float32 CUDA tensors; ll: [B, 4] float32 = (proteus trans, framediff trans, proteus rots, framediff rots), updated
    in place.  Reductions are per sample (the reference samples one protein at a time, B = 1, and sums over the whole tensor).
    Returns (dx_trans, dx_rots, kappa_trans [B], kappa_rots [B], ll)."""
    _lib.require_device()
    if operator not in ("AND", "OR"):
        raise ValueError("operator must be 'AND' or 'OR' (composition.py:195)")
    B = x_trans.shape[0]
    D = x_trans[0].numel()
    dev = x_trans.device
    x = x_trans.contiguous()
    zeros = torch.zeros_like(x)
    out = []
    for comp, (k1, k2), xin, a, b, lift, cols in (("trans", ("pt", "ft"), x, float(a_trans), float(beta_trans), float(lift_trans), (0, 1)),
                                                   ("rots", ("pr", "fr"), zeros, 0.0, float(beta_rots), float(lift_rots), (2, 3))):
        sc = [scores[k1].contiguous(), scores[k2].contiguous()]
        llc = ll[:, cols[0]:cols[1] + 1].contiguous()
        w = torch.empty(B, 2, device=dev, dtype=torch.float32)
        if operator == "OR":
            bias = torch.tensor([float(logp), 0.0], device=dev, dtype=torch.float32)
            xo, llc, w = ops.step_vpsde(xin, eps, sc, llc, a, b, 1.0, dt, ops.MODE_OR, ops.DLOGQ_ITO, temperature=float(T),
                                        logp_bias=bias, ito_scale=float(D * D), weights=w)
        elif lift == 0.0:
            xo, llc, w = ops.step_vpsde(xin, eps, sc, llc, a, b, 1.0, dt, ops.MODE_AND, ops.DLOGQ_ITO, ito_scale=float(D * D), weights=w)
        else:
            # kappa without the lift from a dry run on a copy of ll, then kappa += lift / (2 dt b |s_1 - s_2|^2) (:413,419)
            _, _, w = ops.step_vpsde(xin, eps, sc, llc.clone(), a, b, 1.0, dt, ops.MODE_AND, ops.DLOGQ_ITO, ito_scale=float(D * D), weights=w)
            d = (sc[0] - sc[1]).contiguous()
            div = 2.0 * dt * b * ops.rowdot(d, d)
            w[:, 0] += lift / div
            w[:, 1] = 1.0 - w[:, 0]
            xo, llc, w = ops.step_vpsde(xin, eps, sc, llc, a, b, 1.0, dt, ops.MODE_FIXED, ops.DLOGQ_ITO, ito_scale=float(D * D), weights=w)
        ll[:, cols[0]:cols[1] + 1] = llc
        out.append((xo - xin if comp == "trans" else xo, w[:, 0].clone()))
    return out[0][0], out[1][0], out[0][1], out[1][1], ll


class SuperDiffSampler:
    """CIFAR SuperDiff sampler with one CUDA graph per timestep.

    nets: list of bound score-nets (``ScoreNet(config).bind(params)``); mode 'or' reproduces
    get_joint_stoch_vf (cifar/dynamics.py:115-136, T = 1e6, max-subtracted dlogq), 'and' applies the
    notebook's kappa to image tensors (BASELINE config 3), 'avg' is get_avg_vf (:155-171).
    """

    def __init__(self, nets, batch, image_shape=(32, 32, 3), mode="or", n_steps=1000, dt=None, temperature=1e6,
                 labels=None, device=None, use_graph=True, multi_stream=True):
        _lib.require_device()
        self.nets = list(nets)
        self.M = len(self.nets)
        self.B = int(batch)
        self.shape = (self.B,) + tuple(image_shape)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.mode = {"or": ops.MODE_OR, "and": ops.MODE_AND, "avg": ops.MODE_AVG}[mode]
        self.dlogq_mode = {"or": ops.DLOGQ_CIFAR_MAXSUB, "and": ops.DLOGQ_ITO, "avg": ops.DLOGQ_NONE}[mode]
        self.temperature = float(temperature)
        self.n_steps = int(n_steps)
        self.dt = float(dt) if dt is not None else 1.0 / self.n_steps
        ts = sde.time_grid(self.n_steps, self.dt, "float64")                   # cifar/eval_utils.py:76,85
        self.sched = sde.schedule_table(ts, self.dt, self.device)
        self.counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        dev = self.device
        self.x = torch.zeros(self.shape, device=dev, dtype=torch.float32)
        self.noise = torch.zeros(self.shape, device=dev, dtype=torch.float32)
        self.logq = torch.zeros(self.B, self.M, device=dev, dtype=torch.float32)
        self.weights = torch.zeros(self.B, self.M, device=dev, dtype=torch.float32)
        self.scores = [torch.zeros(self.shape, device=dev, dtype=torch.float32) for _ in range(self.M)]
        self.labels = None if labels is None else labels.to(dev)
        self.use_graph = use_graph
        self.multi_stream = multi_stream
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.M)] if multi_stream else []
        self.graph = None
        self.launches_per_step = None
        # host-noise pipeline: the next step's noise crosses PCIe on a copy stream into one of two staging buffers while the
        # current step computes; step() then only does a device-to-device copy into the buffer the captured graph reads
        self._copy_stream = None
        self._stage = None
        self._stage_ready = None      # per staging buffer: H2D finished
        self._stage_free = None       # per staging buffer: the D2D that consumed it finished
        self._stage_next = 0
        self._staged = None
        self._steps_done = 0          # host mirror of the device step counter (guards the schedule table)

    def _step_body(self):
        # The M score-net forwards are independent: run them on M streams (forked from / joined into the current
        # stream, also under graph capture) so one model's memory-bound GroupNorm passes overlap the other's GEMMs.
        cur = torch.cuda.current_stream()
        if self.multi_stream and self.M > 1:
            for i, net in enumerate(self.nets):
                st = self._streams[i]
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    net(None, self.x, self.labels, sched=self.sched, step_counter=self.counter, out=self.scores[i])
            for st in self._streams:
                cur.wait_stream(st)
        else:
            for i, net in enumerate(self.nets):
                net(None, self.x, self.labels, sched=self.sched, step_counter=self.counter, out=self.scores[i])
        ops.step_vpsde(self.x, self.noise, self.scores, self.logq, 0.0, 0.0, 1.0, 0.0, self.mode, self.dlogq_mode,
                       temperature=self.temperature, x_out=self.x, weights=self.weights, sched=self.sched,
                       step_counter=self.counter)
        # saturating: replaying the graph past the last timestep repeats the last schedule row instead of reading beyond the table
        ops.counter_add(self.counter, 1, rows=self.n_steps)

    def capture(self):
        """Warm up once on a side stream (lazy init, allocator), then capture one timestep.  The warm-up really executes a
        timestep, so the sampler state (x, logq, weights, noise, step counter) is saved before and restored after: a caller
        that did ``reset(x0)`` first and lets ``step()`` capture lazily still starts from x0 at schedule row 0."""
        before = ops.launch_count()
        saved = [t.clone() for t in (self.x, self.logq, self.weights, self.counter)]
        done = self._steps_done
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self._step_body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.launches_per_step = ops.launch_count() - before
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step_body()
        for t, v in zip((self.x, self.logq, self.weights, self.counter), saved):
            t.copy_(v)
        self._steps_done = done

    def reset(self, x0=None, logq0=None):
        """Back to schedule row 0 with logq = 0 (cifar/eval_utils.py:74) or ``logq0``; ``x0`` replaces the state."""
        self.counter.zero_()
        self._steps_done = 0
        if logq0 is None:
            self.logq.zero_()
        else:
            self.logq.copy_(logq0)
        if x0 is not None:
            self.x.copy_(x0.reshape(self.shape))

    def prefetch(self, noise_host):
        """Start the host -> device copy of a later step's noise (pinned host tensor) on the copy stream; the next
        ``step()`` without a ``noise`` argument consumes it.  Call right after ``step()`` so the transfer overlaps compute."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [torch.empty_like(self.noise) for _ in range(2)]
            self._stage_ready = [torch.cuda.Event() for _ in range(2)]
            self._stage_free = [None, None]
        k = self._stage_next
        self._stage_next ^= 1
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[k] is not None:
                self._copy_stream.wait_event(self._stage_free[k])
            self._stage[k].copy_(noise_host, non_blocking=True)
            self._stage_ready[k].record(self._copy_stream)
        self._staged = k

    def step(self, noise=None):
        """One Euler-Maruyama timestep at the schedule row the device counter points to.  ``noise``: device or host tensor
        copied on the current stream; None: use the buffer filled by ``prefetch()`` (or whatever ``self.noise`` holds)."""
        if noise is not None:
            self.noise.copy_(noise, non_blocking=True)
        elif self._staged is not None:
            k, self._staged = self._staged, None
            cur = torch.cuda.current_stream()
            cur.wait_event(self._stage_ready[k])
            self.noise.copy_(self._stage[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._stage_free[k] = ev
        if self._steps_done >= self.n_steps:
            raise RuntimeError(f"SuperDiffSampler.step(): all {self.n_steps} timesteps of the schedule have been taken; "
                               "call reset() before sampling again")
        if self.launches_per_step is None:
            self.capture()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        self._steps_done += 1

    def sample(self, x0=None, noise=None, seed=0):
        """Run all n_steps from x0 (default N(0, I) from `seed`).  Returns (x, logq, weights)."""
        g = torch.Generator(device=self.device)
        g.manual_seed(int(seed))
        if self.launches_per_step is None:
            self.capture()
        if x0 is None:
            x0 = torch.randn(self.shape, generator=g, device=self.device, dtype=torch.float32)
        self.reset(x0)
        host_noise = torch.is_tensor(noise) and not noise.is_cuda and noise.is_pinned()
        if host_noise:      # pipelined: noise[i+1] crosses PCIe while step i computes
            if noise.shape[0] < self.n_steps:
                raise ValueError("noise tensor has fewer steps than n_steps")
            self.prefetch(noise[0])
            for i in range(self.n_steps):
                self.step()
                if i + 1 < self.n_steps:
                    self.prefetch(noise[i + 1])
            return self.x, self.logq, self.weights
        nz = _noise_iter(noise, self.n_steps, self.shape, self.device, seed + 1)
        for i in range(self.n_steps):
            self.step(nz(i))
        return self.x, self.logq, self.weights
