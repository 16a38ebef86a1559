"""Per-timestep SuperDiff math, restated on CPU (test infrastructure).

Two layers:

* ``*_literal`` functions transcribe the reference expression by expression
  (same operation order, dtype follows the inputs: pass float32 tensors for the
  reference's working precision, float64 for the truth the kernels are judged
  against).
* ``step_vpsde_gram`` / ``step_edm_gram`` restate the *same* update through the
  per-sample Gram reductions of SURVEY.md Appendix A.  This is the algebra the
  CUDA kernels implement; tests/test_oracle_identities.py proves in fp64 that
  it equals the literal forms, so the GPU parity tests may compare against
  either.

Scores are stacked ``(M, B, ...)`` like the reference's ``sscores``
(cifar/dynamics.py:123).  Noise is always caller supplied (the reference draws
it from a JAX key at cifar/dynamics.py:126, notebooks/superposition_edu.ipynb:816
and torch.randn_like at applications/images/clip_eval.py:395).
"""
import math

import torch

from . import schedule as S


def _flat(x):
    return x.reshape(x.shape[0], -1)


# ---------------------------------------------------------------------------
# cifar/dynamics.py
# ---------------------------------------------------------------------------

def or_step_cifar_literal(x, logq, sscores, eps, t, dt, temperature=1e6):
    """cifar/dynamics.py:123-136 (get_joint_stoch_vf.joint_vf), after the M
    score evaluations.  Returns (dx, dlogq, weights); dx has x's shape, dlogq
    and weights are (B, M).  ``temperature`` is hard-coded 1e6 at :124."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    M = sscores.shape[0]
    extra = (1,) * (x.dim() - 1)
    weights = torch.softmax(temperature * logq, dim=-1)                    # :124
    w = weights.T.reshape(M, -1, *extra)
    balanced = (w * sscores).sum(0)                                        # :125
    dx = -dt * (a * x - 2 * b * balanced) + math.sqrt(2 * sig * b * dt) * eps   # :127
    xs = x[None]
    dxs = dx[None]
    dlogq = a * (x + dx)[None] - (a * xs - 2 * b * sscores)                # :131
    dlogq = dlogq * (dt * (a * xs - 2 * b * sscores) + 2 * dxs + dt * a * (x + dx)[None])  # :132
    dlogq = dlogq / (4 * sig * b)                                          # :133
    dlogq = dlogq.reshape(M, x.shape[0], -1).sum(-1).T                     # :134
    dlogq = dlogq - dlogq.max(dim=1, keepdim=True).values                  # :135
    return dx, dlogq, weights


def ode_step_cifar_literal(x, logq, sscores, jvps, probes, t, dt, temperature=1e6):
    """cifar/dynamics.py:87-96 (get_joint_vf.joint_vf) after the M (score, jvp) evaluations: ``jvps[i]`` = J_i eps_i,
    ``probes[i]`` = eps_i (Rademacher, :83).  Returns (dx, dlogq (B, M), weights)."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    M = sscores.shape[0]
    red = tuple(range(2, sscores.dim()))
    vfs = a * x[None] - b * sscores                                        # :85
    dlogdx = sscores / (t + 1e-3)                                          # :85
    div = -b * (jvps * probes).reshape(M, x.shape[0], -1).sum(-1)          # :86
    weights = torch.softmax(temperature * logq, dim=-1)                    # :88
    w = weights.T.reshape(M, -1, *([1] * (x.dim() - 1)))
    dx = -dt * (w * vfs).sum(0)                                            # :89
    dlogq = -(-dt) * div + (dlogdx * (dx[None] - (-dt) * vfs)).sum(red)    # :90
    dlogq = dlogq.T                                                        # :91
    dlogq = dlogq - dlogq.max(dim=1, keepdim=True).values                  # :92
    return dx, dlogq, weights


def avg_step_cifar_literal(x, sscores, eps, t, dt, stoch=True):
    """cifar/dynamics.py:155-171 (get_avg_vf.joint_vf).  dlogq is zeros."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    M = sscores.shape[0]
    if stoch:
        vfs = a * x[None] - 2 * b * sscores                                # :163
    else:
        vfs = a * x[None] - b * sscores                                    # :165
    dx = -dt * vfs.mean(0)                                                 # :167
    if stoch:
        dx = dx + math.sqrt(2 * sig * b * dt) * eps                        # :170
    return dx, torch.zeros(x.shape[0], M, dtype=x.dtype)                   # :171


def single_ode_step_cifar_literal(x, sscore, t, dt):
    """cifar/dynamics.py:48-54 (get_vpsde.vector_field): probability-flow ODE
    step of a single model."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    dx = -dt * (a * x - b * sscore)
    return dx, torch.zeros(x.shape[0], 1, dtype=x.dtype)


# ---------------------------------------------------------------------------
# notebooks/superposition_edu.ipynb (toy, M = 2)
# ---------------------------------------------------------------------------

def stoch_dll_toy_literal(t, dt, x, dx, sscore, ndim=None):
    """notebooks/superposition_edu.ipynb:777-780 (get_stoch_dll).  ``ndim`` is
    the notebook's global (= 2, :72); the constant ndim*dt*a is broadcast over
    the D elements and then summed, i.e. contributes ndim*D*dt*a."""
    if ndim is None:
        ndim = x.shape[-1]
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    out = ndim * dt * a - dt * b * (sscore ** 2) / sig
    out = out + ((dx + dt * a * x) * sscore / sig)
    return out.sum(1)


def or_step_toy_literal(x, ll, s1, s2, eps, t, dt, ndim=None):
    """notebooks/superposition_edu.ipynb:813-819.  ``ll`` is (B, 2) holding
    (ll_1[:, i], ll_2[:, i]).  Returns (dx, dll (B,2), kappa (B,))."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    kappa = torch.softmax(torch.stack([ll[:, 0], ll[:, 1]]), dim=0)[0]     # :813
    k = kappa[:, None]
    dx = -dt * (a * x - 2 * b * (s2 + k * (s1 - s2)))                      # :815
    dx = dx + math.sqrt(2 * sig * b * dt) * eps                            # :816
    d1 = stoch_dll_toy_literal(t, dt, x, dx, s1, ndim)                     # :818
    d2 = stoch_dll_toy_literal(t, dt, x, dx, s2, ndim)                     # :819
    return dx, torch.stack([d1, d2], dim=1), kappa


def select_kappa_toy_literal(t, dt, x, s1, s2, eps):
    """notebooks/superposition_edu.ipynb:899-905 (select_kappa); ``eps`` is the
    standard normal draw behind ``noise`` (the same ikey is reused for dx at
    :943)."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    noise = math.sqrt(2 * sig * b * dt) * eps                              # :900
    dx_ind = -dt * (a * x - 2 * b * s2) + noise                            # :901
    kappa = -dt * b * (s1 - s2) * (s1 + s2) / sig                          # :902
    kappa = kappa + ((dx_ind + dt * a * x) * (s1 - s2) / sig)              # :903
    kappa = -kappa.sum(1) / (dt * 2 * b * (s1 - s2) ** 2 / sig).sum(1)     # :904
    return kappa


def and_step_toy_literal(x, s1, s2, eps, t, dt, ndim=None):
    """notebooks/superposition_edu.ipynb:938-946.  Returns (dx, dll (B,2), kappa)."""
    a = S.dlog_alphadt(t)
    b = S.beta(t)
    sig = S.sigma(t)
    kappa = select_kappa_toy_literal(t, dt, x, s1, s2, eps)                # :940
    k = kappa[:, None]
    dx = -dt * (a * x - 2 * b * (s2 + k * (s1 - s2)))                      # :942
    dx = dx + math.sqrt(2 * sig * b * dt) * eps                            # :943
    d1 = stoch_dll_toy_literal(t, dt, x, dx, s1, ndim)                     # :945
    d2 = stoch_dll_toy_literal(t, dt, x, dx, s2, ndim)                     # :946
    return dx, torch.stack([d1, d2], dim=1), kappa


def ll0_toy(x0, ndim=None):
    """notebooks/superposition_edu.ipynb:932: ll_0 = -0.5*|x0|^2 - ndim*log(2*pi)."""
    if ndim is None:
        ndim = x0.shape[-1]
    return -0.5 * (_flat(x0) ** 2).sum(1) - ndim * math.log(2 * math.pi)


# ---------------------------------------------------------------------------
# applications/images/clip_eval.py (Stable-Diffusion latents, EDM sigma)
# ---------------------------------------------------------------------------

def sd_step_literal(latents, z, v_obj, v_bg, v_unc, ll, sigma, dsigma, method,
                    guidance_scale=7.5, lift=0.0, num_inference_steps=50,
                    T=1.0, logp=0.0, kappa_avg=0.5):
    """applications/images/clip_eval.py:393-426 for method in {'and','or','avg'}.

    ``z`` is the standard normal draw (torch.randn_like at :395); ``ll`` is
    (B, 2) = (ll_obj[i], ll_bg[i]).  Returns (dx, ll_next (B,2), kappa (B,)).
    """
    red = tuple(range(1, latents.dim()))
    noise = math.sqrt(2 * abs(dsigma) * sigma) * z                          # :395
    g = guidance_scale
    if method == "and":
        dx_ind = 2 * dsigma * (v_unc + g * (v_bg - v_unc)) + noise          # :398
        kappa = (abs(dsigma) * (v_bg - v_obj) * (v_bg + v_obj)).sum(red) \
            - (dx_ind * (v_obj - v_bg)).sum(red) + sigma * lift / num_inference_steps   # :399
        kappa = kappa / (2 * dsigma * g * ((v_obj - v_bg) ** 2).sum(red))   # :400
    elif method == "or":
        kappa = torch.softmax(torch.stack([T * (ll[:, 0] + logp), T * ll[:, 1]]), 0)[0]  # :402
    elif method == "avg":
        kappa = torch.full((latents.shape[0],), kappa_avg, dtype=latents.dtype)          # :314
    else:
        raise ValueError(method)
    k = kappa.reshape(-1, *([1] * (latents.dim() - 1)))
    vf = v_unc + g * ((v_bg - v_unc) + k * (v_obj - v_bg))                  # :404
    dx = 2 * dsigma * vf + noise                                            # :405
    if method in ("and", "avg"):
        l_obj = ll[:, 0] + (-abs(dsigma) / sigma * v_obj ** 2 - dx * (v_obj / sigma)).sum(red)   # :409
        l_bg = ll[:, 1] + (-abs(dsigma) / sigma * v_bg ** 2 - dx * (v_bg / sigma)).sum(red)      # :410
    else:
        l_obj = ll[:, 0] - (v_obj * (dx + dsigma * v_obj) / sigma).sum(red)  # :412
        l_bg = ll[:, 1] - (v_bg * (dx + dsigma * v_bg) / sigma).sum(red)     # :413
    return dx, torch.stack([l_obj, l_bg], dim=1), kappa


def sd_and_ode_step_literal(latents, v_obj, v_bg, v_unc, dlog_obj, dlog_bg, ll, sigma, dsigma, guidance_scale=7.5, lift=0.0,
                            num_inference_steps=50):
    """applications/images/clip_eval.py:383-390 (method "and_ode") after the three get_vel calls: dlog_* are the Hutchinson
    divergences -(eps * jvp).sum (:103).  Returns (dx, ll_next (B,2), kappa (B,))."""
    red = tuple(range(1, latents.dim()))
    g = guidance_scale
    kappa = sigma * (dlog_obj - dlog_bg) + ((v_obj - v_bg) * (v_obj + v_bg)).sum(red) + lift / dsigma * sigma / num_inference_steps  # :383
    kappa = kappa - ((v_obj - v_bg) * (v_unc + g * (v_bg - v_unc))).sum(red)                                     # :384
    kappa = kappa / (g * ((v_obj - v_bg) ** 2).sum(red))                                                          # :385
    k = kappa.reshape(-1, *([1] * (latents.dim() - 1)))
    vf = v_unc + g * ((v_bg - v_unc) + k * (v_obj - v_bg))                                                        # :387
    dx = dsigma * vf                                                                                              # :388
    l_obj = ll[:, 0] + dsigma * (dlog_obj - ((-v_obj / sigma) * (v_obj - vf)).sum(red))                           # :389
    l_bg = ll[:, 1] + dsigma * (dlog_bg - ((-v_bg / sigma) * (v_bg - vf)).sum(red))                               # :390
    return dx, torch.stack([l_obj, l_bg], dim=1), kappa


# ---------------------------------------------------------------------------
# Gram-reduction restatement -- what the kernels compute
# ---------------------------------------------------------------------------

MODE_OR, MODE_AND, MODE_AVG, MODE_FIXED = 0, 1, 2, 3
DLOGQ_CIFAR_MAXSUB, DLOGQ_ITO, DLOGQ_NONE = 0, 1, 2


# ---------------------------------------------------------------------------
# applications/proteins/superdiff/composition.py -- two-component (translations / rotations) mixing
# ---------------------------------------------------------------------------

def protein_kappa_and_literal(s1, s2, eps, f_x, beta_t, dt, lift=0.0):
    """composition.py:378-420 (CompositionDiffusion.kappa_AND) for one component: s1 = 'proteus' score, s2 = 'framediff'
    score (true scores, no sigma factor), f_x = drift (r3_diffuser.drift_coef for translations, 0 for rotations), beta_t =
    diffusion_coef(t)^2 / 2, ``lift`` = logp * normalised sigma_t / num_inference_steps (:417).  Sums run over the whole
    tensor like the reference's ``.sum()`` (it samples one protein at a time).  Returns the scalar kappa."""
    s1, s2 = s1.double(), s2.double()                                   # :379-380
    noise = math.sqrt(2 * beta_t * dt) * eps                            # :405
    dx_ind = -dt * (f_x - 2 * beta_t * s2) + noise                      # :406
    d = s1 - s2                                                         # :408
    kappa = -dt * beta_t * d * (s1 + s2)                                # :410
    kappa = kappa + (dx_ind + dt * f_x) * d                             # :412
    kappa_div = (dt * 2 * beta_t * d ** 2).sum()                        # :413
    kappa = -(kappa / kappa_div).sum()                                  # :414-415
    return kappa + lift / kappa_div                                     # :419


def protein_stoch_dll_literal(score, dx, f_x, dlog_alphadt, beta_t, dt, component):
    """composition.py:333-358 (compute_stoch_dll) for one model and component; ndim = score.shape[1] * score.shape[2]."""
    ndim = score.shape[1] * score.shape[2]
    if component == "trans":
        out = ndim * dt * dlog_alphadt - dt * beta_t * score ** 2      # :346
        out = out + (dx + dt * f_x) * score                            # :347
    else:
        out = -dt * beta_t * score ** 2                                # :352
        out = out + dx * score                                         # :353
    return out.sum()


def protein_step_literal(x_trans, scores, eps, ll, a_trans, beta_trans, beta_rots, dt, operator, T=1.0, logp=0.0,
                         lift_trans=0.0, lift_rots=0.0):
    """One timestep of the 'composition' branch of CompositionDiffusion.latent_mixing (composition.py:483-531) up to the SE(3)
    update: scores = dict(pt=proteus trans, ft=framediff trans, pr=proteus rots, fr=framediff rots), all [1, L, 3];
    ll = (ll_proteus_trans, ll_framediff_trans, ll_proteus_rots, ll_framediff_rots) BEFORE the step; f_x = a_trans * x
    (FrameDiff's VP drift -b_t x / 2).  Returns (dx_trans, dx_rots, kappa_trans, kappa_rots, ll_next[4])."""
    f_x = a_trans * x_trans
    if operator == "AND":                                                              # :440-442
        kt = protein_kappa_and_literal(scores["pt"], scores["ft"], eps, f_x, beta_trans, dt, lift_trans)
        kr = protein_kappa_and_literal(scores["pr"], scores["fr"], eps, 0.0, beta_rots, dt, lift_rots)
    else:                                                                              # :422-434
        kt = torch.softmax(torch.stack([torch.as_tensor(T * (ll[0] + logp)), torch.as_tensor(T * ll[1])]).double(), 0)[0]
        kr = torch.softmax(torch.stack([torch.as_tensor(T * (ll[2] + logp)), torch.as_tensor(T * ll[3])]).double(), 0)[0]
    dx_trans = -dt * (f_x - 2 * beta_trans * (scores["ft"] + kt * (scores["pt"] - scores["ft"])))   # :514-515
    dx_trans = dx_trans + math.sqrt(2 * beta_trans * dt) * eps                                      # :516
    dx_rots = dt * 2 * beta_rots * (scores["fr"] + kr * (scores["pr"] - scores["fr"]))              # :518
    dx_rots = dx_rots + math.sqrt(2 * beta_rots * dt) * eps                                         # :519
    ll_next = [ll[0] + protein_stoch_dll_literal(scores["pt"], dx_trans, f_x, a_trans, beta_trans, dt, "trans"),   # :526-529
               ll[1] + protein_stoch_dll_literal(scores["ft"], dx_trans, f_x, a_trans, beta_trans, dt, "trans"),
               ll[2] + protein_stoch_dll_literal(scores["pr"], dx_rots, 0.0, 0.0, beta_rots, dt, "rots"),
               ll[3] + protein_stoch_dll_literal(scores["fr"], dx_rots, 0.0, 0.0, beta_rots, dt, "rots")]
    return dx_trans, dx_rots, kt, kr, ll_next


def and_kappa_general(G, N, dt, b, c):
    """General-M AND weights: equalise R_i.  NOT in the
    reference for M > 2 (it only has the M = 2 closed form); reduces to
    select_kappa for M = 2.  G: (B,M,M), N: (B,M)  ->  (B,M) float64."""
    B, M, _ = G.shape
    G = G.double()
    N = N.double()
    A = torch.zeros(B, M, M, dtype=torch.float64)
    rhs = torch.zeros(B, M, dtype=torch.float64)
    diag = torch.diagonal(G, dim1=1, dim2=2)
    for i in range(M - 1):
        A[:, i, :] = 2 * dt * b * (G[:, i, :] - G[:, M - 1, :])
        rhs[:, i] = dt * b * (diag[:, i] - diag[:, M - 1]) - c * (N[:, i] - N[:, M - 1])
    A[:, M - 1, :] = 1.0
    rhs[:, M - 1] = 1.0
    return torch.linalg.solve(A, rhs[..., None])[..., 0]


def step_vpsde_gram(x, eps, sscores, logq, a, b, sig, dt, mode, dlogq_mode,
                    temperature=1.0, logp_bias=None, fixed_weights=None, ito_const=None):
    """Gram-form restatement of the VP-SDE SuperDiff step in float64.

    Covers cifar/dynamics.py:123-136 (MODE_OR + DLOGQ_CIFAR_MAXSUB),
    :155-171 (MODE_AVG + DLOGQ_NONE), notebooks/superposition_edu.ipynb:813-819
    (MODE_OR + DLOGQ_ITO, temperature 1) and :899-905,938-946 (MODE_AND +
    DLOGQ_ITO).  Returns (x_next, logq_next, weights), all float64.

    ito_const: the model-independent constant added per step in DLOGQ_ITO mode;
    defaults to D*D*dt*a (the notebook's ndim*dt*a broadcast over D elements,
    SURVEY.md Appendix A.2).
    """
    M = sscores.shape[0]
    Bn = x.shape[0]
    xd = _flat(x).double()
    ed = _flat(eps).double()
    sd = sscores.reshape(M, Bn, -1).double()
    D = xd.shape[1]
    lq = logq.double()
    c = math.sqrt(2 * sig * b * dt)
    G = torch.einsum("ibd,jbd->bij", sd, sd)
    N = torch.einsum("ibd,bd->bi", sd, ed)
    if mode == MODE_OR:
        z = temperature * lq
        if logp_bias is not None:
            z = temperature * (lq + torch.as_tensor(logp_bias, dtype=torch.float64)[None])
        w = torch.softmax(z, dim=-1)
    elif mode == MODE_AND:
        w = and_kappa_general(G, N, dt, b, c)
    elif mode == MODE_AVG:
        w = torch.full((Bn, M), 1.0 / M, dtype=torch.float64)
    elif mode == MODE_FIXED:
        w = fixed_weights.double()
    else:
        raise ValueError(mode)
    mix = torch.einsum("bi,ibd->bd", w, sd)
    dx = -dt * a * xd + 2 * dt * b * mix + c * ed
    # R_i = [<dx,s_i> + dt*a*<x,s_i> - dt*b*G_ii]/sigma  (Appendix A.2)
    R = (2 * dt * b * torch.einsum("bj,bij->bi", w, G) + c * N
         - dt * b * torch.diagonal(G, dim1=1, dim2=2)) / sig
    if dlogq_mode == DLOGQ_CIFAR_MAXSUB:
        dl = R - R.max(dim=1, keepdim=True).values
    elif dlogq_mode == DLOGQ_ITO:
        k = (D * D * dt * a) if ito_const is None else ito_const
        dl = R + k
    elif dlogq_mode == DLOGQ_NONE:
        dl = torch.zeros_like(R)
    else:
        raise ValueError(dlogq_mode)
    return (xd + dx).reshape(x.shape), lq + dl, w


def step_edm_gram(latents, z, v_obj, v_bg, v_unc, ll, sigma, dsigma, mode,
                  guidance=7.5, lift_term=0.0, temperature=1.0, logp=0.0, kappa_fixed=0.5):
    """Gram-form restatement of applications/images/clip_eval.py:393-426 in
    float64.  ``lift_term`` = sigma*lift/num_inference_steps (:399).  Returns
    (latents_next, ll_next (B,2), kappa (B,))."""
    Bn = latents.shape[0]
    x = _flat(latents).double()
    zz = _flat(z).double()
    vo = _flat(v_obj).double()
    vb = _flat(v_bg).double()
    vu = _flat(v_unc).double()
    l = ll.double()
    g = guidance
    cn = math.sqrt(2 * abs(dsigma) * sigma)
    base = vu + g * (vb - vu)                 # vf at kappa = 0
    diff = vo - vb
    if mode == MODE_AND:
        num = abs(dsigma) * ((vb * vb).sum(1) - (vo * vo).sum(1)) \
            - (2 * dsigma * (base * diff).sum(1) + cn * (zz * diff).sum(1)) + lift_term
        kappa = num / (2 * dsigma * g * (diff * diff).sum(1))
    elif mode == MODE_OR:
        kappa = torch.softmax(torch.stack([temperature * (l[:, 0] + logp), temperature * l[:, 1]]), 0)[0]
    else:
        kappa = torch.full((Bn,), float(kappa_fixed), dtype=torch.float64)
    vf = base + g * kappa[:, None] * diff
    dx = 2 * dsigma * vf + cn * zz
    # NB the two reference conventions differ in the sign of the |v|^2 term
    # (dsigma < 0): and/avg use -|dsigma|/sigma (:409-410,423-424), or uses
    # -dsigma/sigma (:412-413).  SURVEY.md Appendix A.2 calls them identical;
    # they are not, so the mode selects the coefficient.
    q = (-dsigma / sigma) if mode == MODE_OR else (-abs(dsigma) / sigma)
    l_obj = l[:, 0] - (vo * dx).sum(1) / sigma + q * (vo * vo).sum(1)
    l_bg = l[:, 1] - (vb * dx).sum(1) / sigma + q * (vb * vb).sum(1)
    return (x + dx).reshape(latents.shape), torch.stack([l_obj, l_bg], 1), kappa
