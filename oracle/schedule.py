"""SDE schedule of the reference (test infrastructure; see oracle/__init__.py).

Restates cifar/dynamics.py:101-110 (identical copies at :18-27, :60-69,
:141-150 and notebooks/superposition_edu.ipynb:82-93):

    log_alpha(t)   = -0.5*t*beta_0 - 0.25*t**2*(beta_1-beta_0)
    log_sigma(t)   = log(t)                      (the VP sigma is commented out)
    dlog_alphadt   = jax.grad(log_alpha)         = -0.5*beta_0 - 0.5*t*(beta_1-beta_0)
    beta(t)        = 1 + 0.5*t*beta_0 + 0.5*t**2*(beta_1-beta_0)
"""
import math

import numpy as np

BETA_0 = 0.1
BETA_1 = 20.0


def log_alpha(t):
    return -0.5 * t * BETA_0 - 0.25 * t ** 2 * (BETA_1 - BETA_0)


def dlog_alphadt(t):
    # d/dt of log_alpha, what jax.grad returns (cifar/dynamics.py:106)
    return -0.5 * BETA_0 - 0.5 * t * (BETA_1 - BETA_0)


def sigma(t):
    # exp(log_sigma(t)) with log_sigma = log(t)   (cifar/dynamics.py:105)
    return t


def beta(t):
    return 1.0 + 0.5 * t * BETA_0 + 0.5 * t ** 2 * (BETA_1 - BETA_0)


def time_grid(n_steps, dt, accumulate="float64"):
    """Times at which the vector field is evaluated, reproducing the reference's
    repeated ``t += -dt``.

    accumulate='float64': cifar/eval_utils.py:76,85 (Python float).
    accumulate='float32': notebooks/superposition_edu.ipynb:802,820 (``t`` is a
    float32 jnp array, so the drift after 999 steps is +0.93 %).
    """
    if accumulate == "float64":
        t = 1.0
        out = []
        for _ in range(n_steps):
            out.append(t)
            t += -dt
        return np.asarray(out, dtype=np.float64)
    if accumulate == "float32":
        t = np.float32(1.0)
        d = np.float32(dt)
        out = []
        for _ in range(n_steps):
            out.append(float(t))
            t = np.float32(t + (-d))
        return np.asarray(out, dtype=np.float64)
    raise ValueError(accumulate)


def edm_sigmas(num_inference_steps, beta_start=0.00085, beta_end=0.012,
               num_train_timesteps=1000):
    """diffusers EulerDiscreteScheduler sigma table for the SD v1-4 scheduler
    config (scaled_linear betas, 'linspace' timestep spacing), restated from
    its documented algorithm (diffusers is absent and
    its version is unpinned -> parity unpinned).  Call sites in the reference:
    applications/images/clip_eval.py:43,339-340,351-353.

    Returns (sigmas[N+1] float32 with trailing 0, timesteps[N] float32,
    init_noise_sigma).
    """
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                        dtype=np.float32) ** 2
    alphas_cumprod = np.cumprod(1.0 - betas.astype(np.float64))
    sig_train = np.sqrt((1.0 - alphas_cumprod) / alphas_cumprod)
    timesteps = np.linspace(0, num_train_timesteps - 1, num_inference_steps,
                            dtype=np.float32)[::-1].copy()
    sig = np.interp(timesteps, np.arange(num_train_timesteps), sig_train)
    sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
    init_noise_sigma = float(sigmas.max())  # 'linspace' spacing
    return sigmas, timesteps, init_noise_sigma


def log2pi():
    return math.log(2.0 * math.pi)
