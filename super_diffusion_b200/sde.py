"""Host-side SDE schedule of the reference (cifar/dynamics.py:101-110, identical
copies at :18-27, :60-69, :141-150 and notebooks/superposition_edu.ipynb:82-93).

The per-step scalars (a_t, b_t, sigma_t, dt) are computed here on the host and
handed to the fused step kernel (by value, or as a device table for CUDA-graph
replay); sigma_t = t (the VP sigma is commented out in the reference).
"""
import numpy as np
import torch

BETA_0 = 0.1
BETA_1 = 20.0


def log_alpha(t):
    return -0.5 * t * BETA_0 - 0.25 * t ** 2 * (BETA_1 - BETA_0)


def dlog_alphadt(t):
    return -0.5 * BETA_0 - 0.5 * t * (BETA_1 - BETA_0)


def sigma(t):
    return t


def beta(t):
    return 1.0 + 0.5 * t * BETA_0 + 0.5 * t ** 2 * (BETA_1 - BETA_0)


def time_grid(n_steps, dt, accumulate="float64"):
    """Evaluation times of the reference loops, reproducing the accumulation dtype
    of ``t += -dt``: Python float in cifar/eval_utils.py:76,85, float32 array in
    notebooks/superposition_edu.ipynb:802,820."""
    if accumulate == "float64":
        t, out = 1.0, []
        for _ in range(n_steps):
            out.append(t)
            t += -dt
        return np.asarray(out, dtype=np.float64)
    if accumulate == "float32":
        t, d, out = np.float32(1.0), np.float32(dt), []
        for _ in range(n_steps):
            out.append(float(t))
            t = np.float32(t + (-d))
        return np.asarray(out, dtype=np.float64)
    raise ValueError(f"accumulate must be 'float64' or 'float32', got {accumulate!r}")


def schedule_table(ts, dt, device=None):
    """[n, 4] float32 rows (a_t, b_t, sigma_t, dt) for the kernels' `sched` argument."""
    ts = np.asarray(ts, dtype=np.float64)
    tab = np.stack([dlog_alphadt(ts), beta(ts), sigma(ts), np.full_like(ts, dt)], axis=1).astype(np.float32)
    t = torch.from_numpy(tab)
    return t.to(device) if device is not None else t


def edm_sigmas(num_inference_steps, beta_start=0.00085, beta_end=0.012, num_train_timesteps=1000):
    """Euler-discrete sigma table of the SD v1-4 scheduler config used at
    applications/images/clip_eval.py:43,339-340,351-353 (diffusers'
    EulerDiscreteScheduler, scaled_linear betas, 'linspace' spacing; restated,
    diffusers is not a dependency).  Returns (sigmas[N+1], timesteps[N], init_noise_sigma)."""
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=np.float32) ** 2
    alphas_cumprod = np.cumprod(1.0 - betas.astype(np.float64))
    sig_train = np.sqrt((1.0 - alphas_cumprod) / alphas_cumprod)
    timesteps = np.linspace(0, num_train_timesteps - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
    sig = np.interp(timesteps, np.arange(num_train_timesteps), sig_train)
    sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
    return sigmas, timesteps, float(sigmas.max())
