"""Sampling loop — host-side mirror of the reference's cifar/eval_utils.py:47-88.

``get_generator(models, config, vector_field, train=False)`` returns
``artifact_generator(key, labels) -> (x, n)`` exactly like the reference:
x0 ~ N(0, I), logq0 = 0, t0 = 1.0, dt = 5e-3, n = int(1/dt) Euler-Maruyama steps
with ``t`` advanced as a Python float (``t += -dt``, :76,85).  ``dt`` is exposed
because the BASELINE configs use 1000 steps.
"""
import torch

from . import distributed as dist_utils


def local_device_count():
    """The reference divides eval.batch_size by jax.local_device_count() (:48); here one
    process drives one GPU, so the analogue is the torch.distributed world size."""
    return dist_utils.world_size()


def get_generator(models, config, vector_field, train=False, dt=None, device=None, return_logq=False):
    shape = (config.eval.batch_size // local_device_count(), config.data.image_size,
             config.data.image_size, config.data.num_channels)
    n_models = len(models)
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    step_dt = dt if dt is not None else (1e-2 if train else 5e-3)

    def _x0(key):
        g = torch.Generator(device=device)
        seed = key.initial_seed() if isinstance(key, torch.Generator) else int(key)
        g.manual_seed(seed)
        return torch.randn(shape, generator=g, device=device, dtype=torch.float32), seed

    def artifact_generator(key, labels, state=None, x0=None, noise=None):
        """``x0`` / ``noise`` ([n, *shape]) override the draws made from ``key`` (parity tests feed the
        reference's recorded draws); the reference signature is (key, labels)."""
        x, seed = _x0(key)
        if x0 is not None:
            x = x0.to(device, torch.float32).reshape(shape).clone()
        t = 1.0
        n = int(t / step_dt)
        logq = torch.zeros(shape[0], n_models, device=device, dtype=torch.float32)
        args = {"key": seed + 1, "labels": labels, "dt": step_dt}
        if state is not None:
            args["state"] = state
        fast = getattr(vector_field, "step", None)
        for i in range(n):
            if noise is not None:
                args["noise"] = noise[i].to(device, torch.float32).reshape(shape).contiguous()
            if fast is not None:
                x, logq, _ = fast(t, x, logq, args, x_out=x)
            else:
                dx, dlogq = vector_field(t, (x, logq), args)
                x = x + dx
                logq = logq + dlogq
            t += -step_dt
        if return_logq:
            return x, n, logq
        return x, n

    if train:
        def train_generator(key, labels, state):
            return artifact_generator(key, labels, state)
        return train_generator
    return artifact_generator
