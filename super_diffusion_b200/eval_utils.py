"""Sampling loop — host-side mirror of the reference's cifar/eval_utils.py:47-88.

``get_generator(models, config, vector_field, train=False)`` returns
``artifact_generator(key, labels) -> (x, n)`` exactly like the reference:
x0 ~ N(0, I), logq0 = 0, t0 = 1.0, dt = 5e-3, n = int(1/dt) Euler-Maruyama steps
with ``t`` advanced as a Python float (``t += -dt``, :76,85).  ``dt`` is exposed
because the BASELINE configs use 1000 steps.
"""
import torch

from . import distributed as dist_utils
from .superposition import SuperDiffSampler


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

    samplers = {}

    def _graph_sampler(spec, n, labels):
        """One SuperDiffSampler (the whole timestep as one replayed CUDA graph) per (conditioned?) for this generator."""
        key = labels is not None
        smp = samplers.get(key)
        if smp is None:
            lab = None if labels is None else torch.zeros(shape[0], dtype=torch.int32, device=device)
            smp = SuperDiffSampler(spec["nets"], shape[0], image_shape=shape[1:], mode=spec["mode"], n_steps=n, dt=step_dt,
                                   temperature=spec["temperature"], labels=lab, device=device)
            smp.capture()
            samplers[key] = smp
        if labels is not None:
            smp.labels.copy_(labels.to(device=device, dtype=torch.int32))
        return smp

    def artifact_generator(key, labels, state=None, x0=None, noise=None, eager=False, logq_trace=None):
        """``x0`` / ``noise`` ([n, *shape] tensor or a callable i -> tensor; device or pinned host memory) override the draws made from ``key`` (parity tests feed the
        reference's recorded draws); the reference signature is (key, labels).  Vector fields built by this package over
        repo-native score-nets (``vector_field.sampler_spec``) run as a replayed CUDA graph; ``eager=True`` forces the
        per-launch path (same kernels, same noise, bit-identical results -- tests/test_loops_gpu.py).  ``logq_trace``: optional
        [n, B, M] float32 tensor (pinned host memory or device) that receives the log-densities after every step."""
        x, seed = _x0(key)
        if x0 is not None:
            x = x0.to(device, torch.float32).reshape(shape).clone()
        t = 1.0
        n = int(t / step_dt)
        args = {"key": seed + 1, "labels": labels, "dt": step_dt}
        if state is not None:
            args["state"] = state
        spec = getattr(vector_field, "sampler_spec", None)
        if spec is not None and state is None and not eager and shape[0] > 0:
            smp = _graph_sampler(spec, n, labels)
            smp.reset(x)
            for i in range(n):
                # same draw as the eager closure: args['noise'] or torch.randn seeded by (key, t) (dynamics._noise_for)
                if noise is not None:
                    smp.noise.copy_((noise(i) if callable(noise) else noise[i]).reshape(shape), non_blocking=True)
                else:
                    smp.noise.copy_(spec["noise_for"](args, t, smp.x))
                smp.step()
                if logq_trace is not None:
                    logq_trace[i].copy_(smp.logq, non_blocking=True)
                t += -step_dt
            if return_logq:
                return smp.x.clone(), n, smp.logq.clone()
            return smp.x.clone(), n
        logq = torch.zeros(shape[0], n_models, device=device, dtype=torch.float32)
        fast = getattr(vector_field, "step", None)
        for i in range(n):
            if noise is not None:
                args["noise"] = (noise(i) if callable(noise) else noise[i]).to(device, torch.float32).reshape(shape).contiguous()
            if fast is not None:
                x, logq, _ = fast(t, x, logq, args, x_out=x)
            else:
                dx, dlogq = vector_field(t, (x, logq), args)
                x = x + dx
                logq = logq + dlogq
            if logq_trace is not None:
                logq_trace[i].copy_(logq, non_blocking=True)
            t += -step_dt
        if return_logq:
            return x, n, logq
        return x, n

    if train:
        def train_generator(key, labels, state):
            return artifact_generator(key, labels, state)
        return train_generator
    return artifact_generator
