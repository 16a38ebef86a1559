"""Batch sharding across the GPUs of one box.

Samples are independent, so each rank runs every timestep on its own shard with
no per-step communication; one all-gather of the final samples and log-densities
(NCCL over NVLink on GPUs, gloo in the CPU tests) ends the job.  The reference's
counterpart is the communication-free ``jax.vmap(artifact_generator)`` of
cifar/run_lib.py:147,227.
"""
import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def shard_bounds(total, world=None, r=None):
    """Contiguous, near-even split of ``total`` samples: rank r owns [lo, hi)."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    base, rem = divmod(total, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def gather_samples(x_local, total):
    """All-gather ragged shards (sizes from shard_bounds) back into [total, ...] on every rank."""
    if world_size() == 1:
        return x_local
    world = world_size()
    sizes = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    pad[: x_local.shape[0]] = x_local
    out = torch.empty((world * mx,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
