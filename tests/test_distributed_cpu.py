"""N > 1 host logic on CPU: world_size-2 gloo processes shard a batch, "sample" their shard with no
communication, and gather ragged shards back in order (super_diffusion_b200/distributed.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from super_diffusion_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert D.world_size() == world and D.rank() == rank
        lo, hi = D.shard_bounds(total)
        # every rank derives its shard deterministically from global sample indices (no per-step exchange)
        idx = torch.arange(lo, hi, dtype=torch.float32)
        x_local = torch.stack([idx, idx * 2, idx * idx], dim=1)
        logq_local = torch.stack([-idx, torch.zeros_like(idx)], dim=1)
        x = D.gather_samples(x_local, total)
        lq = D.gather_samples(logq_local, total)
        full = torch.arange(total, dtype=torch.float32)
        ok = torch.equal(x, torch.stack([full, full * 2, full * full], 1)) and torch.equal(lq[:, 0], -full)
        q.put((rank, bool(ok), tuple(x.shape)))
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2_gloo():
    for total in (7, 512):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert all(ok for _, ok, _ in res) and all(shape == (total, 3) for *_, shape in res)
