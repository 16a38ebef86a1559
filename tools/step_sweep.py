"""Fused SuperDiff step kernel alone over the number of superposed models (BASELINE config 5: M = 2, 4, 8) and the batch.

For every (B, M, mode) the kernel runs over R independent input sets that together exceed twice the 126 MB L2, visited
round-robin inside one CUDA graph, so every launch reads all of its operands from HBM.  Reported: us per launch, algorithmic
GB/s = 4*B*D*(M+3) bytes / launch (read x, noise, M scores; write x') and the fraction of MEASURED_PEAKS.json's copy bandwidth.

    python tools/step_sweep.py [--batches 512 2048 8192] [--models 2 3 4 8]
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch                                        # noqa: E402
from super_diffusion_b200 import _lib, ops          # noqa: E402

D = 3072


def time_step(B, M, mode, dmode, dev, sched=False, shape=None):
    set_bytes = 4 * B * D * (M + 3)
    R = max(2, -(-2 * 126 * 1024 * 1024 // set_bytes) + 1)
    sets = []
    for _ in range(R):
        sets.append(dict(x=torch.randn(B, D, device=dev), xo=torch.empty(B, D, device=dev),
                         sc=[torch.randn(B, D, device=dev) for _ in range(M)], nz=torch.randn(B, D, device=dev),
                         lq=torch.zeros(B, M, device=dev), w=torch.zeros(B, M, device=dev)))

    # --sched: per-step scalars from the device-side schedule table (row = *counter), the form the CUDA-graph sampler uses
    table = torch.tensor([[-5.0, 5.0, 0.5, 1e-3]] * 8, device=dev) if sched else None
    counter = torch.full((1,), 3, dtype=torch.int32, device=dev) if sched else None

    def one(st):
        ops.step_vpsde(st["x"], st["nz"], st["sc"], st["lq"], -5.0, 5.0, 0.5, 1e-3, mode, dmode, temperature=1e6,
                       x_out=st["xo"], weights=st["w"], sched=table, step_counter=counter, launch_shape=shape)
    for st in sets:
        one(st)
    torch.cuda.synchronize()
    reps = max(1, 48 // R)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one(sets[0])
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for st in sets:
                one(st)
    ts = []
    for i in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(s.elapsed_time(e) * 1e3 / (reps * R))
    del g, sets
    torch.cuda.empty_cache()
    return statistics.median(ts), set_bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[512, 2048, 8192])
    ap.add_argument("--models", type=int, nargs="+", default=[2, 3, 4, 8])
    ap.add_argument("--modes", nargs="+", default=["or", "and", "avg"])
    ap.add_argument("--sched", action="store_true", help="scalars from the device schedule table (the sampler's form)")
    ap.add_argument("--shape", type=int, nargs=3, default=None, metavar=("THREADS", "NV", "CLUSTER"),
                    help="explicit launch shape through sd_step_vpsde_ex instead of the heuristic")
    args = ap.parse_args()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    peak = 6460.2
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    modes = (("or", ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB), ("and", ops.MODE_AND, ops.DLOGQ_ITO),
             ("avg", ops.MODE_AVG, ops.DLOGQ_NONE),
             # diagnostic combinations: log-density update without the softmax, softmax without the update
             ("avg_ito", ops.MODE_AVG, ops.DLOGQ_ITO), ("or_none", ops.MODE_OR, ops.DLOGQ_NONE))
    print(f"fused step kernel, D = {D}, fp32, HBM peak {peak:.1f} GB/s (measured copy)"
          + (", scalars from the device schedule table" if args.sched else ", scalars as arguments")
          + (f", launch shape {tuple(args.shape)}" if args.shape else ""))
    print(f"{'B':>6} {'M':>2} {'mode':>8} {'MB/launch':>10} {'us':>8} {'GB/s':>8} {'frac':>6}")
    for B in args.batches:
        for M in args.models:
            for name, md, dm in modes:
                if name not in args.modes:
                    continue
                us, by = time_step(B, M, md, dm, dev, sched=args.sched, shape=tuple(args.shape) if args.shape else None)
                gbs = by / (us * 1e-6) / 1e9
                print(f"{B:>6} {M:>2} {name:>8} {by / 1e6:>10.1f} {us:>8.2f} {gbs:>8.1f} {gbs / peak:>6.3f}", flush=True)


if __name__ == "__main__":
    main()
