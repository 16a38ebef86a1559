"""Fused EDM / CFG SuperDiff step on Stable-Diffusion latents (BASELINE config 4: 64x64x4 latents, batch 64) alone:
us per launch and algorithmic GB/s = 4*B*D*6 bytes (read latents, z, v_obj, v_bg, v_unc; write latents') over input sets that
rotate through > 2x the L2, launches captured in one CUDA graph.

    python tools/edm_sweep.py [--batches 64 512 2048] [--modes and or avg]
"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch                                        # noqa: E402
from super_diffusion_b200 import _lib, ops          # noqa: E402

D = 64 * 64 * 4


def time_edm(B, mode, dev):
    set_bytes = 4 * B * D * 6
    R = max(2, -(-2 * 126 * 1024 * 1024 // set_bytes) + 1)
    sets = [dict(x=torch.randn(B, D, device=dev), z=torch.randn(B, D, device=dev), vo=torch.randn(B, D, device=dev),
                 vb=torch.randn(B, D, device=dev), vu=torch.randn(B, D, device=dev), ll=torch.zeros(B, 2, device=dev),
                 xo=torch.empty(B, D, device=dev), k=torch.empty(B, device=dev)) for _ in range(R)]

    def one(st):
        ops.step_edm_cfg(st["x"], st["z"], st["vo"], st["vb"], st["vu"], st["ll"], 5.0, -0.2, mode, latents_out=st["xo"],
                         kappa_out=st["k"])
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
    ap.add_argument("--batches", type=int, nargs="+", default=[64, 512, 2048])
    ap.add_argument("--modes", nargs="+", default=["and", "or", "avg"])
    args = ap.parse_args()
    _lib.require_device()
    dev = torch.device("cuda", 0)
    peak = 6460.2
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    print(f"fused EDM step, D = {D}, fp32, HBM peak {peak:.1f} GB/s (measured copy)")
    print(f"{'B':>6} {'mode':>5} {'MB/launch':>10} {'us':>8} {'GB/s':>8} {'frac':>6}")
    for B in args.batches:
        for name in args.modes:
            us, by = time_edm(B, name, dev)
            gbs = by / (us * 1e-6) / 1e9
            print(f"{B:>6} {name:>5} {by / 1e6:>10.1f} {us:>8.2f} {gbs:>8.1f} {gbs / peak:>6.3f}", flush=True)


if __name__ == "__main__":
    main()
