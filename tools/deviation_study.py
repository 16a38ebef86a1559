#!/usr/bin/env python
"""How far do the two score-net precisions drift from the reference arithmetic over a whole SuperDiff-OR trajectory?

VERDICT r01 "next" 1(c): the reference-default 200-step loop (cifar/eval_utils.py:75-77: dt = 5e-3) at batch >= 64 with the
real U-Net (two random-init score-nets, zero-init layers drawn at scale 1), SuperDiff-OR with the reference's T = 1e6 softmax
(cifar/dynamics.py:124), identical x0 and noise on every arm.

  phase 1 (CPU, no GPU needed):  python tools/deviation_study.py --oracle [--dtype float64|float32] [--batch 64]
      runs the oracle loop (oracle/scorenet.py + oracle/steps.py::or_step_cifar_literal) and writes
      tests/golden/cifar_or_200step_oracle_<dtype>.npz (log-density trajectory, weights trajectory, final samples; the fp64 file is
      also the fixture of tests/test_loops_gpu.py::test_cifar_or_200_steps_both_arms_against_the_fp64_oracle).
  phase 2 (GPU box):              python tools/deviation_study.py --gpu
      runs the B200 sampler in both precisions on the same inputs and prints / writes the deviation table
      (profiles/r02_deviation.json, .md).

x0 and the per-step noise come from a CPU torch.Generator with a fixed seed, so both phases (run on different machines)
see identical inputs.  The oracle is test infrastructure: this tool is a checker, not a product path.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SEED_X, SEED_NOISE, SEEDS_MODEL = 1234, 4321, (10, 11)


def inputs(B, n):
    g = torch.Generator().manual_seed(SEED_X)
    x0 = torch.randn(B, 32, 32, 3, generator=g)
    gn = torch.Generator().manual_seed(SEED_NOISE)
    noise = torch.randn(n, B, 32, 32, 3, generator=gn)
    return x0, noise


def models():
    from super_diffusion_b200.configs import vpsde
    from super_diffusion_b200.models import utils as mutils
    cfg = vpsde.get_config()
    out = [mutils.init_model(s, cfg, zero_init_scale=1.0) for s in SEEDS_MODEL]
    return cfg, [m for m, _ in out], [p for _, p in out]


def run_oracle(args):
    from oracle import scorenet as OS
    from oracle import steps as O
    torch.set_num_threads(args.threads)
    dt_ = torch.float64 if args.dtype == "float64" else torch.float32
    cfg, _, params = models()
    params = [OS.params_to(p, dtype=dt_) for p in params]
    B, n, dt = args.batch, args.steps, 1.0 / args.steps
    x0, noise = inputs(B, n)
    x, logq, t = x0.to(dt_), torch.zeros(B, 2, dtype=dt_), 1.0
    lq_tr, w_tr = [], []
    t0 = time.time()
    with torch.no_grad():
        for i in range(n):
            tt = torch.full((B, 1, 1, 1), t, dtype=dt_)
            s = torch.stack([OS.scorenet_apply(p, cfg, tt, x, None) for p in params])
            dx, dlogq, w = O.or_step_cifar_literal(x, logq, s, noise[i].to(dt_), t, dt)
            x = x + dx; logq = logq + dlogq; t += -dt
            lq_tr.append(logq.double().numpy().copy()); w_tr.append(w.double().numpy().copy())
            if i % 10 == 0:
                print(f"step {i}/{n}  {time.time() - t0:.0f}s", flush=True)
    path = os.path.join(ROOT, "tests", "golden", f"cifar_or_200step_oracle_{args.dtype}.npz")
    np.savez_compressed(path, logq=np.stack(lq_tr), weights=np.stack(w_tr), x=x.double().numpy(), B=B, n=n,
                        seconds=time.time() - t0, threads=args.threads)
    print("wrote", path)


def compare(name, lq, w, x, ref):
    """lq, w: [n, B, 2]; x: [B, 32, 32, 3]; ref: the oracle npz."""
    rl, rw, rx = ref["logq"], ref["weights"], ref["x"]
    n, B = rl.shape[0], rl.shape[1]
    win, rwin = w.argmax(-1), rw.argmax(-1)                     # [n, B] OR winners (T = 1e6: weights are one-hot up to ties)
    agree = (win == rwin)
    # the log-density difference between the two models is what the sampler acts on (the row max is subtracted every step)
    gap, rgap = lq[..., 0] - lq[..., 1], rl[..., 0] - rl[..., 1]
    scale = np.abs(rgap).max(axis=1) + 1e-30                    # per step
    gap_err = np.abs(gap - rgap).max(axis=1) / scale
    per_sample = np.linalg.norm((x - rx).reshape(B, -1), axis=1) / np.linalg.norm(rx.reshape(B, -1), axis=1)
    same_seq = agree.all(axis=0)
    out = {
        "arm": name,
        "winner_sequence_match_frac": float(same_seq.mean()),
        "per_step_winner_agreement_min": float(agree.mean(axis=1).min()),
        "per_step_winner_agreement_mean": float(agree.mean()),
        "logq_gap_rel_err_step1": float(gap_err[0]), "logq_gap_rel_err_step10": float(gap_err[min(9, n - 1)]),
        "logq_gap_rel_err_median": float(np.median(gap_err)), "logq_gap_rel_err_final": float(gap_err[-1]),
        "logq_gap_rel_err_max": float(gap_err.max()),
        "final_sample_rel_l2_median": float(np.median(per_sample)), "final_sample_rel_l2_max": float(per_sample.max()),
        "final_sample_rel_l2_median_same_winners": float(np.median(per_sample[same_seq])) if same_seq.any() else None,
        "final_sample_rel_l2_max_same_winners": float(per_sample[same_seq].max()) if same_seq.any() else None,
        "final_logq_gap_rel_err_same_winners": float((np.abs(gap[-1] - rgap[-1])[same_seq] / scale[-1]).max()) if same_seq.any() else None,
    }
    return out


def run_gpu(args):
    from super_diffusion_b200.superposition import SuperDiffSampler
    dev = torch.device("cuda:0")
    cfg, mods, params = models()
    refs = {}
    for d in ("float64", "float32"):
        p = os.path.join(ROOT, "tests", "golden", f"cifar_or_200step_oracle_{d}.npz")
        if os.path.exists(p):
            refs[d] = dict(np.load(p))
    if "float64" not in refs:
        raise SystemExit("run the --oracle phase first (tests/golden/cifar_or_200step_oracle_float64.npz)")
    B, n = int(refs["float64"]["B"]), int(refs["float64"]["n"])
    x0, noise = inputs(B, n)
    rows = []
    if "float32" in refs:
        r32 = refs["float32"]
        rows.append(compare("CPU oracle fp32 (reference working precision)", r32["logq"], r32["weights"], r32["x"], refs["float64"]))
    for prec in ("fp32", "bf16"):
        nets = [m.bind(p, dev, precision=prec) for m, p in zip(mods, params)]
        smp = SuperDiffSampler(nets, B, mode="or", n_steps=n, dt=1.0 / n, temperature=1e6, device=dev)
        smp.capture()
        smp.reset(x0.to(dev))
        lq_tr, w_tr = [], []
        for i in range(n):
            smp.step(noise[i].to(dev))
            lq_tr.append(smp.logq.double().cpu().numpy()); w_tr.append(smp.weights.double().cpu().numpy())
        torch.cuda.synchronize()
        rows.append(compare(f"B200 {prec} score-net", np.stack(lq_tr), np.stack(w_tr), smp.x.double().cpu().numpy(), refs["float64"]))
        del smp, nets
    out = {"batch": B, "steps": n, "temperature": 1e6, "truth": "CPU oracle fp64", "rows": rows}
    with open(os.path.join(ROOT, "profiles", "r02_deviation.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    keys = [k for k in rows[0] if k != "arm"]
    md = [f"SuperDiff-OR, {n} steps, batch {B}, T = 1e6, two random-init score-nets; truth = fp64 CPU oracle on identical x0 / noise\n",
          "| metric | " + " | ".join(r["arm"] for r in rows) + " |", "|---|" + "---|" * len(rows)]
    for k in keys:
        md.append(f"| {k} | " + " | ".join("n/a" if r[k] is None else f"{r[k]:.3g}" for r in rows) + " |")
    txt = "\n".join(md)
    open(os.path.join(ROOT, "profiles", "r02_deviation.md"), "w").write(txt + "\n")
    scratch = os.path.join(ROOT, "gpurun_out")          # the GPU box only sends gpurun_out/ back
    if os.path.isdir(scratch):
        import shutil
        for f in ("r02_deviation.md", "r02_deviation.json"):
            shutil.copy(os.path.join(ROOT, "profiles", f), os.path.join(scratch, f))
    print(txt)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    if a.oracle:
        run_oracle(a)
    if a.gpu:
        run_gpu(a)
