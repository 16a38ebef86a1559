#!/usr/bin/env python
"""Measured score-net error of both precision arms (bf16 / FP32-faithful) against the fp64 CPU oracle and the reference's own
ddpm.py output (tests/golden/ref_scorenet.npz): the numbers the parity gates in tests/ are set from (<= 1.5x these).
Run on the GPU box: python tools/precision_report.py > profiles/r02_precision_report.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import scorenet as OS                              # noqa: E402  (checker only)
from super_diffusion_b200.configs import vpsde                 # noqa: E402
from super_diffusion_b200.models import utils as mutils        # noqa: E402


def errs(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    return {"rel_l2": ((got - ref).norm() / ref.norm()).item(), "max_over_max": ((got - ref).abs().max() / ref.abs().max()).item()}


def main():
    dev = torch.device("cuda:0")
    out = {"oracle_fp64": [], "reference_vectors": []}
    for conditioned, B, t, seed in ((False, 8, 0.73, 3), (True, 16, 0.05, 3), (False, 3, 1.0, 3), (False, 64, 0.44, 9)):
        cfg = vpsde.get_config(conditioned=conditioned)
        model, params = mutils.init_model(seed, cfg, zero_init_scale=1.0)
        params = mutils.perturb_params(params, torch.Generator().manual_seed(seed + 100))
        g = torch.Generator().manual_seed(1)
        x = torch.randn(B, 32, 32, 3, generator=g)
        y = torch.randint(0, 10, (B,), generator=g) if conditioned else None
        with torch.no_grad():
            ref = OS.scorenet_apply(OS.params_to(params, dtype=torch.float64), cfg, torch.full((B, 1, 1, 1), t, dtype=torch.float64), x.double(), y)
            ref32 = OS.scorenet_apply(params, cfg, torch.full((B, 1, 1, 1), t), x, y)
        row = {"conditioned": conditioned, "B": B, "t": t, "oracle_fp32_cpu": errs(ref32, ref)}
        for prec in ("bf16", "fp32"):
            net = model.bind(params, dev, precision=prec)
            o = net(torch.full((B,), t), x.to(dev), y.to(dev) if y is not None else None)
            torch.cuda.synchronize()
            row[prec] = errs(o, ref)
        out["oracle_fp64"].append(row)
    from test_reference_vectors import _load, _our_params, _t
    for name, c in _load("ref_scorenet.npz").items():
        config, model, params = _our_params(c)
        row = {"case": name}
        for prec in ("bf16", "fp32"):
            net = model.bind(params, dev, precision=prec)
            o = net(_t(c["t"]).to(dev), _t(c["x"]).to(dev).contiguous(), _t(c["y"]).to(dev))
            torch.cuda.synchronize()
            row[prec] = errs(o, torch.from_numpy(c["out"]))
        out["reference_vectors"].append(row)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
