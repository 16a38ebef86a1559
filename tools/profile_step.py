"""ncu driver for the fused step kernel alone: a few launches of one (B, M, mode) over rotating input sets.
    python tools/profile_step.py B M {or,and,avg} [launches]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
B, M = int(sys.argv[1]), int(sys.argv[2])
name = sys.argv[3]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
D = 3072
mode, dmode = {"or": (ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB), "and": (ops.MODE_AND, ops.DLOGQ_ITO),
               "avg": (ops.MODE_AVG, ops.DLOGQ_NONE)}[name]
torch.manual_seed(0)
sets = [dict(x=torch.randn(B, D, device=dev), xo=torch.empty(B, D, device=dev),
             sc=[torch.randn(B, D, device=dev) for _ in range(M)], nz=torch.randn(B, D, device=dev),
             lq=torch.zeros(B, M, device=dev), w=torch.zeros(B, M, device=dev)) for _ in range(2)]
for i in range(n):
    st = sets[i % 2]
    ops.step_vpsde(st["x"], st["nz"], st["sc"], st["lq"], -5.0, 5.0, 0.5, 1e-3, mode, dmode, temperature=1e6,
                   x_out=st["xo"], weights=st["w"])
torch.cuda.synchronize()
print("done", sets[0]["xo"].abs().mean().item())
