"""Test infrastructure (lives under tests/ because it checks against oracle/): small driver for compute-sanitizer (memcheck /
racecheck) over the fused step kernels added in r01d: shared-memory AND,
streaming AND, OR with preloaded log-densities, warp kappa solve, streaming EDM step.  Tiny shapes; checks results too."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import schedule as S
from oracle import steps as O
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
t, dt = 0.4, 1e-3
f32 = lambda v: float(torch.tensor(v, dtype=torch.float32))
for M, D, shape in ((3, 3072, None), (8, 3072, None), (4, 1024, (128, 1, -2)), (5, 3072, (256, 1, -1)), (2, 3072, None),
                    (8, 3072, "or"), (3, 3072, "or")):
    B = 3
    x, e = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    s, lq = torch.randn(M, B, D, generator=g), torch.randn(B, M, generator=g)
    mode, dm = (ops.MODE_OR, ops.DLOGQ_CIFAR_MAXSUB) if shape == "or" else (ops.MODE_AND, ops.DLOGQ_ITO)
    xo, l2, w = ops.step_vpsde(x.to(dev), e.to(dev), [si.contiguous() for si in s.to(dev)], lq.to(dev).clone(), S.dlog_alphadt(t),
                               S.beta(t), S.sigma(t), dt, mode, dm, ito_scale=1.0,
                               launch_shape=None if shape in (None, "or") else shape)
    torch.cuda.synchronize()
    xr, lr, wr = O.step_vpsde_gram(x, e, s, lq, f32(S.dlog_alphadt(t)), f32(S.beta(t)), f32(S.sigma(t)), f32(dt), mode, dm,
                                   ito_const=f32(dt) * f32(S.dlog_alphadt(t)))
    assert torch.allclose(xo.cpu().double(), xr, rtol=1e-5, atol=2e-5)
    assert (w.cpu().double() - wr).abs().max() < 2e-4
B, D = 260, 1024
lat, z, vo, vb, vu = (torch.randn(B, D, generator=g).to(dev) for _ in range(5))
for mode in ("and", "or", "avg"):
    ops.step_edm_cfg(lat, z, vo, vb, vu, torch.ones(B, 2, device=dev), 3.2, -0.41, mode)
torch.cuda.synchronize()
print("sanitize driver ok")
