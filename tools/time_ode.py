"""Timing probe (GPU box): score-net forward vs forward+JVP at batch B, and one deterministic SuperDiff step (2 models)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import dynamics, ops
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def ev(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


cfg = vpsde.get_config()
models, states = [], []
for seed in (1, 2):
    m, p = mutils.init_model(seed, cfg, zero_init_scale=1.0)
    models.append(m); states.append(mutils.State(params_ema=p))
net = models[0].bind(states[0].params_ema, dev)
x = torch.randn(B, 32, 32, 3, device=dev)
v = (torch.randint(0, 2, x.shape, device=dev) * 2 - 1).float()
print(f"B={B}: forward {ev(lambda: net(0.5, x)):.2f} ms, forward+jvp {ev(lambda: net.jvp(0.5, x, None, v)):.2f} ms")
vf = dynamics.get_joint_vf(0, models, states)
logq = torch.zeros(B, 2, device=dev)
args = {"key": 1, "labels": None, "dt": 1e-3}
print(f"deterministic SuperDiff step (2 models: 2 forward+jvp, 2 rowdot, ODE step kernel): {ev(lambda: vf.step(0.5, x, logq, args)):.2f} ms "
      f"-> {B / (1000 * ev(lambda: vf.step(0.5, x, logq, args)) * 1e-3):.2f} samples/s at 1000 steps")
st = dynamics.get_joint_stoch_vf(0, models, states)
print(f"stochastic SuperDiff step, eager: {ev(lambda: st.step(0.5, x, logq, args)):.2f} ms")
