"""Score-net forward through the single native call (sd_scorenet_forward) next to the Python-driven op plan, eager and as a
CUDA-graph replay, at batch B (default 512).  CUDA events on the launching stream; plain text, not a bench result."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import native
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def ev_time(fn, iters=8, warm=3):
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
model, params = mutils.init_model(0, cfg, zero_init_scale=1.0)
bound = model.bind(params, dev)
net = native.NativeScoreNet(bound)
x = torch.randn(B, 32, 32, 3, device=dev)
t = torch.full((1,), 0.5, device=dev)
out_p, out_n = torch.empty_like(x), torch.empty_like(x)
flops = B * 12.154e9
ms_py = ev_time(lambda: bound(t, x, out=out_p))
ms_nat = ev_time(lambda: net(t, x, out=out_n))
assert torch.equal(out_p, out_n)


def graphed(fn):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


g_py, g_nat = graphed(lambda: bound(t, x, out=out_p)), graphed(lambda: net(t, x, out=out_n))
# alternate the two graphs (the chip drifts towards its power cap during a run: order matters for a 2 % difference)
rounds = [(ev_time(g_py.replay), ev_time(g_nat.replay)) for _ in range(4)]
print("graph replays, alternating (python, native) ms:", ", ".join(f"({a:.3f}, {b:.3f})" for a, b in rounds))
ms_gpy, ms_gnat = min(a for a, _ in rounds), min(b for _, b in rounds)
print(f"score-net forward, batch {B} (12.154 GFLOP / sample), workspace {net.workspace_bytes(B) / 2**30:.2f} GiB, weights {net.blob.numel() / 1e6:.1f} MB")
for name, ms in (("python op plan, eager", ms_py), ("sd_scorenet_forward, eager", ms_nat), ("python op plan, CUDA graph", ms_gpy),
                 ("sd_scorenet_forward, CUDA graph", ms_gnat)):
    print(f"  {name:34s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
