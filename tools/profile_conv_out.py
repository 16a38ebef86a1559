"""ncu driver for the output conv of the score network (3x3, 128 -> 3 channels, fp32 output; ddpm.py:99): two launches."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
torch.manual_seed(0)
a = torch.randn(B, 32, 32, 128, device=dev).bfloat16()
w = torch.zeros(16, 9 * 128, device=dev)
w[:3] = torch.randn(3, 9 * 128, device=dev) / math.sqrt(9 * 128)
w = w.bfloat16()
bias = torch.randn(3, device=dev)
out = torch.empty(B, 32, 32, 3, device=dev)
for _ in range(2):
    ops.conv_gemm([(a, 9)], w, bias=bias, out_f32=True, n_out=3, out=out)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    ops.conv_gemm([(a, 9)], w, bias=bias, out_f32=True, n_out=3, out=out)
e.record()
torch.cuda.synchronize()
print(f"conv_out B={B}: {s.elapsed_time(e) * 100:.1f} us / launch, |out| = {out.abs().mean().item():.4f}")
