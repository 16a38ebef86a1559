"""ncu driver for ONE conv GEMM launch (after a warm-up launch of the same shape).
    python tools/profile_one.py H C N taps [B] [gn]       (gn: conv + fused GroupNorm epilogue, sd_conv_gemm_gn)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
H, C, N, taps = (int(v) for v in sys.argv[1:5])
B = int(sys.argv[5]) if len(sys.argv) > 5 else 512
torch.manual_seed(0)
a = torch.randn(B, H, H, C, device=dev).bfloat16()
w = (torch.randn(N, taps * C, device=dev) / math.sqrt(taps * C)).bfloat16()
bias = torch.randn(N, device=dev)
rb = torch.randn(B, N, device=dev)
gn = (torch.ones(N, device=dev), torch.zeros(N, device=dev)) if len(sys.argv) > 6 and sys.argv[6] == "gn" else None
for _ in range(2):
    out = ops.conv_gemm([(a, taps)], w, rowbias=rb, want_stats=True, gn=gn) if gn else \
        ops.conv_gemm([(a, taps)], w, bias=bias, rowbias=rb, want_stats=True)
torch.cuda.synchronize()
print("done", out.float().abs().mean().item())
