"""Driver for `ncu --set full --import-source on -k regex:gemm_tcgen05 -s 4 -c 2`: the two launches with the largest share of a
forward at batch 512 -- conv1 of a 32x32 up block (9 x 256 -> 128 channels, fused GroupNorm + swish, global-memory exchange
between the four CTAs of an image) and conv1 of a 16x16 block (9 x 512 -> 256, cta_group::2 pairs, fused GroupNorm)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 512
a32 = torch.randn(B, 32, 32, 256, device=dev).bfloat16()
w32 = (torch.randn(128, 9 * 256, device=dev) / math.sqrt(9 * 256)).bfloat16()
a16 = torch.randn(B, 16, 16, 512, device=dev).bfloat16()
w16 = (torch.randn(256, 9 * 512, device=dev) / math.sqrt(9 * 512)).bfloat16()
for _ in range(3):
    o1 = ops.conv_gemm([(a32, 9)], w32, rowbias=torch.randn(B, 128, device=dev), want_stats=True,
                       gn=(torch.ones(128, device=dev), torch.zeros(128, device=dev)))
    o2 = ops.conv_gemm([(a16, 9)], w16, rowbias=torch.randn(B, 256, device=dev), want_stats=True,
                       gn=(torch.ones(256, device=dev), torch.zeros(256, device=dev)))
torch.cuda.synchronize()
assert o1.gn_fused and o2.gn_fused
print("profile driver done")
