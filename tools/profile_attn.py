"""Driver for `ncu --set full --import-source on -k regex:attn_core -s 2 -c 1`: the fused attention core at the bench shape
(batch 512, 16x16 tokens, 256 channels, residual + bias + GroupNorm sums), two warm-up launches and one profiled."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qk = torch.randn(B, 256, 512, device=dev).bfloat16()
vt = torch.randn(B, 256, 256, device=dev).bfloat16()
res = torch.randn(B, 256, 256, device=dev).bfloat16()
bias = torch.zeros(256, device=dev)
for _ in range(3):
    ops.attention_core(qk[:, :, :256], qk[:, :, 256:], vt, 256 ** -0.5, block=256, bias=bias, residual=res, want_stats=True, C=256)
torch.cuda.synchronize()
print("profile driver done")
