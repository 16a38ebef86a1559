"""Driver for `ncu --set full --import-source on -k regex:gemm_tcgen05 -s 3 -c 1`: a short-K 1x1 layer (the attention q'
projection shape: batch 512, 16x16, 256 -> 256 channels), which runs as operand-swapped 128-channel x 256-pixel units whose
epilogue is not hidden by its four K-blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
a = torch.randn(B, 16, 16, 256, device=dev).bfloat16()
w = (torch.randn(256, 256, device=dev) / 16).bfloat16()
bias = torch.zeros(256, device=dev)
for _ in range(4):
    ops.conv_gemm([(a, 1)], w, bias=bias)
torch.cuda.synchronize()
print("profile driver done")
