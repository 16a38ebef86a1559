"""Do a memory-bound GroupNorm pass and a tcgen05 GEMM on two streams overlap on the SMs?  (GPU box diagnostic)"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
B = 512
a = torch.randn(B, 32, 32, 128, device=dev).bfloat16()
w = (torch.randn(128, 9 * 128, device=dev) / math.sqrt(9 * 128)).bfloat16()
rb = torch.randn(B, 128, device=dev)
x = torch.randn(B, 32, 32, 128, device=dev).bfloat16()
g = torch.ones(128, device=dev); b = torch.zeros(128, device=dev)
h = ops.conv_gemm([(x, 9)], w, rowbias=rb, want_stats=True)      # carries gn_stats
out_gn = torch.empty_like(x)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def gemm(n):
    for _ in range(n):
        ops.conv_gemm([(a, 9)], w, rowbias=rb, want_stats=True)


def gn(n):
    for _ in range(n):
        ops.groupnorm_swish(h, g, b, out=out_gn)


def timed(fa, fb, reps=20):
    def body():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        if fa:
            with torch.cuda.stream(s1):
                fa(reps)
        if fb:
            with torch.cuda.stream(s2):
                fb(reps)
        cur.wait_stream(s1); cur.wait_stream(s2)
    body(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    gr.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gr.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


xf = torch.randn(B * 32 * 32 * 64, device=dev)
xb = torch.empty(B * 32 * 32 * 64, device=dev, dtype=torch.bfloat16)


def cast(n):
    for _ in range(n):
        ops.cast_bf16(xf, out=xb)


def copy(n):
    for _ in range(n):
        xb2.copy_(xb1)


xb1 = torch.empty(B * 32 * 32 * 128, device=dev, dtype=torch.bfloat16); xb2 = torch.empty_like(xb1)
tc, tcc, tgc = timed(None, cast), timed(None, copy), timed(gemm, cast)
tgcp = timed(gemm, copy)
tgg = timed(gn, gn)
print(f"cast alone {tc:.1f} us, GEMM || cast {tgc:.1f} us; torch copy alone {tcc:.1f} us, GEMM || copy {tgcp:.1f} us; GN || GN {tgg:.1f} us per pair")
ta, tb, tab = timed(gemm, None), timed(None, gn), timed(gemm, gn)
print(f"GEMM alone {ta:.1f} us, GroupNorm alone {tb:.1f} us, both streams {tab:.1f} us per pair (sum {ta + tb:.1f}, max {max(ta, tb):.1f})")
