"""Timing probe for one conv GEMM shape (GPU box): variants of the epilogue work, CUDA-graph replay timing."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops, _lib
import ctypes

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


def raw_conv(srcs, w, N, flags, bias=None, rowbias=None, stats=None, out=None):
    lib = _lib.load()
    x0 = srcs[0][0]
    Bn, H, W = x0.shape[:3]
    arr = (_lib.GemmSrc * len(srcs))()
    for i, (t, taps) in enumerate(srcs):
        arr[i].ptr = t.data_ptr(); arr[i].C = t.shape[3]; arr[i].taps = taps
    P = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)
    rc = lib.sd_conv_gemm(arr, len(srcs), Bn, H, W, P(w), N, P(bias), P(rowbias), rowbias.stride(0) if rowbias is not None else 0,
                          ctypes.c_void_p(0), flags, P(out), out.shape[-1], P(stats), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc


for (H, C, N, extra) in [(32, 128, 128, 0), (32, 256, 128, 0), (16, 256, 256, 0), (16, 512, 256, 0), (16, 256, 512, -1)]:
    taps = 1 if extra == -1 else 9
    a = torch.randn(B, H, H, C, device=dev).bfloat16()
    w = (torch.randn(N, taps * C, device=dev) / math.sqrt(taps * C)).bfloat16()
    bias = torch.randn(N, device=dev); rb = torch.randn(B, N, device=dev)
    out = torch.empty(B, H, H, N, device=dev, dtype=torch.bfloat16)
    stats = torch.empty(B, H * H // 128, 2, N, device=dev)
    fl = 2.0 * B * H * H * N * taps * C
    res = {}
    res["full(bias+rowbias+stats)"] = timed(lambda: raw_conv([(a, taps)], w, N, 0, bias, rb, stats, out))
    res["bias+rowbias"] = timed(lambda: raw_conv([(a, taps)], w, N, 0, bias, rb, None, out))
    res["plain"] = timed(lambda: raw_conv([(a, taps)], w, N, 0, None, None, None, out))
    res["skip-epilogue(0x100)"] = timed(lambda: raw_conv([(a, taps)], w, N, 0x100, None, None, None, out))
    print(f"H{H} {taps}x{C} -> N{N}: " + "  ".join(f"{k} {v:.1f}us/{fl / v / 1e6:.0f}TF" for k, v in res.items()), flush=True)
