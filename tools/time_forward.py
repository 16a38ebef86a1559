"""Timing probe (GPU box): score-net forward at B, per-kernel-kind breakdown via CUDA events,
and a launch-shape sweep of the fused step kernel.  Prints plain text; not a bench result."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops, sde
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512


def ev_time(fn, iters=5, warm=2):
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
net = model.bind(params, dev)
x = torch.randn(B, 32, 32, 3, device=dev)
out = torch.empty(B, 32, 32, 3, device=dev)
ms = ev_time(lambda: net(0.5, x, out=out))
flops = B * 12.154e9
print(f"forward B={B}: {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s (12.154 GFLOP/sample)")

# per-op breakdown: wrap ops.* with event timing (serialising; shares matter, not absolutes)
acc = collections.defaultdict(float)
cnt = collections.Counter()
orig = {}
def wrap(name):
    f = getattr(ops, name)
    orig[name] = f
    def g(*a, **k):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = f(*a, **k); e.record(); torch.cuda.synchronize()
        key = name
        if name == "conv_gemm":
            srcs = a[0]; t0 = srcs[0][0]
            key = f"conv_gemm H{t0.shape[1]} K{sum(tp * s_.shape[3] for s_, tp in srcs)} N{a[1].shape[0]}"
        if name == "groupnorm_swish":
            key = f"groupnorm H{a[0].shape[1]} C{a[0].shape[3] + (k['x1'].shape[3] if k.get('x1') is not None else 0)}"
        acc[key] += s.elapsed_time(e); cnt[key] += 1
        return r
    setattr(ops, name, g)
for n in ["conv_gemm", "batched_gemm", "groupnorm_swish", "attention_small", "softmax_rows", "upsample2x",
          "im2col_s2", "conv_in", "time_embedding"]:
    wrap(n)
net(0.5, x, out=out)
acc.clear(); cnt.clear()
net(0.5, x, out=out)
tot = sum(acc.values())
print(f"serialised sum {tot:.3f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"  {k:40s} n={cnt[k]:3d} {v:8.3f} ms {100 * v / tot:5.1f}%")
for n, f in orig.items():
    setattr(ops, n, f)

# fused step sweep
for (Bs, M, mode) in [(512, 2, ops.MODE_OR), (8192, 2, ops.MODE_AND), (8192, 2, ops.MODE_OR), (2048, 8, ops.MODE_OR)]:
    D = 3072
    xs = torch.randn(Bs, D, device=dev); ns = torch.randn(Bs, D, device=dev)
    sc = [torch.randn(Bs, D, device=dev) for _ in range(M)]
    lq = torch.zeros(Bs, M, device=dev); w = torch.zeros(Bs, M, device=dev); xo = torch.empty_like(xs)
    bytes_ = 4 * Bs * D * (M + 3)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    res = []
    for shape in [None, (256, 3, 1), (192, 4, 1), (192, 2, 1), (128, 2, 1), (128, 3, 1), (256, 1, 1), (128, 3, 2), (128, 2, 4), (192, 1, 4), (96, 1, 8), (256, 1, 4), (64, 3, 4), (128, 1, 8), (256, 2, 2)]:
        def f():
            ops.step_vpsde(xs, ns, sc, lq, -5.0, 5.0, 0.5, 1e-3, mode, ops.DLOGQ_CIFAR_MAXSUB if mode == ops.MODE_OR else ops.DLOGQ_ITO,
                           temperature=1e6, x_out=xo, weights=w, launch_shape=shape)
        try:
            f(); torch.cuda.synchronize()
        except Exception as ex:
            continue
        ts = []
        for _ in range(6):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); f(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        best = sorted(ts)[len(ts) // 2]
        res.append((shape, best))
    print(f"step B={Bs} M={M} mode={mode} bytes={bytes_ / 1e6:.1f}MB: " +
          "  ".join(f"{s}:{t * 1e3:.1f}us/{bytes_ / t / 1e6:.0f}GB/s" for s, t in res))
