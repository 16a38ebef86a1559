"""Per-launch timing of one score-net forward (GPU box): record every ops.* call of a forward with its
arguments, then replay each call back to back (no host sync inside the timed region) and print µs, TFLOP/s and
GB/s per call.  Plain text diagnostic, not a bench result.

    python tools/layer_bench.py [B] [reps]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops
from super_diffusion_b200.configs import vpsde
from super_diffusion_b200.models import utils as mutils

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10

cfg = vpsde.get_config()
model, params = mutils.init_model(0, cfg, zero_init_scale=1.0)
net = model.bind(params, dev)
x = torch.randn(B, 32, 32, 3, device=dev)
out = torch.empty(B, 32, 32, 3, device=dev)
net(0.5, x, out=out)
torch.cuda.synchronize()

NAMES = ["conv_gemm", "conv_gemm_s2", "upconv_gemm", "batched_gemm", "attention_probs", "attention_core", "groupnorm_swish", "attention_small",
         "softmax_rows", "upsample2x", "im2col_s2", "im2col_in", "conv_in", "time_embedding", "cast_bf16"]
calls = []
orig = {}


def wrap(name):
    f = getattr(ops, name)
    orig[name] = f

    def g(*a, **k):
        r = f(*a, **k)
        calls.append((name, a, dict(k), r))
        return r
    setattr(ops, name, g)


for n in NAMES:
    if hasattr(ops, n):
        wrap(n)
net(0.5, x, out=out)
torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)


def nbytes(t):
    return t.numel() * t.element_size()


def describe(name, a, k, r):
    """-> (key, flops, bytes)"""
    if name == "conv_gemm":
        srcs, w = a[0], a[1]
        t0 = srcs[0][0]
        K = sum(tp * s.shape[3] for s, tp in srcs)
        N = w.shape[0] if k.get("n_out") is None else k["n_out"]
        Mrows = t0.shape[0] * t0.shape[1] * t0.shape[2]
        segs = "+".join(f"{tp}x{s.shape[3]}" for s, tp in srcs)
        fusedp = getattr(r, "gn_fused", False) or getattr(r, "gn_norm", None) is not None
        keep = len(k["gn"]) > 3 and k["gn"][3] if k.get("gn") is not None else False
        tag = ((" +GN fused" + (" +raw" if keep else "")) if fusedp else " (GN unfused)") if k.get("gn") is not None else ""
        return (f"conv_gemm H{t0.shape[1]} [{segs}] N{N}" + (" stats" if k.get("want_stats") else "") + tag, 2.0 * Mrows * N * K,
                sum(nbytes(s) for s, _ in srcs) + nbytes(w) + nbytes(r))
    if name == "conv_gemm_s2":
        xx, w = a[0], a[1]
        Mrows = xx.shape[0] * xx.shape[1] * xx.shape[2] // 4
        return (f"conv_s2 H{xx.shape[1]} C{xx.shape[3]} N{w.shape[0]}", 2.0 * Mrows * w.shape[0] * w.shape[1], nbytes(xx) + nbytes(w) + nbytes(r))
    if name == "upconv_gemm":
        xx, w4 = a[0], a[1]
        Mrows = xx.shape[0] * xx.shape[1] * xx.shape[2]
        return (f"upconv H{xx.shape[1]} C{xx.shape[3]} N{w4.shape[1]}", 2.0 * Mrows * w4.shape[1] * w4.shape[2] * 4, nbytes(xx) + nbytes(w4) + nbytes(r))
    if name == "batched_gemm":
        A, Bt = a[0], a[1]
        a3 = A if A.dim() == 3 else A.unsqueeze(0)
        b3 = Bt if Bt.dim() == 3 else Bt.unsqueeze(0)
        batch = max(a3.shape[0], b3.shape[0])
        K = k.get("K") or a3.shape[2]
        return (f"batched_gemm b{batch} M{a3.shape[1]} N{b3.shape[1]} K{K}", 2.0 * batch * a3.shape[1] * b3.shape[1] * K,
                batch * (a3.shape[1] + b3.shape[1]) * K * 2 + nbytes(r))
    if name == "attention_probs":
        q = a[0]
        C = k.get("C") or q.shape[2]
        return (f"attn_probs b{q.shape[0]} S{q.shape[1]} C{C}", 2.0 * q.shape[0] * q.shape[1] * q.shape[1] * C,
                2 * q.shape[0] * q.shape[1] * C * 2 + nbytes(r))
    if name == "attention_core":
        q, vt = a[0], a[2]
        C = k.get("C") or q.shape[2]
        return (f"attn_core b{q.shape[0]} S{q.shape[1]} C{C} block{k.get('block')}", 4.0 * q.shape[0] * q.shape[1] * q.shape[1] * C,
                4 * q.shape[0] * q.shape[1] * C * 2 + nbytes(r))
    if name == "groupnorm_swish":
        x0 = a[0]
        x1 = k.get("x1")
        C = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
        has = hasattr(x0, "gn_stats") and (x1 is None or hasattr(x1, "gn_stats"))
        rd = nbytes(x0) + (nbytes(x1) if x1 is not None else 0)
        return (f"groupnorm H{x0.shape[1]} C{C}" + (" (stats given)" if has else ""), 0.0, (1 if has else 2) * rd + nbytes(r))
    return (name, 0.0, 0.0)


def timed(fn):
    """REPS back-to-back launches captured in one CUDA graph (no host launch latency in the timed region)."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / REPS * 1e3   # us


agg = collections.OrderedDict()
total = 0.0
for name, a, k, r in calls:
    key, fl, by = describe(name, a, k, r)
    kk = dict(k)
    if "out" in kk or name in ("groupnorm_swish",):
        pass
    us = timed(lambda: orig[name](*a, **kk))
    total += us
    ent = agg.setdefault(key, [0, 0.0, fl, by])
    ent[0] += 1
    ent[1] += us

print(f"B={B}: {len(calls)} op calls, back-to-back replay sum {total / 1e3:.3f} ms")
print(f"{'op':58s} {'n':>3s} {'us/call':>9s} {'TFLOP/s':>8s} {'GB/s':>7s} {'share':>6s}")
for key, (n, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per = us / n
    print(f"{key:58s} {n:3d} {per:9.1f} {fl / per / 1e6 if fl else 0:8.0f} {by / per / 1e3 if by else 0:7.0f} {us / total * 100:5.1f}%")
kinds = collections.defaultdict(float)
for key, (n, us, fl, by) in agg.items():
    kinds[key.split()[0]] += us
print("by kind:", ", ".join(f"{k} {v / 1e3:.3f} ms" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])))
ms = timed(lambda: net(0.5, x, out=out)) / 1e3
print(f"forward (eager, back to back): {ms:.3f} ms -> {B * 12.154e9 / ms / 1e9:.0f} TFLOP/s")
