"""Diagnostic for the tcgen05 GEMM (run on the GPU box): structured inputs that
reveal descriptor / swizzle / K-advance mistakes, then random-input errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from super_diffusion_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(M, N, K, structured):
    if structured:
        A = torch.zeros(M, K)
        A[torch.arange(M), torch.arange(M) % K] = 1.0
        Bt = (torch.arange(K)[None, :] + 100.0 * torch.arange(N)[:, None]).float()
        Bt = Bt % 251   # exactly representable in bf16? keep small integers (< 256)
    else:
        A = torch.randn(M, K)
        Bt = torch.randn(N, K) / K ** 0.5
    A16, B16 = A.bfloat16(), Bt.bfloat16()
    out = ops.batched_gemm(A16.to(dev), B16.to(dev), out_f32=True)
    torch.cuda.synchronize()
    ref = A16.double() @ B16.double().T
    got = out[0].cpu().double()
    err = (got - ref).abs()
    print(f"M={M} N={N} K={K} structured={structured}: max_err={err.max().item():.4g} ref_max={ref.abs().max().item():.4g} "
          f"bad_rows={int((err.max(1).values > 1e-2).sum())} bad_cols={int((err.max(0).values > 1e-2).sum())}")
    if err.max().item() > 1e-2:
        print(" got[0:8,0:4]=\n", got[:8, :4])
        print(" ref[0:8,0:4]=\n", ref[:8, :4])
        print(" got[:,0]=", got[:, 0].tolist()[:M])
        print(" ref[:,0]=", ref[:, 0].tolist()[:M])
    return err.max().item()


for cfg in [(128, 16, 64, True), (128, 16, 64, False), (128, 64, 128, True), (128, 128, 256, False),
            (256, 256, 512, False), (300, 40, 192, False)]:
    try:
        run(*cfg)
    except Exception as e:  # keep going: later cases may still be informative
        print("EXC", cfg, repr(e))
        break
print("diag done")
