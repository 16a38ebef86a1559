"""Generate the committed golden vectors from the CPU oracle (fp64 math on fp32-rounded inputs).

The reference ships no golden vectors for this path and cannot be imported here
(SURVEY.md §4, F8), so these fixtures pin the *oracle* (and through it the kernels)
against silent drift; they are not outputs of the reference itself ("parity unpinned").

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import schedule as S   # noqa: E402
from oracle import steps as O      # noqa: E402
from oracle import toy             # noqa: E402


def f32(x):
    return float(torch.tensor(x, dtype=torch.float32))


def step_cases():
    out = {}
    cases = [("cifar_or", 2, 3072, 2, O.MODE_OR, O.DLOGQ_CIFAR_MAXSUB, 1e6, 0.62, 5e-3),
             ("cifar_or_m4", 1, 3072, 4, O.MODE_OR, O.DLOGQ_CIFAR_MAXSUB, 1e6, 0.2, 5e-3),
             ("cifar_and", 2, 3072, 2, O.MODE_AND, O.DLOGQ_ITO, 1.0, 0.9, 1e-3),
             ("and_m3", 4, 768, 3, O.MODE_AND, O.DLOGQ_ITO, 1.0, 0.5, 1e-3),
             ("avg", 2, 768, 2, O.MODE_AVG, O.DLOGQ_NONE, 1.0, 0.4, 5e-3),
             ("toy_or", 16, 2, 2, O.MODE_OR, O.DLOGQ_ITO, 1.0, 0.3, 1e-3),
             ("toy_and", 16, 2, 2, O.MODE_AND, O.DLOGQ_ITO, 1.0, 0.3, 1e-3)]
    for name, B, D, M, mode, dmode, temp, t, dt in cases:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        x, eps = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
        s = torch.randn(M, B, D, generator=g)
        logq = (1e-6 if temp > 1 else 1.0) * torch.randn(B, M, generator=g)
        a, b, sg, dtf = f32(S.dlog_alphadt(t)), f32(S.beta(t)), f32(S.sigma(t)), f32(dt)
        ito = float(D * D) if dmode == O.DLOGQ_ITO else 0.0
        xo, lq, w = O.step_vpsde_gram(x, eps, s, logq, a, b, sg, dtf, mode, dmode, temperature=temp,
                                      ito_const=ito * dtf * a)
        out[name] = dict(x=x.numpy(), eps=eps.numpy(), s=s.numpy(), logq=logq.numpy(), t=t, dt=dt, mode=mode,
                         dmode=dmode, temperature=temp, ito_scale=ito, x_out=xo.numpy(), logq_out=lq.numpy(),
                         weights=w.numpy())
    return out


def sd_cases():
    out = {}
    for name, mode in (("sd_and", O.MODE_AND), ("sd_or", O.MODE_OR), ("sd_avg", O.MODE_AVG)):
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        B, D = 3, 4 * 16 * 16
        lat = 14.6 * torch.randn(B, D, generator=g)
        z, vo, vb, vu = (torch.randn(B, D, generator=g) for _ in range(4))
        ll = 1.0 + 0.1 * torch.randn(B, 2, generator=g)
        sigma, dsigma = f32(3.2), f32(-0.41)
        kw = dict(guidance=7.5, lift_term=0.02, temperature=2.0, logp=0.1, kappa_fixed=0.5)
        lo, l2, k = O.step_edm_gram(lat, z, vo, vb, vu, ll, sigma, dsigma, mode, **kw)
        out[name] = dict(lat=lat.numpy(), z=z.numpy(), vo=vo.numpy(), vb=vb.numpy(), vu=vu.numpy(), ll=ll.numpy(),
                         sigma=sigma, dsigma=dsigma, mode=mode, lat_out=lo.numpy(), ll_out=l2.numpy(), kappa=k.numpy(), **kw)
    return out


def toy_loops():
    out = {}
    B, n, dt = 64, 50, 2e-2
    x0 = torch.randn(B, 2, generator=torch.Generator().manual_seed(7))
    noise = torch.randn(n, B, 2, generator=torch.Generator().manual_seed(8))
    fns = [toy.mixture_sscore("up"), toy.mixture_sscore("down")]
    for mode in ("or", "and"):
        x, ll, tr = toy.loop_toy(fns, x0.double(), noise, mode, n, dt, record=True)
        out[f"toy_loop_{mode}"] = dict(x0=x0.numpy(), noise=noise.numpy(), n=n, dt=dt, x=x.numpy(), ll=ll.numpy(),
                                       kappa=tr["kappa"].numpy())
    return out


def main():
    for fname, cases in (("steps.npz", step_cases()), ("sd_steps.npz", sd_cases()), ("toy_loops.npz", toy_loops())):
        flat = {}
        for cname, d in cases.items():
            for k, v in d.items():
                flat[f"{cname}/{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, fname), **flat)
        print(fname, sum(v.nbytes for v in flat.values()) // 1024, "KiB uncompressed")


if __name__ == "__main__":
    main()
