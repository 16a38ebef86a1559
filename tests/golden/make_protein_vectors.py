#!/usr/bin/env python
"""Golden vectors for the protein two-component (translations / rotations) SuperDiff mixing, produced by the reference's OWN
method bodies: ``CompositionDiffusion.kappa_AND`` / ``kappa_OR`` / ``compute_kappas`` / ``compute_stoch_dll`` are cut out of
/root/reference/applications/proteins/superdiff/composition.py by line-exact AST extraction and executed unmodified on a stub
``self`` (the module itself cannot be imported: pytorch3d, proteus, se3diff, openfold, wandb are absent).

What is stubbed (third-party, outside the path): the FrameDiff R3 / SO(3) diffusers -- ``b_t``, ``diffusion_coef``,
``drift_coef`` with FrameDiff's published VP forms (b_t = min_b + t (max_b - min_b), g = sqrt(b_t), f = -b_t x / 2) and a
logarithmic SO(3) sigma schedule -- and ``wandb.log``.  The per-step driver below restates the six inline lines of
``latent_mixing`` that turn kappa into dx (composition.py:508-514) and the ll updates (:521-524); it is part of this script,
not of the oracle.

    python tests/golden/make_protein_vectors.py        # writes tests/golden/ref_protein.npz (needs /root/reference)
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/applications/proteins/superdiff/composition.py"
METHODS = ("compute_stoch_dll", "kappa_AND", "kappa_OR", "compute_kappas")


def load_methods():
    src = open(REF).read()
    tree = ast.parse(src)
    lines = src.splitlines()
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CompositionDiffusion")
    out = {}
    wandb = types.SimpleNamespace(log=lambda *a, **k: None)
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in METHODS:
            body = textwrap.dedent("\n".join(lines[fn.lineno - 1:fn.end_lineno]))
            ns = {"torch": torch, "np": np, "wandb": wandb}
            exec(compile(body, f"{REF}:{fn.lineno}", "exec"), ns)
            out[fn.name] = ns[fn.name]
    assert set(out) == set(METHODS)
    return out


class R3:
    min_b, max_b = 0.1, 20.0
    def b_t(self, t):
        return self.min_b + t * (self.max_b - self.min_b)
    def diffusion_coef(self, t):
        return np.sqrt(self.b_t(t))
    def drift_coef(self, x, t):
        return -0.5 * self.b_t(t) * x


class SO3:
    min_sigma, max_sigma = 0.1, 1.5
    def sigma(self, t):
        return np.log(t * np.exp(self.max_sigma) + (1 - t) * np.exp(self.min_sigma))
    def diffusion_coef(self, t):
        g2 = 2 * (np.exp(self.max_sigma) - np.exp(self.min_sigma)) * self.sigma(t) / np.exp(self.sigma(t))
        return np.sqrt(g2)


def make_self(methods, operator, L, n_steps, logp=0.0, T=1.0):
    s = types.SimpleNamespace()
    s.device = "cpu"
    s.r3_diffuser, s.so3_diffuser = R3(), SO3()
    s.num_inference_steps = n_steps
    s.dt = 1 / n_steps
    s.logp_trans = s.logp_rots = logp
    s.T_trans = s.T_rots = T
    s.kappa_operator = operator
    conf = types.SimpleNamespace
    s.comp_diff_conf = conf(diffuser=conf(r3=conf(max_b=R3.max_b, min_b=R3.min_b), so3=conf(max_sigma=SO3.max_sigma, min_sigma=SO3.min_sigma)))
    s.sigma_r = lambda t: torch.tensor(s.so3_diffuser.sigma(float(t)))
    z = lambda: torch.zeros((1, n_steps + 1))          # composition.py:178-181: float32 accumulators
    s.ll_proteus_trans, s.ll_framediff_trans, s.ll_proteus_rots, s.ll_framediff_rots = z(), z(), z(), z()
    s.val_dict = {m: {c: {} for c in ("trans", "rots")} for m in ("proteus", "framediff")}
    for name, fn in methods.items():
        setattr(s, name, types.MethodType(fn, s))
    return s


def run_case(methods, operator, L=48, n_steps=6, seed=0, logp=0.0, T=1.0):
    g = torch.Generator().manual_seed(seed)
    s = make_self(methods, operator, L, n_steps, logp, T)
    x = torch.randn(1, L, 3, generator=g, dtype=torch.float64)
    W = {k: 0.3 * torch.randn(3, 3, generator=g, dtype=torch.float64) for k in ("pt", "ft", "pr", "fr")}
    rec = {k: [] for k in ("x", "eps", "s_pt", "s_ft", "s_pr", "s_fr", "kappa_trans", "kappa_rots", "dx_trans", "dx_rots",
                           "beta_trans", "beta_rots", "a_trans", "t", "ll", "lift_trans", "lift_rots")}
    dt = 1 / n_steps
    ts = np.linspace(0.01, 1.0, n_steps)[::-1]
    for i, t in enumerate(ts):
        t_ = torch.tensor(t)
        # stand-in score models: smooth functions of the state (the real ones are Proteus / FrameDiff networks)
        sc = {k: torch.tanh(x @ W[k]) * (1.0 + 0.1 * t) + (0.2 if k[0] == "p" else -0.1) for k in W}
        eps = torch.randn(x.shape, generator=g, dtype=torch.float64)
        beta_t_trans = torch.tensor(0.5 * s.r3_diffuser.diffusion_coef(t) ** 2)              # composition.py:485-486
        f_x_trans = s.r3_diffuser.drift_coef(x, t)                                           # :487
        beta_t_rots = torch.tensor(0.5 * s.so3_diffuser.diffusion_coef(t) ** 2)              # :489-490
        s.val_dict["proteus"]["rots"]["score"], s.val_dict["proteus"]["trans"]["score"] = sc["pr"], sc["pt"]     # :492-495
        s.val_dict["framediff"]["rots"]["score"], s.val_dict["framediff"]["trans"]["score"] = sc["fr"], sc["ft"]
        kt, kr = s.compute_kappas(i, t_, x=x, beta_t_rots=beta_t_rots, beta_t_trans=beta_t_trans, eps=eps, f_x_trans=f_x_trans)   # :509
        dx_trans = -dt * (s.r3_diffuser.drift_coef(x, t) - 2 * beta_t_trans * (sc["ft"] + kt * (sc["pt"] - sc["ft"])))          # :514-515
        dx_trans = dx_trans + torch.sqrt(2 * beta_t_trans * dt) * eps                                                            # :516
        dx_rots = dt * 2 * beta_t_rots * (sc["fr"] + kr * (sc["pr"] - sc["fr"]))                                                 # :518
        dx_rots = dx_rots + torch.sqrt(2 * beta_t_rots * dt) * eps                                                               # :519
        s.compute_stoch_dll(t, x, dx_trans, "trans")                                                                             # :522
        s.compute_stoch_dll(t, None, dx_rots, "rots")                                                                            # :523
        s.ll_proteus_trans[:, i + 1] = s.ll_proteus_trans[:, i] + s.val_dict["proteus"]["trans"]["dlldt"]                        # :526-529
        s.ll_framediff_trans[:, i + 1] = s.ll_framediff_trans[:, i] + s.val_dict["framediff"]["trans"]["dlldt"]
        s.ll_proteus_rots[:, i + 1] = s.ll_proteus_rots[:, i] + s.val_dict["proteus"]["rots"]["dlldt"]
        s.ll_framediff_rots[:, i + 1] = s.ll_framediff_rots[:, i] + s.val_dict["framediff"]["rots"]["dlldt"]
        # the scalar kappa_AND adds as lift / kappa_div (composition.py:382-403,417): logp * normalised(-dim/2 log sigma_t) / steps
        dim = L * 3
        def lift_of(sig, mx, mn):
            st, lo, hi = -0.5 * dim * np.log(sig), -0.5 * dim * np.log(mx), -0.5 * dim * np.log(mn)
            return logp * (st - lo) / (hi - lo) / n_steps
        lift_t = lift_of(np.sqrt(s.r3_diffuser.b_t(t)), np.sqrt(R3.max_b), np.sqrt(R3.min_b))
        lift_r = lift_of(s.so3_diffuser.sigma(t), SO3.max_sigma, SO3.min_sigma)
        rec["lift_trans"].append(np.asarray(lift_t)); rec["lift_rots"].append(np.asarray(lift_r))
        for k, v in (("x", x), ("eps", eps), ("s_pt", sc["pt"]), ("s_ft", sc["ft"]), ("s_pr", sc["pr"]), ("s_fr", sc["fr"]),
                     ("kappa_trans", torch.as_tensor(kt)), ("kappa_rots", torch.as_tensor(kr)), ("dx_trans", dx_trans), ("dx_rots", dx_rots),
                     ("beta_trans", beta_t_trans), ("beta_rots", beta_t_rots), ("a_trans", torch.tensor(-0.5 * s.r3_diffuser.b_t(t))),
                     ("t", t_)):
            rec[k].append(np.asarray(v.detach().double().numpy()))
        rec["ll"].append(np.array([float(s.ll_proteus_trans[0, i + 1]), float(s.ll_framediff_trans[0, i + 1]),
                                   float(s.ll_proteus_rots[0, i + 1]), float(s.ll_framediff_rots[0, i + 1])]))
        x = x + dx_trans          # fixture driver only: the reference applies dx on SE(3) (se3_diffuser.reverse, :545-556)
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(L=L, n_steps=n_steps, dt=dt, logp=logp, T=T)
    return out


def main():
    methods = load_methods()
    cases = {"and": run_case(methods, "AND", seed=1), "or": run_case(methods, "OR", seed=2, logp=0.3, T=2.0),
             "and_lift": run_case(methods, "AND", seed=3, logp=0.7)}
    flat = {f"{c}/{k}": v for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(HERE, "ref_protein.npz"), **flat)
    print("wrote ref_protein.npz:", {c: {k: np.shape(v) for k, v in d.items() if k in ("x", "kappa_trans", "ll")} for c, d in cases.items()})
    print("kappa_trans:", {c: d["kappa_trans"].ravel()[:3] for c, d in cases.items()})


if __name__ == "__main__":
    main()
