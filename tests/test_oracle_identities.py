"""The oracle's Gram-form restatement (what the kernels implement) equals the
literal transcription of the reference, in fp64."""
import math

import pytest
import torch

from oracle import schedule as S
from oracle import steps as O

torch.manual_seed(0)


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float64)


@pytest.mark.parametrize("M", [1, 2, 3, 5])
@pytest.mark.parametrize("t", [1.0, 0.37, 0.005])
def test_cifar_or_product_form_equals_gram_form(M, t):
    B, dt = 6, 5e-3
    x = _rand(B, 4, 4, 3, seed=1)
    eps = _rand(B, 4, 4, 3, seed=2)
    s = _rand(M, B, 4, 4, 3, seed=3)
    logq = 1e-6 * _rand(B, M, seed=4)       # so that softmax(1e6 * logq) is not one-hot
    dx, dlogq, w = O.or_step_cifar_literal(x, logq, s, eps, t, dt)
    x2, lq2, w2 = O.step_vpsde_gram(x, eps, s, logq, S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt,
                                    O.MODE_OR, O.DLOGQ_CIFAR_MAXSUB, temperature=1e6)
    assert torch.allclose(x + dx, x2, rtol=0, atol=1e-12)
    assert torch.allclose(logq + dlogq, lq2, rtol=0, atol=1e-10)
    assert torch.allclose(w, w2, atol=1e-12)
    assert torch.allclose(w.sum(1), torch.ones(B, dtype=torch.float64), atol=1e-14)
    assert (dlogq.max(dim=1).values == 0).all()


def test_cifar_avg_equals_gram_form():
    B, M, t, dt = 5, 3, 0.6, 5e-3
    x, eps, s = _rand(B, 8, seed=1), _rand(B, 8, seed=2), _rand(M, B, 8, seed=3)
    dx, dlogq = O.avg_step_cifar_literal(x, s, eps, t, dt, stoch=True)
    x2, lq2, w2 = O.step_vpsde_gram(x, eps, s, torch.zeros(B, M), S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt,
                                    O.MODE_AVG, O.DLOGQ_NONE)
    assert torch.allclose(x + dx, x2, atol=1e-13)
    assert (lq2 == 0).all() and (dlogq == 0).all()


@pytest.mark.parametrize("t", [1.0, 0.2, 0.0010093])
def test_toy_or_equals_gram_form(t):
    B, dt = 16, 1e-3
    x, eps = _rand(B, 2, seed=1), _rand(B, 2, seed=2)
    s1, s2 = _rand(B, 2, seed=3), _rand(B, 2, seed=4)
    ll = _rand(B, 2, seed=5)
    dx, dll, kappa = O.or_step_toy_literal(x, ll, s1, s2, eps, t, dt)
    x2, ll2, w2 = O.step_vpsde_gram(x, eps, torch.stack([s1, s2]), ll, S.dlog_alphadt(t), S.beta(t), S.sigma(t),
                                    dt, O.MODE_OR, O.DLOGQ_ITO, temperature=1.0)
    assert torch.allclose(x + dx, x2, atol=1e-13)
    assert torch.allclose(ll + dll, ll2, atol=1e-10)
    assert torch.allclose(kappa, w2[:, 0], atol=1e-14)
    # M = 2 softmax == the notebook's explicit max-shifted form (superposition_edu.ipynb cell 24)
    mx = torch.maximum(ll[:, 0], ll[:, 1])
    k2 = torch.exp(ll[:, 0] - mx) / (torch.exp(ll[:, 0] - mx) + torch.exp(ll[:, 1] - mx))
    assert torch.allclose(kappa, k2, atol=1e-14)


@pytest.mark.parametrize("D", [2, 48])
def test_toy_and_equals_gram_form_and_equalises(D):
    B, t, dt = 16, 0.43, 1e-3
    x, eps = _rand(B, D, seed=1), _rand(B, D, seed=2)
    s1, s2 = _rand(B, D, seed=3), _rand(B, D, seed=4)
    ll = _rand(B, 2, seed=5)
    dx, dll, kappa = O.and_step_toy_literal(x, s1, s2, eps, t, dt, ndim=D)
    x2, ll2, w2 = O.step_vpsde_gram(x, eps, torch.stack([s1, s2]), ll, S.dlog_alphadt(t), S.beta(t), S.sigma(t),
                                    dt, O.MODE_AND, O.DLOGQ_ITO)
    assert torch.allclose(kappa, w2[:, 0], rtol=1e-10, atol=1e-12)
    assert torch.allclose(x + dx, x2, atol=1e-11)
    assert torch.allclose(ll + dll, ll2, rtol=1e-10, atol=1e-9)
    # AND makes the two density increments equal
    assert torch.allclose(dll[:, 0], dll[:, 1], rtol=0, atol=1e-9)


@pytest.mark.parametrize("M", [3, 4, 8])
def test_general_and_equalises_all_increments(M):
    B, D, t, dt = 7, 96, 0.5, 1e-3
    x, eps, s = _rand(B, D, seed=1), _rand(B, D, seed=2), _rand(M, B, D, seed=3)
    ll = torch.zeros(B, M, dtype=torch.float64)
    x2, ll2, w = O.step_vpsde_gram(x, eps, s, ll, S.dlog_alphadt(t), S.beta(t), S.sigma(t), dt,
                                   O.MODE_AND, O.DLOGQ_ITO)
    assert torch.allclose(w.sum(1), torch.ones(B, dtype=torch.float64), atol=1e-12)
    assert torch.allclose(ll2, ll2[:, :1].expand_as(ll2), atol=1e-8)
    # the increments recomputed literally from dx agree
    dx = (x2 - x)
    for i in range(M):
        d = O.stoch_dll_toy_literal(t, dt, x, dx, s[i], ndim=D)
        assert torch.allclose(d, ll2[:, i], atol=1e-8)


@pytest.mark.parametrize("method,mode", [("and", O.MODE_AND), ("or", O.MODE_OR), ("avg", O.MODE_AVG)])
def test_sd_literal_equals_gram_form(method, mode):
    B = 4
    lat, z = 14.6 * _rand(B, 4, 8, 8, seed=1), _rand(B, 4, 8, 8, seed=2)
    vo, vb, vu = _rand(B, 4, 8, 8, seed=3), _rand(B, 4, 8, 8, seed=4), _rand(B, 4, 8, 8, seed=5)
    ll = 1.0 + 0.1 * _rand(B, 2, seed=6)
    sigma, dsigma = 3.2, -0.41
    dx, ll1, kappa = O.sd_step_literal(lat, z, vo, vb, vu, ll, sigma, dsigma, method, guidance_scale=7.5,
                                       lift=0.3, num_inference_steps=50, T=2.0, logp=0.1)
    lat2, ll2, k2 = O.step_edm_gram(lat, z, vo, vb, vu, ll, sigma, dsigma, mode, guidance=7.5,
                                    lift_term=sigma * 0.3 / 50, temperature=2.0, logp=0.1, kappa_fixed=0.5)
    assert torch.allclose(kappa, k2, rtol=1e-11, atol=1e-12)
    assert torch.allclose(lat + dx, lat2, atol=1e-11)
    assert torch.allclose(ll1, ll2, rtol=1e-11, atol=1e-9)
    if method == "and":   # AND equalises the two increments when lift = 0
        _, ll0, _ = O.sd_step_literal(lat, z, vo, vb, vu, ll, sigma, dsigma, "and", lift=0.0)
        inc = ll0 - ll
        assert torch.allclose(inc[:, 0], inc[:, 1], atol=1e-8)


def test_time_grid_float32_drift():
    # 999 float32 decrements of 1e-3 leave t = 0.0010093, not 0.001
    g32 = S.time_grid(1000, 1e-3, "float32")
    g64 = S.time_grid(1000, 1e-3, "float64")
    assert abs(g32[-1] - 0.0010093) < 2e-7
    assert abs(g64[-1] - 0.001) < 1e-12
    assert len(g32) == 1000 and g32[0] == 1.0


def test_edm_sigma_table_shape_and_monotone():
    sig, ts, init = S.edm_sigmas(50)
    assert sig.shape == (51,) and ts.shape == (50,) and sig[-1] == 0.0
    assert (sig[:-1] > sig[1:]).all()
    assert abs(init - sig[0]) < 1e-7 and 14.0 < init < 15.0   # SD v1 sigma_max ~ 14.6


def test_schedule_values():
    # beta(t) = sigma * d/dt log(sigma/alpha) with sigma = t   (cifar/dynamics.py:25-27)
    for t in (1.0, 0.5, 0.01):
        assert math.isclose(S.beta(t), t * (1.0 / t - S.dlog_alphadt(t)), rel_tol=1e-12)
