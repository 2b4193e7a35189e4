"""The reference's own unit tests for the linear-algebra helpers on the path
(tests/test_utils.py:12-52 of the reference), restated against `ode_uncertainty_b200.utils`:
factors are compared through `c @ c.T`, exactly like the reference does (:31)."""
import numpy as np
import torch

from ode_uncertainty_b200 import utils


def _rand_10x10():
    return torch.as_tensor(np.random.default_rng(7).normal(size=(10, 10)))      # reference: key(7)


def test_sqrt_L_sum_qr():
    a = _rand_10x10()
    b = a @ a.T
    b_sqrt = torch.linalg.cholesky(b)
    c = utils.sqrt_L_sum_qr(a, b_sqrt)
    ref = torch.linalg.cholesky(a @ a.T + b)
    assert torch.allclose(c @ c.T, ref @ ref.T, rtol=1e-5, atol=1e-8)
    assert torch.allclose(torch.triu(c, 1), torch.zeros_like(c)) and bool((torch.diagonal(c) >= 0).all())


def test_sqrt_L_sum_qr_zero_and_three_blocks():
    a = _rand_10x10()
    c = utils.sqrt_L_sum_qr(a, torch.zeros(10, 10, dtype=torch.float64))
    assert torch.allclose(c @ c.T, a @ a.T, rtol=1e-5, atol=1e-8)
    d = torch.diag(torch.arange(1.0, 11.0, dtype=torch.float64))
    c3 = utils.sqrt_L_sum_qr_3(a, d, 0.5 * a.T)
    assert torch.allclose(c3 @ c3.T, a @ a.T + d @ d.T + 0.25 * a.T @ a, rtol=1e-10, atol=1e-10)
    # the Householder sign quirk Q1 of the reference (qr_sum(1e-12 I, 0.1 I) = -0.1 I) does not leak out
    q = utils.sqrt_L_sum_qr(utils.const_diag(3, 1e-12), utils.const_diag(3, 0.1))
    assert torch.allclose(q, utils.const_diag(3, 0.1))


def test_negative_log_gaussian_sqrt_against_density():
    from scipy.stats import multivariate_normal
    a = _rand_10x10()
    P = (a @ a.T).numpy() + np.eye(10)
    L = np.linalg.cholesky(P)
    rng = np.random.default_rng(1)
    x, m = rng.normal(size=10), rng.normal(size=10)
    got = float(utils.negative_log_gaussian_sqrt(x, m, L))
    assert abs(got + multivariate_normal.logpdf(x, mean=m, cov=P)) < 1e-9
    assert abs(float(utils.negative_log_gaussian_sqrt(x, m, -L)) - got) < 1e-12     # sign-safe (|diag|)


def test_normalize_roundtrip_in_sorted_key_order():
    vals = {"g_Na": np.array([25.0, 20.0]), "C": np.array([1.0]), "V_T": np.array([-60.0, -61.0])}
    lo = {k: v - 2.0 for k, v in vals.items()}
    hi = {k: v + 6.0 for k, v in vals.items()}
    nz = utils.normalize(vals, lo, hi)
    assert all(np.allclose(nz[k], 0.25) for k in vals)
    back = utils.inv_normalize(nz, lo, hi)
    assert all(np.allclose(back[k], vals[k]) for k in vals)
    flat = utils.normalize(np.array([1.0, 2.0]), np.array([0.0, 0.0]), np.array([2.0, 4.0]))
    assert np.allclose(flat, 0.5)
