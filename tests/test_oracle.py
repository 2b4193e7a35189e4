"""Pins the oracle itself: the reference's own known answers for the sub-steps of the path
 and the documented quirks."""
import math

import numpy as np
import torch

import cases
from oracle import ref_torch as R


def test_rkf45_logistic_known_answer():
    # reference tests/test_solvers.py:11-43 (analytic solution src/ode/logistic.py:43-70):
    # RKF45, h=0.1, 100 steps, default allclose tolerances (rtol 1e-5, atol 1e-8)
    p = {"growth_rate": torch.tensor(1.0), "capacity": torch.tensor(1.0)}
    x0 = torch.tensor([[0.01]])
    xs, _ = R.run_rk(R.ode_logistic, p, "RKF45", 0.1, 0.0, x0, 100)
    t = torch.arange(101) * 0.1
    exact = 1.0 / (1.0 + (1.0 / 0.01 - 1.0) * torch.exp(-t))
    np.testing.assert_allclose(xs[:, 0, 0].numpy(), exact.numpy(), rtol=1e-5, atol=1e-8)


def _rlc_exact(t, R_, L_, C_, x0, v0):
    # src/ode/rlc_circuit.py:63-110: under-, critically- and over-damped closed forms
    alpha = R_ / (2 * L_)
    w0 = 1 / math.sqrt(L_ * C_)
    if alpha < w0:
        wd = math.sqrt(w0 ** 2 - alpha ** 2)
        return np.exp(-alpha * t) * (x0 * np.cos(wd * t) + (v0 + alpha * x0) / wd * np.sin(wd * t))
    if alpha == w0:
        return np.exp(-alpha * t) * (x0 + (v0 + alpha * x0) * t)
    s1 = -alpha + math.sqrt(alpha ** 2 - w0 ** 2)
    s2 = -alpha - math.sqrt(alpha ** 2 - w0 ** 2)
    a2 = (v0 - s1 * x0) / (s2 - s1)
    a1 = x0 - a2
    return a1 * np.exp(s1 * t) + a2 * np.exp(s2 * t)


def test_rkf45_rlc_known_answers():
    # reference tests/test_solvers.py:46-148: h=0.01, 100 steps, rtol=1e-4, atol=1e-7
    for R_, L_, C_ in [(1.0, 1.0, 0.25), (2.0, 1.0, 1.0), (3.0, 1.0, 1.0)]:
        p = {"resistance": torch.tensor(R_), "inductance": torch.tensor(L_),
             "capacitance": torch.tensor(C_)}
        x0 = torch.tensor([[1.0], [0.0]])
        xs, _ = R.run_rk(R.ode_rlc, p, "RKF45", 0.01, 0.0, x0, 100)
        t = np.arange(101) * 0.01
        np.testing.assert_allclose(xs[:, 0, 0].numpy(), _rlc_exact(t, R_, L_, C_, 1.0, 0.0),
                                   rtol=1e-4, atol=1e-7)


def test_sqrt_L_sum_qr_matches_cholesky_through_products():
    # reference tests/test_utils.py:12-41 (fixture seed 7, compare c @ c.T)
    g = torch.Generator().manual_seed(7)
    a = torch.randn(10, 10, generator=g)
    b_full = torch.randn(10, 10, generator=g)
    b = b_full @ b_full.T + 10 * torch.eye(10)
    c = R.sqrt_L_sum_qr(a, torch.linalg.cholesky(b))
    np.testing.assert_allclose((c @ c.T).numpy(), (a @ a.T + b).numpy(), rtol=1e-10)
    z = R.sqrt_L_sum_qr(a, torch.zeros(10, 10))
    np.testing.assert_allclose((z @ z.T).numpy(), (a @ a.T).numpy(), rtol=1e-10)
    c3 = R.sqrt_L_sum_qr_3(a, torch.linalg.cholesky(b), 2 * torch.eye(10))
    np.testing.assert_allclose((c3 @ c3.T).numpy(), (a @ a.T + b + 4 * torch.eye(10)).numpy(), rtol=1e-10)


def test_householder_sign_convention_quirk_q1():
    # qr_sum(1e-12 I, 0.1 I) = -0.1 I under LAPACK's dlarfg convention
    c = R.sqrt_L_sum_qr(1e-12 * torch.eye(3), 0.1 * torch.eye(3))
    np.testing.assert_allclose(c.numpy(), -0.1 * np.eye(3), atol=1e-15)


def test_tableau_consistency():
    # rows of A sum to c; b rows sum to 1 except HeunEuler's propagating row (0.5)
    for name, (A, b, c) in R.TABLEAUX.items():
        np.testing.assert_allclose(A.sum(1).numpy(), c.numpy(), atol=1e-15)
        np.testing.assert_allclose(float(b[0].sum()), 1.0, atol=1e-15)
        np.testing.assert_allclose(float(b[1].sum()), 0.5 if name == "HeunEuler" else 1.0, atol=1e-15)
    nnz = lambda m: int((m != 0).sum())
    A, b, _ = R.TABLEAUX["RKF45"]
    assert (nnz(A), nnz(b[1]), nnz(b[0] - b[1])) == (15, 4, 5)   # SURVEY 8(d) flop model


def test_guard_quirk_only_on_flagged_cases():
    """The sign-sensitive zero-gain guard never fires on the cases run with the
    reference's verbatim guard, and provably differs on the two cases run with intended
    semantics."""
    for name, spec in cases.CASES.items():
        g = cases.load_golden(name)
        if spec.get("guard") == "intended":
            assert int(g["guard_mismatch_steps"]) > 0, name
        else:
            assert int(g["guard_mismatch_steps"]) == 0 and int(g["guard_fired_steps"]) == 0, name


def test_golden_is_reproducible_from_oracle():
    for name in ("lv_heun_none", "vdp_dopri65_predict"):
        live = cases.run_oracle(cases.CASES[name])
        gold = cases.load_golden(name)
        for k in ("x", "P", "eps", "nll"):
            np.testing.assert_allclose(live[k], gold[k], rtol=1e-13, atol=0)


def test_noise_schedules_values():
    from ode_uncertainty_b200.noise_schedules import (CosineAnnealingSchedule,
                                                      ExponentialDecaySchedule, LinearDecaySchedule)
    lin = LinearDecaySchedule(-2.0, 3.0)     # gamma in {1e-2, 1e-5, 1e-8}
    assert [lin.step(i) for i in range(3)] == [10 ** -2.0, 10 ** -5.0, 10 ** -8.0]
    ex = ExponentialDecaySchedule(0.0, 8.0)
    np.testing.assert_allclose([ex.step(i) for i in range(3)],
                               [1.0, 10 ** (-8 * math.log10(2)), 10 ** (-8 * math.log10(3))])
    co = CosineAnnealingSchedule(0.0, -10.0, 4)
    np.testing.assert_allclose([co.step(i) for i in (0, 3, 4)], [1.0, 1e-10, 1.0])
