"""Plain-RK least-squares baseline loss (scripts/run_parameter_estimation_baseline.py
::nll :552-632) served by the gradient kernels as the degenerate filter P0 = 0, Q = 0,
disable_cov_update (zero gain, S = R).  Checked against the reference's own baseline nll() and its
reverse-mode gradient (tests/golden/ref_baseline_*.npz, oracle/make_golden_ref.py baseline)."""
import os

import numpy as np
import pytest

import cases
import util as U
from ode_uncertainty_b200 import runners
from test_grad import REF_GRAD


def _check(backend, name, batch=1):
    ref = dict(np.load(os.path.join(cases.GOLDEN, f"ref_baseline_{name}.npz")))
    spec = dict(cases.CASES[name], disable=True)
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    ob = REF_GRAD[name]()
    _, _, perm = runners.param_layout(ob)
    theta_sorted = ref["pn_base"] * (ref["hi"] - ref["lo"]) + ref["lo"]
    idx_builder = np.array([int(np.nonzero(perm == j)[0][0]) for j in range(perm.size)])
    n = m["x0"].numel()
    xb = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, 0)
    nll, g = U.run_grad(backend, plan, xb, m["T"], idx_builder, t0=m["t0"], P0_sqrt=np.zeros((n, n)),
                        theta_shared=theta_sorted[perm], gamma_sqrt=0.0, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(),
                        ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
    want = float(ref["nll_base"])
    assert abs(nll[batch - 1] - want) <= 1e-9 * abs(want)
    g_norm = g[batch - 1] * (ref["hi"] - ref["lo"])
    scale = np.max(np.abs(ref["grad_norm_base"]))
    np.testing.assert_allclose(g_norm, ref["grad_norm_base"], rtol=1e-6, atol=1e-6 * scale)


@pytest.mark.parametrize("name", list(REF_GRAD))
def test_baseline_loss_and_gradient_match_reference(name):
    _check("hostemu", name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REF_GRAD))
def test_cuda_baseline_loss_and_gradient(name):
    _check("gpu", name, batch=5)


@pytest.mark.gpu
def test_cuda_optimize_baseline_recovers_lotka_volterra():
    from ode_uncertainty_b200 import Plan, _native as N, estimation, ode as O, solvers as S
    truth = np.array([1.5, 1.0, 3.0, 1.0])
    h, T = 0.01, 400
    plan = Plan(N.ODE_LOTKA_VOLTERRA, N.SOLVER_RKF45, h)
    xs = runners.solve_trajectory(plan, [1.0, 1.0], T, theta_shared=truth)
    ys = xs[1:] + np.random.default_rng(3).normal(0.0, 0.05, (T, 2))
    rngs = {"alpha": (0.5, 3.0), "beta": (0.3, 2.0), "gamma": (1.0, 5.0), "delta": (0.3, 2.0)}
    res = estimation.optimize_baseline(S.RKF45(step_size=h), O.LotkaVolterra(), x0="[[1.0, 1.0]]",
                                       ts_y=h * np.arange(1, T + 1), ys_x=ys, measurement_matrix=np.eye(2),
                                       params_range=rngs, t0=0.0, tN=T * h, obs_noise_var=0.05 ** 2,
                                       lbfgs_maxiter=100, num_random_runs=4, seed=5)
    assert res["params_optims"].shape == (4, 4) and res["nll_optims"].shape == (4,)
    best = int(np.argmin(res["nll_optims"]))
    np.testing.assert_allclose(res["params_optims"][best], truth[[0, 1, 3, 2]], rtol=0.05)
