"""GPU: tempered parameter estimation (ode_uncertainty_b200/estimation.py, the caller of the hot path:
scripts/run_parameter_estimation.py:49-308, 540-682) on a Lotka-Volterra problem shaped like
configs/params/lotkavolterra4.yaml: every random run advances in lock step, one forward-mode
gradient launch per optimiser round."""
import numpy as np
import pytest

from oracle import ref_cpp as RC
from ode_uncertainty_b200 import estimation, ode as O, solvers as S
from ode_uncertainty_b200.filters import SQRT_EKF
from ode_uncertainty_b200.noise_schedules import LinearDecaySchedule

pytestmark = pytest.mark.gpu


def test_tempered_lbfgsb_recovers_lotka_volterra_parameters():
    truth = np.array([1.5, 1.0, 3.0, 1.0])
    h, T = 0.01, 400
    xs, _ = RC.rk_run("LotkaVolterra", "RKF45", h, [1.0, 1.0], T, theta=truth)
    rng = np.random.default_rng(3)
    ys = xs[1:] + rng.normal(0.0, 0.05, (T, 2))
    ts = h * np.arange(1, T + 1)
    rngs = {"alpha": (0.5, 3.0), "beta": (0.3, 2.0), "gamma": (1.0, 5.0), "delta": (0.3, 2.0)}
    res = estimation.optimize(SQRT_EKF(disable_cov_update=True), S.RKF45(step_size=h), O.LotkaVolterra(),
                              x0="[[1.0, 1.0]]", ts_y=ts, ys_x=ys, measurement_matrix=np.eye(2), params_range=rngs,
                              gamma_noise_weights=[1.0, 1.0], t0=0.0, tN=T * h, num_tempering_stages=3,
                              obs_noise_var=0.05 ** 2, gamma_noise_schedule=LinearDecaySchedule(-2.0, 2.0),
                              lbfgs_maxiter=60, num_random_runs=4, seed=5)
    assert res["params_optims"].shape == (4, 3, 4) and res["nll_optims"].shape == (4, 3)
    assert list(res["params_name"]) == ["alpha", "beta", "delta", "gamma"]       # sorted-key order
    assert res["gammas"][-1] == 0.0 and res["gammas"][0] == pytest.approx(1e-2)
    best = int(np.argmin(res["nll_optims"][:, -1]))
    est = res["params_optims"][best, -1]
    np.testing.assert_allclose(est, truth[[0, 1, 3, 2]], rtol=0.05)
    # lock step: far fewer launches than optimiser evaluations
    assert res["kernel_launches"] < res["num_nll_evals"].sum()
    assert np.all(res["num_lbfgs_iters"] <= 60)
