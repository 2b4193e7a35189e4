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


def _lv_problem(T=400, h=0.01):
    truth = np.array([1.5, 1.0, 3.0, 1.0])
    xs, _ = RC.rk_run("LotkaVolterra", "RKF45", h, [1.0, 1.0], T, theta=truth)
    rng = np.random.default_rng(3)
    ys = xs[1:] + rng.normal(0.0, 0.05, (T, 2))
    kw = dict(x0="[[1.0, 1.0]]", ts_y=h * np.arange(1, T + 1), ys_x=ys, measurement_matrix=np.eye(2),
              params_range={"alpha": (0.5, 3.0), "beta": (0.3, 2.0), "gamma": (1.0, 5.0), "delta": (0.3, 2.0)},
              gamma_noise_weights=[1.0, 1.0], t0=0.0, tN=T * h, num_tempering_stages=3, obs_noise_var=0.05 ** 2,
              gamma_noise_schedule=LinearDecaySchedule(-2.0, 2.0), seed=5)
    return truth, kw


def test_device_lockstep_lbfgs_reaches_scipys_optimum_without_host_reads():
    """the optimiser loop on the device (csrc/lbfgs.cu).  All restarts advance by one
    `odeu_lbfgs_step` launch per batched objective evaluation; with check_every = 0 nothing is read back
    inside a tempering stage.  SciPy's L-BFGS-B (the reference's optimiser) is the checker: from the same
    starting points both must end in the same optimum of the final (gamma = 0) stage."""
    truth, kw = _lv_problem()
    fb, sb, ob = SQRT_EKF(disable_cov_update=True), S.RKF45(step_size=0.01), O.LotkaVolterra()
    R = 6
    ref = estimation.optimize(fb, sb, ob, lbfgs_maxiter=200, num_random_runs=R, optimizer="scipy", **kw)
    dev = estimation.optimize(fb, sb, ob, lbfgs_maxiter=200, num_random_runs=R, optimizer="device", check_every=0, **kw)
    assert dev["host_reads_inside_stages"] == 0
    assert dev["params_optims"].shape == ref["params_optims"].shape == (R, 3, 4)
    np.testing.assert_array_equal(dev["params_inits"], ref["params_inits"])
    # every restart that SciPy brings to the global basin must end in the same point on the device
    f_ref, f_dev = ref["nll_optims"][:, -1], dev["nll_optims"][:, -1]
    best = f_ref.min()
    good = np.nonzero(f_ref <= best + 1e-6 * abs(best))[0]
    assert good.size >= 2
    for r in good:
        assert abs(f_dev[r] - f_ref[r]) <= 1e-6 * abs(f_ref[r]), (r, f_dev[r], f_ref[r])
        np.testing.assert_allclose(dev["params_optims"][r, -1], ref["params_optims"][r, -1], rtol=2e-4)
    np.testing.assert_allclose(dev["params_optims"][good[0], -1], truth[[0, 1, 3, 2]], rtol=0.05)
    assert set(np.unique(dev["lbfgs_status"])) <= {1, 2, 3, 4}
    # with a flag read every 4 evaluations the stage stops once every restart has converged: evaluations are
    # shared by all restarts (launches ~ evaluations of the slowest restart, not their sum), same optimum
    dev4 = estimation.optimize(fb, sb, ob, lbfgs_maxiter=200, num_random_runs=R, optimizer="device", check_every=4, **kw)
    np.testing.assert_array_equal(dev4["params_optims"], dev["params_optims"])
    assert dev4["kernel_launches"] < dev4["num_nll_evals"].sum()
    assert dev4["kernel_launches"] <= dev4["num_nll_evals"].max(axis=0).sum() + 3 * 4


def test_device_lbfgs_at_c3_restart_count():
    """4,096 restarts (the C3 batch) advance in lock step: one launch pair per evaluation, no host thread
    per restart (the round-1 driver started one Python thread per run)."""
    truth, kw = _lv_problem(T=100)
    kw["num_tempering_stages"] = 2
    fb, sb, ob = SQRT_EKF(disable_cov_update=True), S.RKF45(step_size=0.01), O.LotkaVolterra()
    res = estimation.optimize(fb, sb, ob, lbfgs_maxiter=40, num_random_runs=4096, optimizer="device", check_every=0, **kw)
    assert res["host_reads_inside_stages"] == 0 and res["params_optims"].shape == (4096, 2, 4)
    f = res["nll_optims"][:, -1]
    assert np.isfinite(f).mean() > 0.99
    best = int(np.nanargmin(f))
    # (a horizon of 100 steps identifies the parameters only loosely; the point here is the restart count)
    np.testing.assert_allclose(res["params_optims"][best, -1], truth[[0, 1, 3, 2]], rtol=0.4)
    assert f[best] < np.nanmedian(f)
    assert res["kernel_launches"] <= 2 * (int(1.5 * 40) + 8)
