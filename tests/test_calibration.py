"""Calibration sweep (scripts/run_calibration_conrad_baseline_calibration.py:126-222):
the filter NLL - mean over all steps of the nan_to_num'ed per-step terms - for a range of static
process-noise levels, one trajectory per level (per-trajectory `cov_scale`), against Oracle-B run
once per level."""
import numpy as np
import pytest

import util as U
from oracle import ref_cpp as RC
from ode_uncertainty_b200 import _native as N


def _problem(T=120):
    h = 0.01
    xs, _ = RC.rk_run("LotkaVolterra", "RKF45", h, [1.0, 1.0], T, theta=[1.5, 1.0, 3.0, 1.0])
    rng = np.random.default_rng(5)
    ys = xs[1:, :1] + rng.normal(0.0, 0.03, (T, 1))
    return h, T, ys


def _oracle(levels, h, T, ys):
    out = []
    for c in levels:
        r = RC.ekf_run("LotkaVolterra", "RKF45", h, [[1.0, 1.0]], T, H=[[1.0, 0.0]], R_sqrt=[[0.03]], ys=ys,
                       correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T), cov="static_diagonal",
                       scale=float(c), theta_default=[1.5, 1.0, 3.0, 1.0], guard="intended")
        out.append(r["nll"][0] / T)
    return np.array(out)


def _sweep(backend, levels, h, T, ys):
    plan = U.make_plan(ode_id=N.ODE_LOTKA_VOLTERRA, solver_id=N.SOLVER_RKF45, step_size=h,
                       cov_fn_id=N.COV_STATIC_DIAGONAL, cov_scale=123.0)     # overridden per trajectory
    B = len(levels)
    r = U.run_ekf(backend, plan, np.repeat([[1.0, 1.0]], B, 0), T, H=[[1.0, 0.0]], R_sqrt=[[0.03]], ys=ys,
                  correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T), cov_scale_batch=levels,
                  nll_nan_to_num=True)
    return r["nll"] / T


def test_sweep_kernel_source_matches_oracle_per_level():
    h, T, ys = _problem()
    levels = np.logspace(-8, 0, 9)
    np.testing.assert_allclose(_sweep("hostemu", levels, h, T, ys), _oracle(levels, h, T, ys), rtol=1e-9)


def test_nan_to_num_semantics():
    """A NaN observation poisons that step's term: nan_to_num drops it (and everything after it,
    because the state itself became NaN) instead of returning NaN for the whole level."""
    h, T, ys = _problem(40)
    ys = ys.copy()
    ys[25, 0] = np.nan
    levels = np.array([1e-3, 1e-2])
    got = _sweep("hostemu", levels, h, T, ys)
    ref = _oracle(levels, h, 25, ys[:25]) * 25 / T          # the 25 healthy steps, mean over all T
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, ref, rtol=1e-9)


@pytest.mark.gpu
def test_calibration_runner_on_gpu():
    from ode_uncertainty_b200 import ode as O, runners, solvers as S
    from ode_uncertainty_b200.filters import SQRT_EKF
    h, T, ys = _problem()
    ys_x = np.concatenate([ys, np.zeros_like(ys)], axis=1)       # data file holds the full state; H picks x_0
    res = runners.calibration(filter_builder=SQRT_EKF(), solver_builder=S.RKF45(step_size=h), ode_builder=O.LotkaVolterra(),
                              x0="[[1.0, 1.0]]", t0=0.0, tN=T * h, ts_y=h * np.arange(1, T + 1), ys_x=ys_x,
                              measurement_matrix="[[1, 0]]", obs_noise_var=0.03 ** 2, min_noise_log=-8.0,
                              max_noise_log=0.0, num_noise_levels=33)
    assert res["noise_levels"].shape == (33,) and res["nll_conrad"].shape == (33,)
    np.testing.assert_allclose(res["nll_conrad"], _oracle(res["noise_levels"], h, T, ys), rtol=1e-9)
    ours = RC.ekf_run("LotkaVolterra", "RKF45", h, [[1.0, 1.0]], T, H=[[1.0, 0.0]], R_sqrt=[[0.03]], ys=ys,
                      correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T),
                      theta_default=[1.5, 1.0, 3.0, 1.0], guard="intended")["nll"][0] / T
    assert abs(res["nll_ours"] - ours) <= 1e-8 * abs(ours)
