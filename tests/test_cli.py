"""Config front end: YAML files with the reference's `class_path` / `init_args`
trees (structure of configs/ekf_trajectory_conrad_baseline/rkf45/lorenz.yaml and
configs/params/lotkavolterra4.yaml) select the B200 plugins."""
import os

import numpy as np
import pytest

from ode_uncertainty_b200 import cli

C1_YAML = """
output: {out}
filter_builder:
  class_path: src.filters.SQRT_EKF
  init_args:
    cov_update_fn_builder:
      class_path: src.covariance_update_functions.DiagonalCovarianceUpdate
      init_args:
        scale: 1.0
    static_cov_update_fn_builder:
      class_path: src.covariance_update_functions.StaticDiagonalCovarianceUpdate
      init_args:
        scale: 1.0
solver_builder:
  class_path: src.solvers.RKF45
  init_args:
    step_size: 0.01
ode_builder:
  class_path: src.ode.Lorenz
x0: '[[1.0, 1.0, 1.0]]'
P0: null
t0: 0.0
tN: 1.0
y_path: null
measurement_matrix: '[[1, 0, 0], [0, 1, 0], [0, 0, 1]]'
obs_noise_var: 0.0
save_interval: 1
disable_pbar: false
"""


def test_config_tree_builds_the_b200_plugins(tmp_path):
    from ode_uncertainty_b200 import filters, ode, solvers
    from ode_uncertainty_b200.covariance_update_functions import DiagonalCovarianceUpdate
    p = tmp_path / "c1.yaml"
    p.write_text(C1_YAML.format(out=str(tmp_path / "res" / "lorenz.h5")))
    cfg = cli.load_config(str(p), {"tN": "2.5", "save_interval": "5"})
    assert isinstance(cfg["filter_builder"], filters.SQRT_EKF)
    assert isinstance(cfg["filter_builder"].cov_update_fn_builder, DiagonalCovarianceUpdate)
    assert isinstance(cfg["solver_builder"], solvers.RKF45) and cfg["solver_builder"].h == 0.01
    assert isinstance(cfg["ode_builder"], ode.Lorenz)
    assert cfg["tN"] == 2.5 and cfg["save_interval"] == 5 and "disable_pbar" not in cfg
    assert cfg["output"].endswith("lorenz.npz")


def test_unknown_plugins_fail_loudly(tmp_path):
    p = tmp_path / "bad.yaml"
    p.write_text("solver_builder:\n  class_path: src.solvers.DiffraxSolverBuilder\n  init_args: {name: Tsit5, step_size: 0.01}\n")
    with pytest.raises(ValueError, match="DiffraxSolverBuilder"):      # only Kvaerno3 / ImplicitEuler are served
        cli.load_config(str(p))
    p.write_text("solver_builder:\n  class_path: src.solvers.NoSuchSolver\n")
    with pytest.raises(ValueError, match="NoSuchSolver"):
        cli.load_config(str(p))
    p.write_text("ode_builder:\n  class_path: somewhere.else.Thing\n")
    with pytest.raises(ValueError, match="no B200 plugin"):
        cli.load_config(str(p))


def test_nested_schedule_and_hh_builder(tmp_path):
    from ode_uncertainty_b200 import noise_schedules, ode
    p = tmp_path / "pe.yaml"
    p.write_text("""
gamma_noise_schedule:
  class_path: src.noise_schedules.LinearDecaySchedule
  init_args:
    init_noise_log: -2.0
    decay_rate: 3
ode_builder:
  class_path: src.ode.MultiCompartmentHodgkinHuxley
  init_args:
    model: reduced-1
    num_compartments: 2
params_range:
  g_Na: [0.5, 80.0]
num_processes: 4
""")
    cfg = cli.load_config(str(p))
    assert isinstance(cfg["gamma_noise_schedule"], noise_schedules.LinearDecaySchedule)
    assert cfg["gamma_noise_schedule"].step(1) == pytest.approx(1e-5)
    assert isinstance(cfg["ode_builder"], ode.MultiCompartmentHodgkinHuxley) and cfg["ode_builder"].state_dim == 14
    assert cfg["params_range"] == {"g_Na": [0.5, 80.0]} and "num_processes" not in cfg


@pytest.mark.gpu
def test_run_filter_from_yaml_equals_direct_call(tmp_path):
    from ode_uncertainty_b200 import ode as O, runners, solvers as S
    from ode_uncertainty_b200.filters import SQRT_EKF
    out = tmp_path / "res" / "lorenz.h5"
    p = tmp_path / "c1.yaml"
    p.write_text(C1_YAML.format(out=str(out)))
    res = cli.main(["run_filter", "--config", str(p)])
    ref = runners.run_filter(filter_builder=SQRT_EKF(), solver_builder=S.RKF45(step_size=0.01), ode_builder=O.Lorenz(),
                             x0="[[1.0, 1.0, 1.0]]", t0=0.0, tN=1.0, save_interval=1)
    np.testing.assert_array_equal(res["x"], ref["x"])
    np.testing.assert_array_equal(res["P"], ref["P"])
    saved = np.load(str(out)[:-3] + ".npz")
    assert saved["x"].shape == (101, 1, 1, 3) and set(["t", "x", "eps", "P_sqrt"]).issubset(saved.files)


@pytest.mark.gpu
def test_baseline_parameter_estimation_from_yaml(tmp_path):
    """configs/params_baseline/*.yaml shape (scripts/run_parameter_estimation_baseline.py:40-262) with an
    embedded RK solver: observations from an .npz `y_path`, results with the reference's dataset names."""
    from ode_uncertainty_b200 import Plan, _native as N, runners
    truth, h, T = np.array([1.5, 1.0, 3.0, 1.0]), 0.01, 300
    xs = runners.solve_trajectory(Plan(N.ODE_LOTKA_VOLTERRA, N.SOLVER_RKF45, h), [1.0, 1.0], T, theta_shared=truth)
    yp = tmp_path / "obs.npz"
    np.savez(yp, t=h * np.arange(1, T + 1), x=xs[1:] + np.random.default_rng(2).normal(0, 0.05, (T, 2)))
    out = tmp_path / "res" / "lv.h5"
    p = tmp_path / "base.yaml"
    p.write_text(f"""
output: {out}
solver_builder:
  class_path: src.solvers.RKF45
  init_args:
    step_size: {h}
ode_builder:
  class_path: src.ode.LotkaVolterra
x0: '[[1.0, 1.0]]'
t0: 0.0
tN: {T * h}
y_path: {yp}
measurement_matrix: '[[1, 0], [0, 1]]'
params_range:
  alpha: [0.5, 3.0]
  beta: [0.3, 2.0]
  gamma: [1.0, 5.0]
  delta: [0.3, 2.0]
params_optimized:
  alpha: true
  beta: true
  gamma: false
  delta: true
obs_noise_var: 0.0025
initial_state_parametrized: false
lbfgs_maxiter: 80
num_random_runs: 3
seed: 621
num_processes: 4
disable_pbar: true
verbose: false
""")
    res = cli.main(["run_parameter_estimation_baseline", "optimize", "--config", str(p)])
    assert res["params_optims"].shape == (3, 3) and list(res["params_name"]) == ["alpha", "beta", "delta"]
    best = int(np.argmin(res["nll_optims"]))
    np.testing.assert_allclose(res["params_optims"][best], [1.5, 1.0, 1.0], rtol=0.05)
    saved = np.load(str(out)[:-3] + ".npz")
    assert set(["params_inits", "params_optims", "params_default", "params_name", "nll_optims", "num_lbfgs_iters",
                "num_nll_evals", "num_nll_jac_evals"]).issubset(saved.files)


@pytest.mark.gpu
def test_evaluate_subcommand_grid_matches_single_evaluations(tmp_path):
    """`run_parameter_estimation evaluate` (scripts/run_parameter_estimation.py:311-537): NLL on the
    parameter grid for every tempering stage, the whole grid as ONE batch per stage; each grid value must
    equal the single-parameter-set evaluation (Oracle-B) at that point."""
    from oracle import ref_cpp as RC
    from ode_uncertainty_b200 import Plan, _native as N, runners
    truth, h, T = np.array([1.5, 1.0, 3.0, 1.0]), 0.01, 120
    xs = runners.solve_trajectory(Plan(N.ODE_LOTKA_VOLTERRA, N.SOLVER_RKF45, h), [1.0, 1.0], T, theta_shared=truth)
    ys = xs[1:] + np.random.default_rng(2).normal(0, 0.05, (T, 2))
    yp = tmp_path / "obs.npz"
    np.savez(yp, t=h * np.arange(1, T + 1), x=ys)
    p = tmp_path / "eval.yaml"
    p.write_text(f"""
output: {tmp_path / "eval.h5"}
filter_builder:
  class_path: src.filters.SQRT_EKF
  init_args:
    disable_cov_update: true
solver_builder:
  class_path: src.solvers.RKF45
  init_args:
    step_size: {h}
ode_builder:
  class_path: src.ode.LotkaVolterra
x0: '[[1.0, 1.0]]'
t0: 0.0
tN: {T * h}
y_path: {yp}
measurement_matrix: '[[1, 0], [0, 1]]'
params_range:
  alpha: [0.5, 3.0]
  beta: [0.3, 2.0]
  gamma: [1.0, 5.0]
  delta: [0.3, 2.0]
params_optimized:
  alpha: true
  beta: false
  gamma: true
  delta: false
num_param_evals:
  alpha: 5
  beta: 1
  gamma: 4
  delta: 1
num_tempering_stages: 2
final_gamma_zero: true
obs_noise_var: 0.0025
gamma_noise_schedule:
  class_path: src.noise_schedules.LinearDecaySchedule
  init_args:
    init_noise_log: -2.0
    decay_rate: 2
gamma_noise_weights: '[1, 1]'
""")
    res = cli.main(["run_parameter_estimation", "evaluate", "--config", str(p)])
    assert res["param_evals"].shape == (20, 2) and res["nll_evals"].shape == (2, 20)
    assert res["gammas"].tolist() == [1e-2, 0.0] and res["timings"].shape == (39,)
    np.testing.assert_allclose(np.unique(res["param_evals"][:, 0]), np.linspace(0.5, 3.0, 5))     # alpha (sorted keys: alpha, gamma)
    np.testing.assert_allclose(np.unique(res["param_evals"][:, 1]), np.linspace(1.0, 5.0, 4))
    for gi in (0, 7, 19):
        th = np.array([res["param_evals"][gi, 0], 1.0, res["param_evals"][gi, 1], 1.0])           # builder order alpha, beta, gamma, delta
        for si, g in enumerate(res["gammas"]):
            o = RC.ekf_run("LotkaVolterra", "RKF45", h, [[1.0, 1.0]], T, theta=th, Q_sqrt=np.eye(2), gamma_sqrt=g ** 0.5,
                           H=np.eye(2), R_sqrt=np.eye(2) * 0.05, ys=ys, correct_flags=np.ones(T, np.uint8),
                           xy_index_map=np.arange(T), disable=True, guard="intended")
            assert abs(res["nll_evals"][si, gi] - o["nll"][0]) <= 1e-8 * abs(o["nll"][0])
