"""Implicit (stiff) solver plugins - `DiffraxSolverBuilder(name="Kvaerno3" | "ImplicitEuler")`,
src/solvers/diffrax_solver.py:16-140, SURVEY 8(f) N3.

PARITY UNPINNED: diffrax is a third-party dependency that exists neither under /root/reference nor
in this image, and no test of the reference touches it.  What is checked here is the CUDA
restatement (csrc/dirk.cuh) against this repository's OWN restatement of the published methods
(oracle/ref_torch.py::dirk_step, Newton run to machine precision, derivatives by torch.func) - mean,
covariance, NLL and the NLL gradient - plus method-level properties that do not depend on either
restatement (order of accuracy against an analytic solution, stability where the explicit solver
diverges)."""
import os

import numpy as np
import pytest
import torch

import cases
import util as U
from oracle import ref_torch as R

IMPL = list(cases.IMPLICIT_CASES)


def _gold(name):
    """Oracle-A trajectories are slow (Newton + jacfwd under vmap(jvp)); cached under tests/golden."""
    p = os.path.join(cases.GOLDEN, f"oracleA_implicit_{name}.npz")
    if not os.path.exists(p) or os.environ.get("ODEU_WRITE_GOLDEN"):
        out = cases.run_oracle(cases.IMPLICIT_CASES[name])
        np.savez_compressed(p, **out)
    return dict(np.load(p))


def _run(backend, name, batch=1, **extra):
    spec = cases.IMPLICIT_CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, axis=0)
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
              theta_shared=R.flat_params(m["params"]).numpy(), save_interval=1)
    if m["L"] > 0:
        kw.update(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
    kw.update(extra)
    return U.run_ekf(backend, plan, x0, m["T"], **kw), m, plan


@pytest.mark.parametrize("backend", ["hostemu", pytest.param("gpu", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("name", IMPL)
def test_implicit_filter_matches_own_oracle(name, backend):
    out, m, _ = _run(backend, name, batch=3 if backend == "gpu" else 1)
    cases.compare(out, _gold(name), cases.IMPLICIT_CASES[name], b=0)
    assert np.all(out["traj"]["eps"] == 0.0)                      # diffrax_solver.py:128


def test_kvaerno3_is_third_order_and_l_stable_on_the_linear_test_equation():
    """Independent of any restatement: x' = -lambda x through the Lotka-Volterra plugin is not
    available, so use the pendulum's small-angle limit?  No - use the solver directly on Van der Pol with
    damping 0 (x'' = -x, harmonic oscillator): the error against cos(t) must fall 8x per halving of h."""
    from ode_uncertainty_b200 import Plan, _native as N
    errs = []
    for h in (0.1, 0.05, 0.025):
        plan = Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_KVAERNO3, step_size=h)
        T = int(round(2.0 / h))
        out = U.run_ekf("hostemu", plan, np.array([[1.0, 0.0]]), T, theta_shared=[0.0], P0_sqrt=np.zeros((2, 2)))
        errs.append(abs(out["xT"][0, 0] - np.cos(T * h)))
    assert 6.5 < errs[0] / errs[1] < 9.5 and 6.5 < errs[1] / errs[2] < 9.5, errs
    # the step Jacobian of the harmonic oscillator is the stability matrix R(hA): P0 = I propagates to R R^T
    plan = Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_KVAERNO3, step_size=0.1, disable_cov_update=True)
    out = U.run_ekf("hostemu", plan, np.array([[1.0, 0.0]]), 1, theta_shared=[0.0], P0_sqrt=np.eye(2))
    Rm = np.linalg.cholesky(out["PT"][0])
    assert abs(np.linalg.det(out["PT"][0]) - 1.0) < 1e-3          # |R| ~ 1 on the imaginary axis at h = 0.1


def test_implicit_solver_survives_where_rkf45_diverges():
    from ode_uncertainty_b200 import Plan, _native as N
    x0 = np.array([[2.0, 0.0]])
    kw = dict(theta_shared=[200.0], P0_sqrt=np.zeros((2, 2)))
    exp = U.run_ekf("hostemu", Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.05), x0, 400, **kw)
    imp = U.run_ekf("hostemu", Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_KVAERNO3, step_size=0.05), x0, 400, **kw)
    assert not np.isfinite(exp["xT"]).all()
    assert np.isfinite(imp["xT"]).all() and abs(imp["xT"][0, 0]) <= 2.01


@pytest.mark.parametrize("backend", ["hostemu", pytest.param("gpu", marks=pytest.mark.gpu)])
def test_implicit_nll_gradient_matches_own_oracle_autograd(backend):
    """d NLL / d theta through the implicit steps (parameter tangents by the implicit function theorem,
    d J / d theta through the nested dual) against reverse-mode autograd through the oracle's Newton."""
    name = "vdp_stiff_kvaerno3_obs"
    spec = dict(cases.IMPLICIT_CASES[name])
    spec["T"] = 12
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    th = torch.tensor(spec["theta"], dtype=torch.float64, requires_grad=True)

    def loss(thv):
        params = {"damping": thv[0]}
        st = R.init_state(m["t0"], m["x0"], m["P0s"], m["Q"], m["gamma"] ** 0.5, m["Rs"])
        _, nll, _ = R.run_filter(m["ode"], params, m["solver"], m["h"], m["cov"], m["scale"], m["disable"], st, m["H"],
                                 m["ys"][:spec["T"]], m["flags"], m["ymap"], spec["T"], spec["T"], guard="intended")
        return nll

    val = loss(th)
    (g_ref,) = torch.autograd.grad(val, th)
    nll, g = U.run_grad(backend, plan, m["x0"].reshape(1, -1).numpy(), spec["T"], [0], t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                        theta_shared=spec["theta"], Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
                        R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy()[:spec["T"]], correct_flags=m["flags"][:spec["T"]],
                        xy_index_map=m["ymap"][:spec["T"]])
    assert abs(nll[0] - float(val)) <= 1e-9 * abs(float(val))
    assert abs(g[0, 0] - float(g_ref[0])) <= 1e-6 * abs(float(g_ref[0])), (g[0, 0], float(g_ref[0]))


def test_yaml_front_end_selects_the_implicit_plugin():
    from ode_uncertainty_b200 import cli, solvers, _native as N
    sb = cli.instantiate({"class_path": "src.solvers.DiffraxSolverBuilder", "init_args": {"name": "Kvaerno3", "step_size": 0.01}})
    assert isinstance(sb, solvers.DiffraxSolverBuilder) and sb.solver_id == N.SOLVER_KVAERNO3 and sb.h == 0.01
    assert solvers.DiffraxSolverBuilder().solver_id == N.SOLVER_IMPLICIT_EULER          # the builder's default (:19)
    with pytest.raises(ValueError, match="serves"):
        solvers.DiffraxSolverBuilder(name="Tsit5")


@pytest.mark.gpu
def test_shipped_hodgkin_huxley_configuration_shape_runs_end_to_end(tmp_path):
    """The shape of configs/params/hodgkinhuxley11_full.yaml - full Hodgkin-Huxley model, `Kvaerno3`,
    step_size 0.01, 11 optimised parameters, process-noise tempering - through the YAML front end
    (`cli run_parameter_estimation optimize`), on a shortened horizon: the explicit RKF45 diverges on
    this model at h = 0.01, the implicit plugin must not, and the optimiser must lower
    the NLL from its random start."""
    import yaml
    from ode_uncertainty_b200 import cli, ode as O, solvers, runners, Plan, _native as N
    ob = O.HodgkinHuxley(model="full")
    plan = Plan(N.ODE_HODGKIN_HUXLEY, N.SOLVER_KVAERNO3, 0.01, ode_variant=0)
    t0, tN = 9.5, 12.5                                  # across the stimulus onset at t = 10
    T = int(round((tN - t0) / 0.01))
    x0 = ob.build_initial_value(np.array([[-70.0]]), ob.params).reshape(-1)
    xs = runners.solve_trajectory(plan, x0, T, t0=t0, theta_shared=ob.flat_params(ob.params))
    assert np.isfinite(xs).all() and xs[:, 0].max() > -68.0          # the stimulus (t >= 10) depolarises the membrane
    rng = np.random.default_rng(621)
    np.savez(tmp_path / "obs.npz", t=t0 + 0.01 * np.arange(1, T + 1), x=(xs[1:] + rng.normal(0, 0.1 ** 0.5, xs[1:].shape)).reshape(T, 1, 8))
    rng_keys = {"C": [0.4, 3.0], "A": [1.9e-05, 30.2e-05], "g_Na": [0.5, 80.0], "E_Na": [50.0, 100.0], "g_K": [1.0e-04, 15.0],
                "E_K": [-110.0, -70.0], "g_leak": [1.0e-04, 0.6], "E_leak": [-100.0, -35.0], "V_T": [-90.0, -40.0],
                "g_M": [1.0e-04, 0.6], "tau_max": [50.0, 5000.0], "g_L": [-1.0e-04, 0.6], "E_Ca": [100.0, 150.0],
                "g_T": [-1.0e-04, 0.6], "V_x": [0.0, 4.0]}
    opt = {k: k not in ("C", "A", "tau_max", "V_x") for k in rng_keys}                   # 11 optimised parameters
    cfg = {"output": str(tmp_path / "out.h5"),
           "filter_builder": {"class_path": "src.filters.SQRT_EKF", "init_args": {
               "cov_update_fn_builder": {"class_path": "src.covariance_update_functions.DiagonalCovarianceUpdate", "init_args": {"scale": 1.0}},
               "disable_cov_update": True}},
           "solver_builder": {"class_path": "src.solvers.DiffraxSolverBuilder", "init_args": {"name": "Kvaerno3", "step_size": 0.01}},
           "ode_builder": {"class_path": "src.ode.HodgkinHuxley", "init_args": {"model": "full"}},
           "x0": "[[-70.0]]", "P0": None, "t0": t0, "tN": tN, "y_path": str(tmp_path / "obs.h5"),
           "measurement_matrix": "[[1, 0, 0, 0, 0, 0, 0, 0]]", "params_range": rng_keys, "params_optimized": opt,
           "gamma_noise_weights": "[1, 1, 1, 1, 1, 1, 1, 1]", "num_tempering_stages": 2, "final_gamma_zero": True,
           "obs_noise_var": 0.1, "lbfgs_maxiter": 3, "num_random_runs": 3, "seed": 7, "initial_state_parametrized": False,
           "gamma_noise_schedule": {"class_path": "src.noise_schedules.LinearDecaySchedule", "init_args": {"init_noise_log": -2.0, "decay_rate": 3}}}
    (tmp_path / "cfg.yaml").write_text(yaml.safe_dump(cfg))
    res = cli.main(["run_parameter_estimation", "optimize", "--config", str(tmp_path / "cfg.yaml")])
    assert res["params_optims"].shape == (3, 2, 11)
    assert np.isfinite(res["nll_optims"]).all()
    assert (res["num_nll_evals"] > 0).all()
