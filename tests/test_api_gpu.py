"""GPU: the reference-shaped plugin API end to end (run_filter / unroll / nll / single steps)."""
import os

import numpy as np
import pytest
import torch

import cases
from oracle import ref_cpp as RC
from oracle import ref_torch as R
from ode_uncertainty_b200 import ode as O
from ode_uncertainty_b200 import runners, solvers as S
from ode_uncertainty_b200.covariance_update_functions import DiagonalCovarianceUpdate
from ode_uncertainty_b200.filters import (SQRT_EKF, ParticleFilter, parametrized_solver_handle, solver_handle)

pytestmark = pytest.mark.gpu


def test_run_filter_c1_config_matches_oracle():
    """configs/ekf_trajectory_conrad_baseline/rkf45/lorenz.yaml (C1), shortened to 200 steps."""
    spec = cases.CASES["c1_lorenz_rkf45_predict"]
    gold = cases.load_golden("c1_lorenz_rkf45_predict")
    res = runners.run_filter(filter_builder=SQRT_EKF(cov_update_fn_builder=DiagonalCovarianceUpdate(1.0)),
                             solver_builder=S.RKF45(step_size=0.01), ode_builder=O.Lorenz(),
                             x0="[[1.0, 1.0, 1.0]]", t0=0.0, tN=2.0, save_interval=1)
    assert res["x"].shape == (201, 1, 1, 3) and res["t"].shape == (201, 1)
    assert res["P_sqrt"].shape == (201, 1, 3, 3) and res["y_hat"].shape == (201, 1, 0)
    out = {"traj": {"t": res["t"][:, 0], "x": res["x"].reshape(201, 1, 3), "eps": res["eps"].reshape(201, 1, 3),
                    "P": res["P"], "y_hat": res["y_hat"], "S": np.zeros((201, 1, 0, 0))},
           "nll": res["nll"], "xT": res["x"][-1].reshape(1, 3), "PT": res["P"][-1]}
    cases.compare(out, gold, spec)
    Ps = res["P_sqrt"][:, 0]
    np.testing.assert_allclose(Ps @ Ps.transpose(0, 2, 1), res["P"][:, 0], rtol=1e-10, atol=1e-40)


def test_run_filter_with_observation_file(tmp_path):
    spec = cases.CASES["lv_rkf45_temper_q_only"]
    m = cases.materialize(spec)
    # data file like scripts/run_ode_solver.py writes it: t, x at the observation times
    ts = m["t0"] + m["h"] * np.arange(1, m["T"] + 1)
    path = os.path.join(tmp_path, "obs.npz")
    np.savez(path, t=ts, x=m["ys"].numpy())
    out_path = os.path.join(tmp_path, "res.npz")
    res = runners.run_filter(output=out_path, filter_builder=SQRT_EKF(), solver_builder=S.RKF45(0.01),
                             ode_builder=O.LotkaVolterra(), x0="[[1.0, 1.0]]", t0=0.0, tN=m["T"] * m["h"],
                             y_path=path, measurement_matrix="[[1, 0], [0, 1]]", obs_noise_var=0.1)
    saved = np.load(out_path)
    assert set(["t", "x", "eps", "P_sqrt", "y_hat", "S_sqrt"]).issubset(saved.files)
    b = RC.ekf_run("LotkaVolterra", "RKF45", 0.01, [[1.0, 1.0]], m["T"], H=np.eye(2), R_sqrt=np.eye(2) * 0.1 ** 0.5,
                   ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"],
                   theta_default=[1.5, 1.0, 3.0, 1.0], save_interval=1)
    np.testing.assert_allclose(res["x"].reshape(m["T"] + 1, 2), b["traj"]["x"][:, 0], rtol=1e-9)
    assert abs(res["nll"][0] - b["nll"][0]) < 1e-8 * abs(b["nll"][0])


def test_single_step_predict_and_correct_match_oracle():
    ode_b, sb = O.Lorenz(), S.RKF45(step_size=0.01)
    fb = SQRT_EKF()
    sb.setup(ode_b.build(), ode_b.params)
    solver = solver_handle(sb)
    predict, correct = fb.build_predict(), fb.build_correct()
    cov = fb.build_cov_update_fn()
    P0s = np.eye(3) * 1e-3
    st = fb.init_state(sb.init_state(0.0, np.array([[1.0, 1.0, 1.0]])), P0s, np.zeros((3, 3)), 0.0, np.eye(2) * 0.1)
    # oracle
    ode, params, _ = R.ODES["Lorenz"]
    ost = R.init_state(0.0, torch.tensor([[1.0, 1.0, 1.0]]), torch.tensor(P0s), torch.zeros(3, 3), 0.0, torch.eye(2) * 0.1)
    osolver = lambda t, x: R.rk_step(ode, params, "RKF45", 0.01, t, x)
    H = torch.eye(3)[[0, 2]]
    for k in range(3):
        st = predict(solver, cov, st)
        ost = R.ekf_predict(osolver, R.cov_update_sqrt("diagonal", 1.0), False, ost)
        np.testing.assert_allclose(st["x"].cpu().numpy(), ost["x"].numpy(), rtol=1e-13)
        np.testing.assert_allclose(st["P"].cpu().numpy()[0], (ost["P_sqrt"][0] @ ost["P_sqrt"][0].T).numpy(), rtol=1e-9)
        y = torch.tensor([1.0 + 0.1 * k, 1.0])
        st["y"], ost["y"] = y, y
        st = correct(H.numpy(), st)
        ost = R.ekf_correct(H, ost)
        np.testing.assert_allclose(st["x"].cpu().numpy(), ost["x"].numpy(), rtol=1e-12)
        np.testing.assert_allclose(st["P"].cpu().numpy()[0], (ost["P_sqrt"][0] @ ost["P_sqrt"][0].T).numpy(), rtol=1e-9)
        np.testing.assert_allclose(st["S"].cpu().numpy()[0], (ost["S_sqrt"][0] @ ost["S_sqrt"][0].T).numpy(), rtol=1e-12)
        nlg = R.negative_log_gaussian_sqrt(ost["y"], ost["y_hat"][0], ost["S_sqrt"][0])
        np.testing.assert_allclose(float(st["nlg"][0]), float(nlg), rtol=1e-11)


def test_solver_step_and_ode_rhs_match_oracle():
    dev = torch.device("cuda:0")
    for name in ("Lorenz", "VanDerPol", "LotkaVolterra", "Pendulum", "LCAO", "HodgkinHuxley/reduced-1",
                 "HodgkinHuxley/full", "HodgkinHuxley/reduced-4"):
        ode, params, shape = R.ODES[name]
        b = O.HodgkinHuxley(model=name.split("/")[1]) if "/" in name else getattr(O, name)()
        x = torch.tensor(cases.CASES[{"Lorenz": "c1_lorenz_rkf45_predict", "VanDerPol": "vdp_rkf45_obs",
                                      "LotkaVolterra": "lv_heun_none", "Pendulum": "pendulum_rkf45_obs",
                                      "LCAO": "lcao_rkf45_obs", "HodgkinHuxley/reduced-1": "hh_r1_rkf45_temper",
                                      "HodgkinHuxley/full": "hh_full_rkf45_small_h",
                                      "HodgkinHuxley/reduced-4": "hh_r4_rkf45_temper"}[name]]["x0"]).reshape(shape)
        f = b.build()
        got = f(12.0, x.to(dev), b.params).cpu()
        # gate derivatives vanish at the HH steady state: compare against the size of the terms (O(1))
        np.testing.assert_allclose(got.numpy(), ode(torch.tensor(12.0), x, params).numpy(), rtol=1e-12, atol=1e-15)
        for cls in (S.RKF45, S.Dopri65, S.BS32, S.HeunEuler):
            sb = cls(step_size=0.01)
            sb.setup(f, b.params)
            nxt = sb.build()({"t": torch.tensor(11.0), "x": x.to(dev)})
            t1, x1, e1 = R.rk_step(ode, params, cls.tableau, 0.01, torch.tensor(11.0), x)
            np.testing.assert_allclose(nxt["x"].cpu().numpy(), x1.numpy(), rtol=1e-13)
            np.testing.assert_allclose(nxt["eps"].cpu().numpy(), e1.numpy(), rtol=1e-6, atol=16 * np.spacing(np.abs(x1.numpy()).max()))
            assert float(nxt["t"]) == float(t1)


def test_batched_nll_matches_reference_code_and_oracle_b():
    """ekf_nll at the default parameters equals the reference's own nll() fixture; perturbed
    parameter sets (one launch) equal per-set Oracle-B runs."""
    for name in ("lv_rkf45_temper_q_only", "lv_rkf45_temper_eps_plus_q", "hh_r4_rkf45_temper"):
        spec = cases.CASES[name]
        m = cases.materialize(spec)
        ref = dict(np.load(os.path.join(cases.GOLDEN, f"ref_{name}.npz")))
        if name.startswith("lv"):
            ob = O.LotkaVolterra()
        else:
            ob = O.HodgkinHuxley(model="reduced-4")
        fb = SQRT_EKF(cov_update_fn_builder=DiagonalCovarianceUpdate(spec.get("scale", 1.0)),
                      disable_cov_update=spec.get("disable", False))
        sb = S.RKF45(step_size=0.01)
        keys_s, sizes, perm = runners.param_layout(ob)
        lo = {k: np.minimum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(-1) for k in keys_s}
        hi = {k: np.maximum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(-1) for k in keys_s}
        lo_f = np.concatenate([lo[k] for k in keys_s]); hi_f = np.concatenate([hi[k] for k in keys_s])
        np.testing.assert_allclose(lo_f, ref["lo"]); np.testing.assert_allclose(hi_f, ref["hi"])
        default_sorted = np.concatenate([ob.params[k].reshape(-1) for k in keys_s])
        pn0 = (default_sorted - lo_f) / (hi_f - lo_f)
        rng = np.random.default_rng(5)
        pn = np.clip(pn0[None, :] + 0.02 * rng.normal(size=(6, pn0.size)), 0, 1)
        pn[0] = pn0
        common = dict(params_min=lo, params_max=hi, x0=m["x0"].numpy(), P0_sqrt=m["P0s"].numpy(), t0=m["t0"],
                      num_steps=m["T"], measurement_matrix=m["H"].numpy(), ys=m["ys"].numpy(),
                      correct_flags=m["flags"], xy_index_map=m["ymap"], R_sqrt=m["Rs"].numpy(),
                      Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5)
        nll = runners.ekf_nll(fb, sb, ob, params_norm=pn, **common).cpu().numpy()
        assert abs(nll[0] - float(ref["nll_fn"])) <= 1e-9 * abs(float(ref["nll_fn"]))
        theta_sorted = pn * (hi_f - lo_f) + lo_f
        b = RC.ekf_run(spec["ode"], "RKF45", 0.01, np.repeat(m["x0"].reshape(1, -1).numpy(), 6, 0), m["T"],
                       t0=m["t0"], P0_sqrt=m["P0s"].numpy(), theta=theta_sorted[:, perm], Q_sqrt=m["Q"].numpy(),
                       gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(),
                       correct_flags=m["flags"], xy_index_map=m["ymap"], scale=spec.get("scale", 1.0),
                       disable=spec.get("disable", False))
        np.testing.assert_allclose(nll, b["nll"], rtol=1e-8)


def test_particle_filter_unroll_shapes_and_particle_zero():
    fb = ParticleFilter(num_particles=64)
    tr = runners.run_filter(filter_builder=fb, solver_builder=S.RKF45(0.01), ode_builder=O.Lorenz(),
                            x0="[[1.0, 1.0, 1.0]]", tN=0.5, save_interval=5)
    assert tr["x"].shape == (11, 64, 1, 3) and tr["t"].shape == (11, 64)
    xs, _ = RC.rk_run("Lorenz", "RKF45", 0.01, [1.0, 1.0, 1.0], 50, theta=[10.0, 8.0 / 3, 28.0])
    np.testing.assert_allclose(tr["x"][:, 0, 0], xs[::5], rtol=1e-12)
    assert np.abs(tr["x"][-1, 1:] - tr["x"][-1, 0]).max() > 0


def test_solve_trajectory_equals_predict_only_filter_mean():
    """runners.solve_trajectory (plain RK steps of the ensemble kernel, scripts/run_ode_solver.py:56-74)
    gives the same solution as the mean of the prediction-only filter."""
    import numpy as np
    import torch
    from ode_uncertainty_b200 import Plan, _native as N, ekf_run, runners
    for ode_id, kw, x0 in ((N.ODE_LORENZ, {}, [1.0, 1.0, 1.0]),
                           (N.ODE_HODGKIN_HUXLEY, dict(ode_variant=1), [-70.0, 0.01, 0.99, 0.01, 0.03, 0.0, 0.4])):
        plan = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01, **kw)
        xs = runners.solve_trajectory(plan, x0, 300)
        r = ekf_run(plan, torch.tensor([x0], dtype=torch.float64, device="cuda"), 300, P0_sqrt=np.eye(plan.n) * 1e-12,
                    save_interval=1, save_keys=("x",), want_final=False)
        ref = r.traj["x"][:, 0, :].cpu().numpy()
        assert xs.shape == ref.shape == (301, plan.n)
        np.testing.assert_allclose(xs, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())


def test_single_compartment_multi_hh_matches_reference_code():
    """MultiCompartmentHodgkinHuxley(num_compartments = 1) (src/ode/hodgkin_huxley.py:284-439 with an empty
    coupling_coeffs): the plan is mapped onto the single-compartment kernels (csrc/api.cu).  Right-hand side and
    initial value against the reference's OWN class (fixture: oracle/make_golden_mhh1.py)."""
    from ast import literal_eval
    g = dict(np.load(os.path.join(cases.GOLDEN, "ref_mhh1_rhs.npz")))
    args = literal_eval(str(g["args"]))
    dev = torch.device("cuda:0")
    for model in ("reduced-1", "reduced-4", "full"):
        b = O.MultiCompartmentHodgkinHuxley(model=model, num_compartments=1, **args)
        f = b.build()
        n = g[f"x_{model}"].shape[1]
        assert b.shape == (1, n)
        np.testing.assert_allclose(b.build_initial_value(np.array([[-68.0]]), b.params).reshape(-1), g[f"x0_{model}"], rtol=1e-12)
        for x, t, want in zip(g[f"x_{model}"], g[f"t_{model}"], g[f"f_{model}"]):
            got = f(float(t), torch.tensor(x).reshape(1, n).to(dev), b.params).cpu().numpy().reshape(-1)
            np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)
