"""CPU suite: the kernel SOURCE (ekf_trajectory, compiled for the host by tests/host_emu.cu)
against Oracle-A golden vectors.  No GPU needed; the product path itself is covered by
tests/test_parity_gpu.py through the C ABI."""
import numpy as np
import pytest

import cases

FAST = ["c1_lorenz_rkf45_predict", "lorenz_rkf45_obs_full", "lorenz_dopri65_obs_partial",
        "lorenz_bs32_outer", "lorenz_heun_static", "vdp_rkf45_obs", "vdp_dopri65_predict",
        "lv_rkf45_temper_q_only", "lv_rkf45_temper_eps_plus_q", "lv_bs32_gamma0_final_stage",
        "lv_heun_none", "pendulum_rkf45_obs", "lcao_rkf45_obs", "hh_r4_rkf45_temper",
        "hh_r1_rkf45_temper", "hh_full_rkf45_small_h", "c3_mhh_r1_rkf45_temper"]


@pytest.mark.parametrize("name", FAST)
def test_kernel_source_matches_oracle_golden(name):
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    out = cases.run_product("hostemu", spec, save_interval=1, batch=2)
    cases.compare(out, gold, spec, b=0)
    cases.compare(out, gold, spec, b=1)


def test_save_interval_strides_like_reference():
    # scripts/run_filter.py:219-222: concat(initial, states)[::save_interval]
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    full = cases.run_product("hostemu", spec, save_interval=1)
    strided = cases.run_product("hostemu", spec, save_interval=7)
    for k in ("t", "x", "eps", "P", "y_hat", "S"):
        np.testing.assert_array_equal(strided["traj"][k], full["traj"][k][::7])
    none = cases.run_product("hostemu", spec, save_interval=0)
    assert none["traj"] is None
    np.testing.assert_array_equal(none["xT"], full["xT"])
    np.testing.assert_array_equal(none["nll"], full["nll"])


def test_resume_from_saved_state_is_bitwise():
    """Checkpoint/resume: running T1 then T2 steps from the returned (t, x, P) equals T1+T2."""
    import util as U
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    kw = dict(Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
              R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy())
    x0 = m["x0"].reshape(1, -1).numpy()
    T, T1 = m["T"], 60
    whole = U.run_ekf("hostemu", plan, x0, T, t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                      correct_flags=m["flags"], xy_index_map=m["ymap"], **kw)
    a = U.run_ekf("hostemu", plan, x0, T1, t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                  correct_flags=m["flags"], xy_index_map=m["ymap"], **kw)
    b = U.run_ekf("hostemu", plan, a["xT"], T - T1, t0=a["tT"], P0=a["PT"],
                  correct_flags=m["flags"][T1:], xy_index_map=m["ymap"][T1:], **kw)
    np.testing.assert_array_equal(b["xT"], whole["xT"])
    np.testing.assert_array_equal(b["PT"], whole["PT"])
    assert abs((a["nll"] + b["nll"])[0] - whole["nll"][0]) <= 1e-12 * abs(whole["nll"][0])


def test_per_trajectory_params_and_observations():
    """theta [B,p] and ys [T_obs,B,L] per trajectory give the same result as B separate runs."""
    import util as U
    spec = cases.CASES["lv_rkf45_temper_q_only"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    rng = np.random.default_rng(3)
    B = 3
    theta = plan.default_params[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, plan.p)))
    ys = m["ys"].numpy()[:, None, :] + 0.01 * rng.normal(size=(m["ys"].shape[0], B, m["L"]))
    x0 = m["x0"].reshape(1, -1).numpy() + 0.05 * rng.normal(size=(B, m["n"]))
    common = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(),
                  gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(),
                  correct_flags=m["flags"], xy_index_map=m["ymap"])
    batched = U.run_ekf("hostemu", plan, x0, m["T"], theta=theta, ys=ys, ys_per_trajectory=True, **common)
    for b in range(B):
        one = U.run_ekf("hostemu", plan, x0[b:b + 1], m["T"], theta_shared=theta[b], ys=ys[:, b], **common)
        np.testing.assert_array_equal(batched["xT"][b], one["xT"][0])
        np.testing.assert_array_equal(batched["PT"][b], one["PT"][0])
        np.testing.assert_array_equal(batched["nll"][b], one["nll"][0])


def test_nan_propagates_per_trajectory_without_error():
    """Numerical failure is never an error: a NaN trajectory stays NaN, its
    neighbours are untouched."""
    import util as U
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), 3, axis=0)
    x0[1, 0] = np.nan
    out = U.run_ekf("hostemu", plan, x0, 20, P0_sqrt=m["P0s"].numpy(), H=m["H"].numpy(),
                    R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                    xy_index_map=m["ymap"])
    assert np.isnan(out["xT"][1]).all() and np.isnan(out["nll"][1])
    assert np.isfinite(out["xT"][[0, 2]]).all()
    np.testing.assert_array_equal(out["xT"][0], out["xT"][2])


def test_invalid_arguments_raise_value_error():
    import util as U
    from ode_uncertainty_b200 import Plan, _native as N
    with pytest.raises(ValueError):
        Plan(ode_id=N.ODE_LORENZ, solver_id=17, step_size=0.01)
    with pytest.raises(ValueError):
        Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=-1.0)
    with pytest.raises(ValueError):
        Plan(ode_id=N.ODE_MULTI_HH, solver_id=N.SOLVER_RKF45, step_size=0.01, ode_variant=1,
             num_compartments=5)
    plan = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
    with pytest.raises(ValueError):  # L > n
        U.run_ekf("hostemu", plan, np.ones((1, 3)), 5, H=np.ones((4, 3)), R_sqrt=np.eye(4),
                  ys=np.zeros((5, 4)), correct_flags=np.ones(5, np.uint8),
                  xy_index_map=np.arange(5))


@pytest.mark.parametrize("name", ["lorenz_rkf45_obs_full", "vdp_rkf45_obs", "lv_rkf45_temper_q_only",
                                  "c1_lorenz_rkf45_predict"])
def test_time_segmented_run_is_bitwise_identical(name):
    """The dynamic scheduler cuts a run into (block, time-segment) items whose state travels
    through the workspace: replayed sequentially on the host, results equal the one-piece run."""
    import util as U
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), 35, axis=0) + 0.01 * np.arange(35)[:, None]
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5)
    if m["L"] > 0:
        kw.update(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                  xy_index_map=m["ymap"])
    whole = U.run_ekf("hostemu", plan, x0, m["T"], **kw)
    seg = U.run_ekf("hostemu", plan, x0, m["T"], segmented=True, **kw)
    for k in ("xT", "PT", "epsT", "yhatT", "ST"):
        np.testing.assert_array_equal(seg[k], whole[k])
    # the log-determinant part of the NLL is accumulated as a pivot product per segment
    # (LogProd, ekf_core.cuh), so segmenting changes its rounding, not the state
    np.testing.assert_allclose(seg["nll"], whole["nll"], rtol=1e-12, atol=1e-12)
    assert seg["tT"] == whole["tT"]
    if m["L"] > 0:
        # the final-state y_hat / S come from the LAST observation step, which only that step writes
        # (ekf_thread.cuh): observations every third step that stop after 60 % of the run - the last
        # one lies in an earlier time segment than the end of the run
        T = m["T"]
        flags = np.zeros(T, np.uint8)
        flags[2:int(0.6 * T):3] = 1
        last = int(np.nonzero(flags)[0][-1])
        kw2 = dict(kw, correct_flags=flags, xy_index_map=np.minimum(np.arange(T), m["ys"].shape[0] - 1))
        whole = U.run_ekf("hostemu", plan, x0, T, **kw2)
        seg = U.run_ekf("hostemu", plan, x0, T, segmented=True, **kw2)
        upto = U.run_ekf("hostemu", plan, x0, last + 1, **{**kw2, "correct_flags": flags[:last + 1],
                                                            "xy_index_map": kw2["xy_index_map"][:last + 1]})
        for k in ("xT", "PT", "yhatT", "ST"):
            np.testing.assert_array_equal(seg[k], whole[k])
        assert np.abs(whole["ST"]).max() > 0
        np.testing.assert_array_equal(whole["yhatT"], upto["yhatT"])      # = those of the run cut at the last observation
        np.testing.assert_array_equal(whole["ST"], upto["ST"])


@pytest.mark.parametrize("name", ["hh_r1_rkf45_temper", "hh_full_rkf45_small_h", "c3_mhh_r1_rkf45_temper"])
def test_row_kernel_minimal_and_full_output_runs_match_oracle(name):
    """Medium-size systems (row kernel, ekf_rows.cuh): the run that requests only nll / xT / PT against the run with
    the full output contract and against the oracle's final state, on a ragged batch of perturbed initial conditions."""
    import util as U
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    B = 6
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0)
    x0[1:, 0] += 0.3 * np.arange(1, B)
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
              H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
              xy_index_map=m["ymap"])
    coop = U.run_ekf("hostemu", plan, x0, m["T"], minimal=True, **kw)
    thr = U.run_ekf("hostemu", plan, x0, m["T"], **kw)
    np.testing.assert_allclose(coop["nll"], thr["nll"], rtol=1e-11)
    np.testing.assert_allclose(coop["xT"], thr["xT"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(coop["PT"], thr["PT"], rtol=1e-9, atol=1e-12 * np.abs(thr["PT"]).max())
    assert abs(coop["nll"][0] - float(gold["nll"])) <= 1e-9 * abs(float(gold["nll"]))
    np.testing.assert_allclose(coop["xT"][0], gold["x"][-1], rtol=1e-10)
    np.testing.assert_allclose(coop["PT"][0], gold["P"][-1], rtol=1e-9, atol=1e-12 * np.abs(gold["P"][-1]).max())


def test_two_wide_row_kernel_source_is_bitwise_equal_to_one_wide():
    """ODEU_ROWS_2WIDE=1 (two trajectories per thread through the V2d scalar, vec2.cuh): the same
    statements on two lanes, so the replayed source must agree bit for bit with the one-wide
    kernel.  Runs in subprocesses because the switch is read once per process."""
    import subprocess
    import sys
    import tempfile
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests")
import cases, util as U
spec = cases.CASES["c3_mhh_r1_rkf45_temper"]; m = cases.materialize(spec); plan = cases.make_plan_for(spec)
B = 11
x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0); x0[1:, 0] += 0.3 * np.arange(1, B)
kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
          R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
c = U.run_ekf("hostemu", plan, x0, m["T"], minimal=True, **kw)
np.save(sys.argv[1], np.concatenate([c["nll"].ravel(), c["xT"].ravel(), c["PT"].ravel()]))
'''
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as td:
        outs = []
        for tag, env in (("one", {}), ("two", {"ODEU_ROWS_2WIDE": "1"})):
            out = os.path.join(td, tag + ".npy")
            e = dict(os.environ); e.pop("ODEU_ROWS_2WIDE", None); e.update(env)
            subprocess.check_call([sys.executable, "-c", code, out], cwd=root, env=e)
            outs.append(np.load(out))
    np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize("backend", ["hostemu", pytest.param("gpu", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("name", ["hh_r1_rkf45_temper", "c3_mhh_r1_rkf45_temper"])
def test_row_kernel_serves_the_full_output_contract(name, backend):
    """run_filter on the Hodgkin-Huxley family (strided trajectory slots, eps, pre-update
    y_hat / S, t, resume from a per-trajectory P0) is served by the row-parallel kernel, not by the slow
    one-column-per-pass thread kernel: every output against the Oracle-A fixture, strided saves equal the
    dense ones, and a run cut in two (second leg resumed from xT / PT / tT) equals the uncut run bitwise."""
    import util as U
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    batch = 19 if backend == "gpu" else 2
    dense = cases.run_product(backend, spec, save_interval=1, batch=batch)
    cases.compare(dense, gold, spec, b=batch - 1)
    strided = cases.run_product(backend, spec, save_interval=7, batch=batch)
    for k in ("t", "x", "eps", "P", "y_hat", "S"):
        np.testing.assert_array_equal(strided["traj"][k], dense["traj"][k][::7], err_msg=k)
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, axis=0)
    kw = dict(Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy())
    T, T1 = m["T"], 17
    a = U.run_ekf(backend, plan, x0, T1, t0=m["t0"], P0_sqrt=m["P0s"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"], **kw)
    b = U.run_ekf(backend, plan, a["xT"], T - T1, t0=a["tT"], P0=a["PT"], correct_flags=m["flags"][T1:], xy_index_map=m["ymap"][T1:], **kw)
    np.testing.assert_array_equal(b["xT"], dense["xT"])
    np.testing.assert_array_equal(b["PT"], dense["PT"])
    assert b["tT"] == dense["tT"]
    np.testing.assert_array_equal(b["yhatT"], dense["traj"]["y_hat"][-1])
    np.testing.assert_array_equal(b["ST"], dense["traj"]["S"][-1])
    np.testing.assert_array_equal(b["epsT"], dense["traj"]["eps"][-1])
