"""guard_mode "reference": the factor-form kernels (csrc/ekf_sqrt.cuh) carry P_sqrt through the same
three Householder QRs per step as the reference (src/utils.py:233-274, LAPACK dlarfg signs), so the
zero-gain guard `all(S_sqrt < 1e-16)` (src/filters/sqrt_ekf.py:350-353) fires exactly where the
reference's does - including the healthy all-negative factors on which the reference silently drops
an observation.  Pinned against (i) fixtures produced by the REFERENCE'S OWN CODE
(tests/golden/ref_*.npz), now including the two quirk cases, (ii) a LAPACK-backed Oracle-A run of the
headline Van der Pol workload in which the guard fires, (iii) Oracle-B in reference mode."""
import os

import numpy as np
import pytest

import cases
import util as U
from oracle import ref_cpp as RC
from oracle import ref_torch as R

SMALL = [n for n, s in cases.CASES.items() if cases.ODE_IDS[s["ode"]][0] not in (5, 6)]   # n <= 4
QUIRK = [n for n in SMALL if cases.CASES[n].get("guard") == "intended"]


def _ref(name):
    p = os.path.join(cases.GOLDEN, f"ref_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} missing (python oracle/make_golden_ref.py)")
    return dict(np.load(p))


def _run(backend, spec, guard, **kw):
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    batch = kw.pop("batch", 1)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, axis=0)
    a = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, guard=guard)
    if m["L"] > 0:
        a.update(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                 xy_index_map=m["ymap"])
    a.update(kw)
    return U.run_ekf(backend, plan, x0, m["T"], **a), m


@pytest.mark.parametrize("name", SMALL)
def test_factor_form_source_matches_reference_code(name):
    """Every n <= 4 parity case, quirk cases included, against what the reference's own code computed."""
    spec = cases.CASES[name]
    out, _ = _run("hostemu", spec, "reference", save_interval=1)
    cases.compare(out, _ref(name), spec, b=0)


@pytest.mark.parametrize("name", QUIRK)
def test_quirk_cases_fire_like_the_reference_and_differ_from_intended(name):
    spec = cases.CASES[name]
    out, m = _run("hostemu", spec, "reference", save_interval=1)
    live = dict(spec)
    live.pop("guard")
    gold = cases.run_oracle(live)                           # Oracle-A, verbatim guard (LAPACK)
    assert int(out["guard_fired"][0]) == int(gold["guard_fired_steps"]) > 0
    assert int(out["guard_mismatch"][0]) == int(gold["guard_mismatch_steps"]) > 0
    intended, _ = _run("hostemu", spec, "intended", save_interval=1)
    assert U.rel_err(intended["traj"]["x"], out["traj"]["x"]) > 1e-6     # the quirk is observable
    # factor-form arithmetic with the INTENDED predicate = the full-covariance kernels
    fi, _ = _run("hostemu", spec, "intended_factor", save_interval=1)
    assert int(fi["guard_fired"][0]) == 0
    assert U.rel_err(fi["traj"]["x"], intended["traj"]["x"]) < 1e-10


def _c2(system, B, T):
    import bench
    w = bench.workload_inputs(system, B, T, 0)
    ys = bench.observations(system, T, w)
    return w, ys


def _c2_kw(w, ys, T):
    return dict(t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"], R_sqrt=w["R_sqrt"], ys=ys,
                correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T, dtype=np.int64))


def test_headline_vdp_guard_fires_where_lapack_fires():
    """BASELINE config 2, Van der Pol half, trajectory 0: Oracle-A (LAPACK QR) drops the observation
    of step 1006; the factor-form kernel source and Oracle-B must reproduce the whole trajectory."""
    from ode_uncertainty_b200 import Plan, _native as N
    g = dict(np.load(os.path.join(cases.GOLDEN, "oracleA_c2_vdp_guardref.npz")))
    T = int(g["T"])
    assert np.nonzero(g["fired"])[0].tolist() == [1006]
    w, ys = _c2("VanDerPol", 8, T)
    plan = Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.01)
    kw = _c2_kw(w, ys, T)
    out = U.run_ekf("hostemu", plan, w["x0"][:1], T, guard="reference", save_interval=1, **kw)
    assert int(out["guard_fired"][0]) == 1 and int(out["guard_mismatch"][0]) == 1
    np.testing.assert_allclose(out["traj"]["x"][1:, 0], g["x"], rtol=1e-10, atol=1e-10 * np.abs(g["x"]).max())
    # the factor itself, signs included
    np.testing.assert_allclose(out["PT_sqrt"][0], g["P_sqrt"][-1], rtol=1e-6, atol=1e-9 * np.abs(g["P_sqrt"][-1]).max())
    assert abs(out["nll"][0] - float(g["nll"])) <= 1e-9 * abs(float(g["nll"]))
    ob = RC.ekf_run("VanDerPol", "RKF45", 0.01, w["x0"][:1], T, theta=[5.0], guard="reference", save_interval=1, **kw)
    assert ob["guard_fired_steps"] == 1
    np.testing.assert_allclose(ob["traj"]["x"][1:, 0], g["x"], rtol=1e-10, atol=1e-10 * np.abs(g["x"]).max())
    # the intended-guard run differs from what the reference computes from step 1007 on
    oi = U.run_ekf("hostemu", plan, w["x0"][:1], T, guard="intended", save_interval=1, **kw)
    d = np.abs(oi["traj"]["x"][1:, 0] - g["x"]).max(axis=1)
    assert d[:1006].max() < 1e-10 and d[1006:].max() > 1e-8


@pytest.mark.parametrize("system", ["Lorenz", "VanDerPol"])
def test_headline_workload_factor_form_vs_oracle_b(system):
    """Several trajectories of the bench workload over 2,500 steps, static and (block, time-segment)
    scheduled replay: state, NLL and the guard's firing count against Oracle-B in reference mode."""
    from ode_uncertainty_b200 import Plan, _native as N
    B, T = 5, 2500
    w, ys = _c2(system, B, T)
    plan = Plan(ode_id=N.ODE_LORENZ if system == "Lorenz" else N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.01)
    kw = _c2_kw(w, ys, T)
    th = {"Lorenz": [10.0, 8.0 / 3, 28.0], "VanDerPol": [5.0]}[system]
    o = RC.ekf_run(system, "RKF45", 0.01, w["x0"], T, theta=th, guard="reference", **kw)
    for seg in (False, True):
        out = U.run_ekf("hostemu", plan, w["x0"], T, guard="reference", segmented=seg, **kw)
        assert int(out["guard_fired"].sum()) == o["guard_fired_steps"]
        assert int(out["guard_mismatch"].sum()) == o["guard_mismatch_steps"]
        np.testing.assert_allclose(out["xT"], o["xT"], rtol=1e-9, atol=1e-9 * np.abs(o["xT"]).max())
        np.testing.assert_allclose(out["PT"], o["PT"], rtol=1e-7, atol=1e-7 * np.abs(o["PT"]).max())
        np.testing.assert_allclose(out["nll"], o["nll"], rtol=1e-9)
    if system == "Lorenz":
        assert o["guard_fired_steps"] == 0 and o["fragile_qr_columns"] == 0
    else:
        assert o["guard_fired_steps"] > 0


def test_factor_form_resume_is_bitwise():
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    full, m = _run("hostemu", spec, "reference", batch=3)
    plan = cases.make_plan_for(spec)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), 3, axis=0)
    kw = dict(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), guard="reference")
    T, T1 = m["T"], 61
    a = U.run_ekf("hostemu", plan, x0, T1, P0_sqrt=m["P0s"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"], **kw)
    b = U.run_ekf("hostemu", plan, a["xT"], T - T1, t0=a["tT"], P0_sqrt_batch=a["PT_sqrt"],
                  correct_flags=m["flags"][T1:], xy_index_map=m["ymap"][T1:], **kw)
    np.testing.assert_array_equal(b["xT"], full["xT"])
    np.testing.assert_array_equal(b["PT_sqrt"], full["PT_sqrt"])
    np.testing.assert_array_equal(b["PT"], full["PT"])


def test_reference_guard_rejected_where_not_served():
    """n > 4 plans (row / thread kernels of the Hodgkin-Huxley family) keep the intended
    guard; asking them for the factor form is an error, not a silent downgrade."""
    spec = cases.CASES["hh_r4_rkf45_temper"]
    out, _ = _run("hostemu", spec, "intended")
    assert np.isfinite(out["nll"]).all()
    spec = cases.CASES["hh_r1_rkf45_temper"]
    with pytest.raises(ValueError, match="state dimension <= 4"):
        _run("hostemu", spec, "reference")


@pytest.mark.parametrize("L", [1, 2])
def test_exactly_singular_innovation_trips_the_guard(L):
    """P0 = 0, R = 0, no process noise: S = 0 exactly.  The reference's guard fires (all zeros pass
    `< 1e-16`), K = 0, the state is untouched (sqrt_ekf.py:351-353).  The leading-identity fast path
    used rsqrt(0) * 0 = NaN here and poisoned x and P (advisor, round 1)."""
    from ode_uncertainty_b200 import Plan, _native as N
    plan = Plan(ode_id=N.ODE_LOTKA_VOLTERRA, solver_id=N.SOLVER_RKF45, step_size=0.01, disable_cov_update=True)
    T = 5
    x0 = np.array([[1.0, 1.0], [1.2, 0.7]])
    H = np.eye(2)[:L]
    kw = dict(P0_sqrt=np.zeros((2, 2)), H=H, R_sqrt=np.zeros((L, L)), ys=np.ones((T, L)),
              correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T, dtype=np.int64))
    free = U.run_ekf("hostemu", plan, x0, T, P0_sqrt=np.zeros((2, 2)))          # prediction only
    for guard in ("intended", "reference"):
        out = U.run_ekf("hostemu", plan, x0, T, guard=guard, **kw)
        np.testing.assert_array_equal(out["xT"], free["xT"])
        np.testing.assert_array_equal(out["PT"], np.zeros((2, 2, 2)))
    # generic measurement update (H not a leading identity block) takes the same decision
    Hg = np.array([[0.0, 1.0]])
    kwg = dict(kw, H=Hg, R_sqrt=np.zeros((1, 1)), ys=np.ones((T, 1)))
    out = U.run_ekf("hostemu", plan, x0, T, **kwg)
    np.testing.assert_array_equal(out["xT"], free["xT"])



def test_exactly_singular_innovation_on_the_row_kernel():
    """Same degenerate update on the Hodgkin-Huxley row kernel (ekf_rows.cuh, n = 7, L = 1): P0 = 0, R = 0,
    no process noise.  rsqrt(0) used to poison x and P there (advisor, round 1: "apply the same check to
    the rows and grad variants"); now the pivot and its column are zero, the guard trips, K = 0 and the
    state equals the prediction-only run bit for bit."""
    from ode_uncertainty_b200 import Plan, _native as N
    from ode_uncertainty_b200 import ode as O
    ob = O.HodgkinHuxley(model="reduced-1")
    plan = Plan(N.ODE_HODGKIN_HUXLEY, N.SOLVER_RKF45, 0.01, ode_variant=1, disable_cov_update=True)
    n, T = plan.n, 4
    x0 = np.repeat(ob.build_initial_value(np.array([[-70.0]]), ob.params).reshape(1, -1), 3, axis=0)
    H = np.zeros((1, n)); H[0, 0] = 1.0
    kw = dict(P0_sqrt=np.zeros((n, n)), H=H, R_sqrt=np.zeros((1, 1)), ys=np.full((T, 1), -60.0),
              correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T, dtype=np.int64))
    free = U.run_ekf("hostemu", plan, x0, T, minimal=True, P0_sqrt=np.zeros((n, n)))
    out = U.run_ekf("hostemu", plan, x0, T, minimal=True, **kw)
    np.testing.assert_array_equal(out["xT"], free["xT"])
    np.testing.assert_array_equal(out["PT"], np.zeros((3, n, n)))

# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL)
def test_cuda_factor_form_matches_reference_code(name):
    spec = cases.CASES[name]
    out, _ = _run("gpu", spec, "reference", save_interval=1, batch=33)
    cases.compare(out, _ref(name), spec, b=32)
    if name in QUIRK:
        assert int(out["guard_fired"][32]) > 0


@pytest.mark.gpu
def test_cuda_headline_vdp_guard_fires_where_lapack_fires():
    from ode_uncertainty_b200 import Plan, _native as N
    g = dict(np.load(os.path.join(cases.GOLDEN, "oracleA_c2_vdp_guardref.npz")))
    T = int(g["T"])
    w, ys = _c2("VanDerPol", 8, T)
    plan = Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.01)
    out = U.run_ekf("gpu", plan, w["x0"], T, guard="reference", save_interval=1, **_c2_kw(w, ys, T))
    assert int(out["guard_fired"][0]) == 1
    np.testing.assert_allclose(out["traj"]["x"][1:, 0], g["x"], rtol=1e-10, atol=1e-10 * np.abs(g["x"]).max())
    np.testing.assert_allclose(out["PT_sqrt"][0], g["P_sqrt"][-1], rtol=1e-6, atol=1e-9 * np.abs(g["P_sqrt"][-1]).max())


@pytest.mark.gpu
def test_cuda_factor_form_scheduler_and_resume():
    """Dynamic (block, time-segment) scheduling of the factor-form kernel equals the static launch bit
    for bit (state, factor, guard counters); resume from PT_sqrt is bitwise."""
    import torch
    from ode_uncertainty_b200 import Plan, ekf_run, _native as N
    dev = torch.device("cuda:0")
    B, T = 65536, 400
    w, ys = _c2("VanDerPol", B, T)
    plan = Plan(ode_id=N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.01)
    kw = dict(t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"], R_sqrt=w["R_sqrt"], ys=torch.as_tensor(ys).to(dev),
              correct_flags=torch.ones(T, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T, device=dev),
              guard="reference")
    x0 = torch.as_tensor(w["x0"]).to(dev)
    a = ekf_run(plan, x0, T, dynamic=True, **kw)
    b = ekf_run(plan, x0, T, dynamic=False, **kw)
    for k in ("xT", "PT", "PT_sqrt", "epsT", "guard_fired", "guard_mismatch"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert torch.allclose(a.nll, b.nll, rtol=1e-12, atol=1e-12)
    T1 = 150
    kw1 = dict(kw, ys=kw["ys"][:T1], correct_flags=kw["correct_flags"][:T1], xy_index_map=kw["xy_index_map"][:T1])
    c = ekf_run(plan, x0[:4096], T1, **kw1)
    kw2 = dict(kw, ys=kw["ys"][T1:], correct_flags=kw["correct_flags"][T1:], xy_index_map=kw["xy_index_map"][:T - T1])
    kw2.pop("P0_sqrt"); kw2["t0"] = float(c.tT)
    d = ekf_run(plan, c.xT.contiguous(), T - T1, P0_sqrt_batch=c.PT_sqrt.contiguous(), **kw2)
    assert torch.equal(d.xT, b.xT[:4096]) and torch.equal(d.PT_sqrt, b.PT_sqrt[:4096])
