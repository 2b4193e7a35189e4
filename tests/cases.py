"""Parity case table shared by the CPU (host-emulated kernel source) and GPU (product path) suites.

Each case is run through Oracle-A (oracle/ref_torch.py) once by oracle/make_golden.py and cached
as tests/golden/<name>.npz; the tests compare against the cached vectors (and, for a subset,
against a live oracle run so a stale fixture cannot hide a regression).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_torch as R  # noqa: E402
from ode_uncertainty_b200 import _native as N  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SOLVERS = {"RKF45": N.SOLVER_RKF45, "Dopri65": N.SOLVER_DOPRI65, "BS32": N.SOLVER_BS32,
           "HeunEuler": N.SOLVER_HEUN_EULER, "Kvaerno3": N.SOLVER_KVAERNO3, "ImplicitEuler": N.SOLVER_IMPLICIT_EULER}
COVS = {"diagonal": N.COV_DIAGONAL, "outer": N.COV_OUTER, "static_diagonal": N.COV_STATIC_DIAGONAL}
ODE_IDS = {
    "Lorenz": (N.ODE_LORENZ, 0, 0), "VanDerPol": (N.ODE_VAN_DER_POL, 0, 0),
    "LotkaVolterra": (N.ODE_LOTKA_VOLTERRA, 0, 0), "Pendulum": (N.ODE_PENDULUM, 0, 0),
    "LCAO": (N.ODE_LCAO, 2, 0), "HodgkinHuxley/full": (N.ODE_HODGKIN_HUXLEY, 0, 0),
    "HodgkinHuxley/reduced-1": (N.ODE_HODGKIN_HUXLEY, 1, 0),
    "HodgkinHuxley/reduced-4": (N.ODE_HODGKIN_HUXLEY, 4, 0),
    "MultiHH/reduced-1/2": (N.ODE_MULTI_HH, 1, 2), "MultiHH/reduced-4/2": (N.ODE_MULTI_HH, 4, 2),
}


def _hh_x0(model):
    return R.hh_initial_value(model, -70.0, R.ODES[f"HodgkinHuxley/{model}"][1]).flatten().tolist()


def _mhh_x0(model):
    p = R.multi_hh_default_params(2)
    xs = []
    for c in range(2):
        pc = {k: torch.broadcast_to(v, (2,) + tuple(v.shape[1:]))[c] for k, v in p.items()}
        xs += R.hh_initial_value(model, -70.0, pc).flatten().tolist()
    return xs


# name -> spec.  `obs`: None (prediction only) or (rows of H as state indices, cadence).
# guard="intended": cases on which the reference's sign-sensitive zero-gain guard
# (src/filters/sqrt_ekf.py:351-353) fires for a healthy negative Householder factor
# and silently drops observations; there the oracle is run with the guard's intended meaning
# (what the CUDA path implements) and tests/test_oracle.py pins that the verbatim guard differs.
CASES = {
    # C1: configs/ekf_trajectory_conrad_baseline/rkf45/lorenz.yaml, shortened
    "c1_lorenz_rkf45_predict": dict(ode="Lorenz", solver="RKF45", T=200, x0=[1., 1., 1.]),
    "lorenz_rkf45_obs_full": dict(ode="Lorenz", solver="RKF45", T=150, x0=[1., 1., 1.],
                                  obs=([0, 1, 2], 1), Rvar=1e-3),
    "lorenz_dopri65_obs_partial": dict(ode="Lorenz", solver="Dopri65", T=100, x0=[-3., 2., 20.],
                                       obs=([0, 2], 1), Rvar=1e-2),
    "lorenz_bs32_outer": dict(ode="Lorenz", solver="BS32", T=100, x0=[1., 1., 1.], cov="outer",
                              scale=2.0, obs=([1], 3), Rvar=1e-2, guard="intended"),
    "lorenz_heun_static": dict(ode="Lorenz", solver="HeunEuler", T=100, x0=[1., 1., 1.],
                               cov="static_diagonal", scale=1e-3, obs=([0, 1, 2], 1), Rvar=1e-3),
    "vdp_rkf45_obs": dict(ode="VanDerPol", solver="RKF45", T=150, x0=[2., 10.], t0=10.0,
                          obs=([0], 1), Rvar=1e-3),
    "vdp_dopri65_predict": dict(ode="VanDerPol", solver="Dopri65", T=100, x0=[2., 10.], t0=10.0),
    # process-noise tempering branches (configs/params/lotkavolterra4.yaml shape)
    "lv_rkf45_temper_q_only": dict(ode="LotkaVolterra", solver="RKF45", T=150, x0=[1., 1.],
                                   disable=True, Qw=[1., 1.], gamma=1e-2, obs=([0, 1], 1), Rvar=0.1),
    "lv_rkf45_temper_eps_plus_q": dict(ode="LotkaVolterra", solver="RKF45", T=150, x0=[1., 1.],
                                       disable=False, Qw=[1., 0.5], gamma=1e-5, scale=3.0,
                                       obs=([0], 1), Rvar=0.1),
    "lv_bs32_gamma0_final_stage": dict(ode="LotkaVolterra", solver="BS32", T=100, x0=[1., 1.],
                                       disable=True, Qw=[1., 1.], gamma=0.0, obs=([0, 1], 1),
                                       Rvar=0.1),
    "lv_heun_none": dict(ode="LotkaVolterra", solver="HeunEuler", T=80, x0=[1., 1.], disable=True,
                         obs=([0], 2), Rvar=0.1, P0=0.01, guard="intended"),
    "pendulum_rkf45_obs": dict(ode="Pendulum", solver="RKF45", T=120, x0=[0.7853981633974483, 0.0],
                               obs=([0], 1), Rvar=1e-2),
    "lcao_rkf45_obs": dict(ode="LCAO", solver="RKF45", T=100, x0=[1., -0.5, 0., 0.3],
                           obs=([0, 1], 1), Rvar=1e-2),
    "hh_r4_rkf45_temper": dict(ode="HodgkinHuxley/reduced-4", solver="RKF45", T=80, t0=9.6,
                               x0=_hh_x0("reduced-4"), disable=True, Qw=[1.] * 4, gamma=1e-2,
                               obs=([0], 1), Rvar=0.1),
    "hh_r1_rkf45_temper": dict(ode="HodgkinHuxley/reduced-1", solver="RKF45", T=80, t0=9.6,
                               x0=_hh_x0("reduced-1"), disable=True, Qw=[1.] * 7, gamma=1e-5,
                               obs=([0], 1), Rvar=0.1),
    "hh_full_rkf45_small_h": dict(ode="HodgkinHuxley/full", solver="RKF45", T=60, t0=9.99, h=2.5e-4,
                                  x0=_hh_x0("full"), disable=True, Qw=[1.] * 8, gamma=1e-8,
                                  obs=([0], 1), Rvar=0.1),
    # C3 shape: 2-compartment reduced-1, n=14, L=2 (configs/params/hodgkinhuxley6_c2_r1.yaml)
    "c3_mhh_r1_rkf45_temper": dict(ode="MultiHH/reduced-1/2", solver="RKF45", T=40, t0=9.8,
                                   x0=_mhh_x0("reduced-1"), disable=True, Qw=[1.] * 14, gamma=1e-2,
                                   obs=([0, 7], 1), Rvar=0.1),
}


# Implicit solver plugins (DiffraxSolverBuilder).  PARITY UNPINNED: diffrax is absent, the
# oracle is this repository's own restatement (oracle/ref_torch.py::dirk_step).  Kept apart from CASES
# (no reference-code fixture can exist); run against a live / cached Oracle-A by tests/test_implicit.py.
IMPLICIT_CASES = {
    # stiff regime of Van der Pol (damping 50): explicit RKF45 at h = 0.05 blows up, Kvaerno3 does not
    "vdp_stiff_kvaerno3_obs": dict(ode="VanDerPol", solver="Kvaerno3", T=30, h=0.05, x0=[2., 0.], theta=[50.0],
                                   obs=([0], 1), Rvar=1e-3, disable=True, Qw=[1., 1.], gamma=1e-4),
    "lv_implicit_euler_predict": dict(ode="LotkaVolterra", solver="ImplicitEuler", T=25, h=0.05, x0=[1., 1.]),
    # the shipped Hodgkin-Huxley shape (configs/params/hodgkinhuxley11_full.yaml): full model, Kvaerno3, h = 0.01
    "hh_full_kvaerno3_temper": dict(ode="HodgkinHuxley/full", solver="Kvaerno3", T=25, t0=9.9, h=0.01,
                                    x0=_hh_x0("full"), disable=True, Qw=[1.] * 8, gamma=1e-5, obs=([0], 1), Rvar=0.1),
}


def _lcao_x0(D, seed=7):
    """BASELINE config 5: positions ~ N(0, 1), velocities 0."""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, 1.0, D).tolist() + [0.0] * D


# Large-state cases (n = 2 D; served by odeu_ekf_dense_run).  Kept apart from CASES: their
# fixtures store the covariance of the LAST step only (tests/test_dense.py).
DENSE_CASES = {
    # C5 shape at reduced size: first 16 positions observed at every step, P0 = 1e-6 I
    "c5_lcao64_rkf45_obs": dict(ode="LCAO/64", solver="RKF45", T=12, x0=_lcao_x0(64), P0=1e-6,
                                obs=(list(range(16)), 1), Rvar=1e-2),
    # C5 itself (n = 256), few steps
    "c5_lcao128_rkf45_obs": dict(ode="LCAO/128", solver="RKF45", T=5, x0=_lcao_x0(128), P0=1e-6,
                                 obs=(list(range(16)), 1), Rvar=1e-2),
    # tempering branch (diagonal Q), scattered observed components
    "lcao64_dopri65_temper": dict(ode="LCAO/64", solver="Dopri65", T=8, x0=_lcao_x0(64, 11), P0=1e-4,
                                  disable=True, Qw=[0.5 + 0.01 * i for i in range(128)], gamma=1e-3,
                                  obs=([3, 70, 127, 64, 10], 1), Rvar=1e-1, guard="intended"),
    # prediction only, embedded error as process noise, Heun-Euler (b[1] = [0.5, 0] kept verbatim)
    "lcao64_bs32_predict": dict(ode="LCAO/64", solver="BS32", T=10, x0=_lcao_x0(64, 3), P0=1e-6, scale=2.0),
}


def ode_and_params(name):
    if name.startswith("LCAO/"):
        D = int(name.split("/")[1])
        return R.ode_lcao, R.ODES["LCAO"][1], (2, D)
    if name.startswith("MultiHH/"):
        _, model, nc = name.split("/")
        return R.make_ode_multi_hh(model, int(nc)), R.multi_hh_default_params(int(nc)), (1, int(nc) * R.HH_DIM[model])
    return R.ODES[name]


def materialize(spec):
    """Expand a spec into concrete arrays (deterministic; seeds fixed)."""
    ode, params, shape = ode_and_params(spec["ode"])
    if "theta" in spec:        # non-default parameters, flat in builder order
        flat, o, params = torch.tensor(spec["theta"], dtype=torch.float64), 0, dict(params)
        for k, v in list(params.items()):
            params[k] = flat[o:o + v.numel()].reshape(v.shape)
            o += v.numel()
    n = shape[0] * shape[1]
    h = spec.get("h", 0.01)
    T = spec["T"]
    t0 = spec.get("t0", 0.0)
    x0 = torch.tensor(spec["x0"], dtype=torch.float64).reshape(shape)
    P0s = torch.eye(n) * (1e-12 if "P0" not in spec else spec["P0"] ** 0.5)
    Qw = spec.get("Qw")
    Q = torch.zeros(n, n) if Qw is None else torch.diag(torch.tensor(Qw, dtype=torch.float64))
    gamma = spec.get("gamma", 0.0)
    obs = spec.get("obs")
    if obs is not None:
        rows, every = obs
        H = torch.eye(n)[rows]
        L = len(rows)
        Rvar = spec.get("Rvar", 1e-3)
        Rs = torch.eye(L) * Rvar ** 0.5
        xs, _ = R.run_rk(ode, params, spec["solver"], h, t0, x0, T)
        rng = np.random.default_rng(8)
        ys_all = xs[1:].reshape(T, -1) @ H.T + torch.tensor(rng.normal(0, Rvar ** 0.5, (T, L)))
        steps = np.arange(1, T + 1)
        flags = (steps % every == 0).astype(np.uint8)
        # observations exist only at flagged steps, like a coarser data file
        idx = np.nonzero(flags)[0]
        ys = ys_all[idx]
        ymap = np.zeros(T, dtype=np.int64)
        ymap[idx] = np.arange(len(idx))
    else:
        H, L, Rs = torch.eye(n), 0, torch.zeros(0, 0)
        ys = torch.zeros(1, 0)
        flags, ymap = np.zeros(T, dtype=np.uint8), np.zeros(T, dtype=np.int64)
    return dict(ode=ode, params=params, shape=shape, n=n, h=h, T=T, t0=t0, x0=x0, P0s=P0s, Q=Q,
                gamma=gamma, H=H, L=L, Rs=Rs, ys=ys, flags=flags, ymap=ymap,
                cov=spec.get("cov", "diagonal"), scale=spec.get("scale", 1.0),
                disable=spec.get("disable", False), solver=spec["solver"])


def run_oracle(spec):
    m = materialize(spec)
    st = R.init_state(m["t0"], m["x0"], m["P0s"], m["Q"], m["gamma"] ** 0.5, m["Rs"])
    traj, nll, quirks = R.run_filter(m["ode"], m["params"], m["solver"], m["h"], m["cov"], m["scale"],
                                     m["disable"], st, m["H"], m["ys"], m["flags"], m["ymap"], m["T"], 1,
                                     guard=spec.get("guard", "reference"))
    out = {k: v.numpy() for k, v in traj.items()}
    out["nll"] = np.array(float(nll))
    out["guard_mismatch_steps"] = np.array(quirks["guard_mismatch_steps"])
    out["guard_fired_steps"] = np.array(quirks["guard_fired_steps"])
    return out


def golden_path(name):
    return os.path.join(GOLDEN, f"oracleA_{name}.npz")


def load_golden(name):
    p = golden_path(name)
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} missing: run `python oracle/make_golden.py`")
    return dict(np.load(p))


def make_plan_for(spec):
    from ode_uncertainty_b200 import Plan
    if spec["ode"].startswith("LCAO/"):
        ode_id, variant, nc = N.ODE_LCAO, int(spec["ode"].split("/")[1]), 0
    else:
        ode_id, variant, nc = ODE_IDS[spec["ode"]]
    return Plan(ode_id=ode_id, solver_id=SOLVERS[spec["solver"]], step_size=spec.get("h", 0.01),
                ode_variant=variant, num_compartments=nc, cov_fn_id=COVS[spec.get("cov", "diagonal")],
                cov_scale=spec.get("scale", 1.0), disable_cov_update=spec.get("disable", False))


def run_product(backend, spec, save_interval=1, batch=1):
    """Run the kernel source (hostemu) or the product path (gpu) on the case; `batch` replicates
    the trajectory so coalesced multi-thread paths are exercised with identical expected output."""
    import util as U
    m = materialize(spec)
    plan = make_plan_for(spec)
    theta = R.flat_params(m["params"]).numpy()
    assert theta.shape[0] == plan.p
    np.testing.assert_allclose(theta, plan.default_params, rtol=0, atol=0)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, axis=0)
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
              save_interval=save_interval)
    if m["L"] > 0:
        kw.update(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(),
                  correct_flags=m["flags"], xy_index_map=m["ymap"])
    return U.run_ekf(backend, plan, x0, m["T"], **kw)


def compare(out, gold, spec, b=0):
    """Tolerances (float64), per saved step, relative to the largest entry of the quantity at
    that step:
      t: exact; x, y_hat: 1e-10; S: 1e-9; nll: 1e-9
      eps: |d eps| <= 16 ulp(max|x|) + 1e-9 eps  -- eps is a difference of two O(|x|) numbers
           (rksolver.py:146-147), so it carries cancellation noise of a few ulp(x) in ANY
           implementation, the reference included;
      P: 1e-9 when eps does not enter P (disable_cov_update), else 5e-6 (eps^2 enters P and
         inherits the relative cancellation noise of eps, ~ulp(x)/eps ~ 1e-7).
    """
    tr = out["traj"]
    np.testing.assert_array_equal(tr["t"], gold["t"])
    Ts = gold["t"].shape[0]

    def stepwise(a, g, rtol, what, atol_step=None):
        a = a.reshape(Ts, -1)
        g = g.reshape(Ts, -1)
        if g.shape[1] == 0:
            return
        scale = np.max(np.abs(g), axis=1)
        err = np.max(np.abs(a - g), axis=1)
        tol = rtol * scale + (0 if atol_step is None else atol_step)
        bad = np.nonzero(~(err <= tol))[0]
        assert bad.size == 0, f"{what}: step {bad[0]} err {err[bad[0]]:.3e} tol {np.broadcast_to(tol, err.shape)[bad[0]]:.3e}"

    xmax = np.max(np.abs(gold["x"]), axis=1)
    ulp = np.spacing(xmax)
    stepwise(tr["x"][:, b], gold["x"], 1e-10, "x")
    stepwise(tr["eps"][:, b], gold["eps"], 1e-9, "eps", atol_step=16 * ulp)
    eps_in_P = not spec.get("disable", False)
    # absolute floor for P when eps^2 enters it: d(eps^2) <= 2 eps d_eps + d_eps^2 with
    # d_eps = 16 ulp(x), accumulated over the steps taken so far
    emax = np.max(np.abs(gold["eps"]), axis=1)
    s = spec.get("scale", 1.0) if spec.get("cov", "diagonal") != "static_diagonal" else 0.0
    p_floor = np.cumsum(s * s * (2 * emax * 16 * ulp + (16 * ulp) ** 2)) if eps_in_P else None
    stepwise(tr["P"][:, b], gold["P"], 5e-6 if eps_in_P else 1e-9, "P", atol_step=p_floor)
    if gold["y_hat"].size:
        stepwise(tr["y_hat"][:, b], gold["y_hat"], 1e-10, "y_hat")
        stepwise(tr["S"][:, b], gold["S"], 5e-6 if eps_in_P else 1e-9, "S", atol_step=p_floor)
    g = float(gold["nll"])
    assert abs(out["nll"][b] - g) <= 1e-9 * max(1.0, abs(g)), f"nll {out['nll'][b]} vs {g}"
    # final-state outputs agree with the last saved slot
    np.testing.assert_array_equal(out["xT"][b], tr["x"][-1, b])
    np.testing.assert_array_equal(out["PT"][b], tr["P"][-1, b])
