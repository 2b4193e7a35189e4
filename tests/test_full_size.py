"""Parity at BASELINE.json's FULL sizes (configs 2 and 3): the whole batch runs on the GPU at the
benchmark's horizon; a sample of trajectories / parameter sets is re-run by Oracle-B (the C++
restatement of the reference's square-root filter) over the same full horizon.  With an observation
at every step the filters contract onto the data, so the comparison stays meaningful for Lorenz-63
over all 10,000 steps (free-running chaos, SURVEY F7, does not apply).  Tolerances: 1e-8 relative
on the final mean, 1e-7 on the final covariance and the NLL (rounding accumulated over 10^4 steps
of two different factorisations: full covariance here, Householder square-root in the oracle)."""
import numpy as np
import pytest
import torch

from oracle import ref_cpp as RC

pytestmark = pytest.mark.gpu


def _bench_inputs(system, B, T):
    import bench
    w = bench.workload_inputs(system, B, T, 0)
    ys = bench.observations(system, T, w)
    return w, ys


@pytest.mark.parametrize("system", ["Lorenz", "VanDerPol"])
def test_config2_full_size_sample_against_oracle(system):
    from ode_uncertainty_b200 import Plan, _native as N, ekf_run
    dev = torch.device("cuda:0")
    B, T = 65536, 10000
    w, ys = _bench_inputs(system, B, T)
    plan = Plan(ode_id=N.ODE_LORENZ if system == "Lorenz" else N.ODE_VAN_DER_POL, solver_id=N.SOLVER_RKF45, step_size=0.01)
    flags, ymap = np.ones(T, np.uint8), np.arange(T, dtype=np.int64)
    r = ekf_run(plan, torch.as_tensor(w["x0"]).to(dev), T, t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"], R_sqrt=w["R_sqrt"],
                ys=torch.as_tensor(ys).to(dev), correct_flags=torch.as_tensor(flags).to(dev),
                xy_index_map=torch.as_tensor(ymap).to(dev), guard="reference")
    assert torch.isfinite(r.nll).all() and torch.isfinite(r.PT).all()
    idx = np.array([0, 1, 31, 32, 63, 64, 4095, 4096, 12345, 32767, 32768, 50000, 65534, 65535])
    th = {"Lorenz": [10.0, 8.0 / 3, 28.0], "VanDerPol": [5.0]}[system]
    o = RC.ekf_run(system, "RKF45", 0.01, w["x0"][idx], T, t0=w["t0"], P0_sqrt=w["P0_sqrt"], theta=th, H=w["H"],
                   R_sqrt=w["R_sqrt"], ys=ys, correct_flags=flags, xy_index_map=ymap, guard="reference")
    # BOTH sides apply the zero-gain guard exactly as the reference writes it (`all(S_sqrt < 1e-16)` on a
    # factor with LAPACK's Householder signs, sqrt_ekf.py:350-353): the kernels run in factor form
    # (guard_mode reference), Oracle-B is the reference's own square-root formulation.  On Lorenz the
    # predicate never fires; on Van der Pol with H = I it fires for healthy all-negative factors (the
    # reference silently drops those observations) and the kernel must drop the SAME ones.
    fired = r.guard_fired[idx].cpu().numpy()
    assert int(fired.sum()) == o["guard_fired_steps"]
    assert int(r.guard_mismatch[idx].sum()) == o["guard_mismatch_steps"]
    if system == "Lorenz":
        assert o["guard_fired_steps"] == 0 and int(r.guard_fired.sum()) == 0
    else:
        assert o["guard_fired_steps"] > 0
    x, P, nll = r.xT[idx].cpu().numpy(), r.PT[idx].cpu().numpy(), r.nll[idx].cpu().numpy()
    np.testing.assert_allclose(x, o["xT"], rtol=1e-8, atol=1e-8 * np.abs(o["xT"]).max())
    np.testing.assert_allclose(P, o["PT"], rtol=1e-7, atol=1e-7 * np.abs(o["PT"]).max())
    np.testing.assert_allclose(nll, o["nll"], rtol=1e-7)
    # the filters have forgotten their initial condition: every trajectory sits on the same posterior
    assert float((r.xT - r.xT[0]).abs().max()) < 1e-6 * float(r.xT[0].abs().max())


def test_config3_full_size_sample_against_oracle():
    """B = 4,096 parameter sets x T = 10,000 steps of the 2-compartment reduced-1 model (n = 14, L = 2,
    tempering branch, row-parallel kernel): sampled parameter sets against Oracle-B."""
    from ode_uncertainty_b200 import Plan, _native as N, ekf_run
    from ode_uncertainty_b200 import ode as O
    dev = torch.device("cuda:0")
    B, T = 4096, 10000
    ob = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)
    plan = Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=2, disable_cov_update=True)
    th0 = ob.flat_params(ob.params)
    x0 = ob.build_initial_value(np.array([[-70.0, -70.0]]), ob.params).reshape(-1)
    xs, _ = RC.rk_run("MultiHH/reduced-1/2", "RKF45", 0.01, x0, T, theta=th0)
    rng = np.random.default_rng(621)
    ys = xs[1:][:, [0, 7]] + rng.normal(0, 0.1 ** 0.5, (T, 2))
    off, o = {}, 0
    for k in ob.params:
        off[k] = o
        o += ob.params[k].size
    rng = np.random.default_rng(7)
    theta = np.repeat(th0[None, :], B, 0)
    for k in ["g_Na", "g_K", "g_leak", "g_M", "g_L"]:
        sl = slice(off[k], off[k] + 2)
        theta[:, sl] = th0[sl] * (1 + 0.2 * rng.uniform(-1, 1, (B, 2)))
    theta[:, off["V_T"]:off["V_T"] + 2] += rng.uniform(-3, 3, (B, 2))
    H = np.zeros((2, 14)); H[0, 0] = 1; H[1, 7] = 1
    flags, ymap = np.ones(T, np.uint8), np.arange(T, dtype=np.int64)
    kw = dict(P0_sqrt=np.eye(14) * 1e-12, Q_sqrt=np.eye(14), gamma_sqrt=0.1, H=H, R_sqrt=np.eye(2) * 0.1 ** 0.5)
    x0b = np.repeat(x0[None, :], B, 0)
    r = ekf_run(plan, torch.as_tensor(x0b).to(dev), T, theta=torch.as_tensor(theta).to(dev), ys=torch.as_tensor(ys).to(dev),
                correct_flags=torch.as_tensor(flags).to(dev), xy_index_map=torch.as_tensor(ymap).to(dev), minimal=True, **kw)
    assert torch.isfinite(r.nll).all()
    idx = np.array([0, 15, 16, 2047, 4095])
    o = RC.ekf_run("MultiHH/reduced-1/2", "RKF45", 0.01, x0b[idx], T, theta=theta[idx], ys=ys, correct_flags=flags,
                   xy_index_map=ymap, disable=True, guard="intended", **kw)
    assert o["guard_mismatch_steps"] == 0
    np.testing.assert_allclose(r.nll[idx].cpu().numpy(), o["nll"], rtol=1e-7)
    np.testing.assert_allclose(r.xT[idx].cpu().numpy(), o["xT"], rtol=1e-8, atol=1e-8 * np.abs(o["xT"]).max())
    np.testing.assert_allclose(r.PT[idx].cpu().numpy(), o["PT"], rtol=1e-7, atol=1e-7 * np.abs(o["PT"]).max())
