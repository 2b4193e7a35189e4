"""`ys_per_trajectory` on the CUDA path (it was only exercised through the host
emulation): every kernel family that reads observations - thread kernels (full-covariance and
factor form), row kernel (loss and loss + gradient) and the dense DMMA path - must give, for a
batch with per-trajectory observation sequences ys [T_obs, B, L], exactly what B separate
single-sequence runs give; the thread kernel is also checked against Oracle-B fed the same
per-trajectory observations."""
import numpy as np
import pytest
import torch

import cases
import util as U
from oracle import ref_cpp as RC

pytestmark = pytest.mark.gpu


def _common(m):
    return dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
                H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])


@pytest.mark.parametrize("name,guard", [("lorenz_rkf45_obs_full", "intended"), ("lorenz_rkf45_obs_full", "reference"),
                                        ("lorenz_dopri65_obs_partial", "intended"), ("vdp_rkf45_obs", "reference"),
                                        ("lv_rkf45_temper_eps_plus_q", "reference")])
def test_thread_kernels_per_trajectory_observations(name, guard):
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    rng = np.random.default_rng(5)
    B = 70                                                    # ragged: not a warp multiple
    ys = m["ys"].numpy()[:, None, :] + 0.02 * rng.normal(size=(m["ys"].shape[0], B, m["L"]))
    x0 = m["x0"].reshape(1, -1).numpy() + 0.05 * rng.normal(size=(B, m["n"]))
    kw = _common(m)
    batched = U.run_ekf("gpu", plan, x0, m["T"], ys=ys, ys_per_trajectory=True, guard=guard, **kw)
    for b in (0, 31, 32, 69):
        one = U.run_ekf("gpu", plan, x0[b:b + 1], m["T"], ys=ys[:, b], guard=guard, **kw)
        for k in ("xT", "PT", "nll", "yhatT", "ST"):
            np.testing.assert_array_equal(batched[k][b], one[k][0], err_msg=k)
    from oracle import ref_torch as R
    o = RC.ekf_run(spec["ode"], spec["solver"], m["h"], x0, m["T"], theta=R.flat_params(m["params"]).numpy(), ys=ys,
                   ys_per_trajectory=True, cov=m["cov"], scale=m["scale"], disable=m["disable"], guard=guard, **kw)
    np.testing.assert_allclose(batched["xT"], o["xT"], rtol=1e-9, atol=1e-10 * np.abs(o["xT"]).max())
    np.testing.assert_allclose(batched["nll"], o["nll"], rtol=1e-9)


def test_row_kernel_loss_and_gradient_per_trajectory_observations():
    """C3 shape (2-compartment Hodgkin-Huxley, n = 14, L = 2): odeu_ekf_run (row kernel, minimal outputs)
    and odeu_ekf_grad_run with ys [T_obs, B, L]."""
    from ode_uncertainty_b200 import ekf_grad_run, ekf_run
    dev = torch.device("cuda:0")
    spec = cases.CASES["c3_mhh_r1_rkf45_temper"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    rng = np.random.default_rng(6)
    B = 37
    ys = m["ys"].numpy()[:, None, :] + 0.3 * rng.normal(size=(m["ys"].shape[0], B, m["L"]))
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0)
    x0[:, 0] += rng.uniform(-1, 1, B)
    kw = _common(m)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
    tk = dict(kw, correct_flags=t(kw["correct_flags"], torch.uint8), xy_index_map=t(kw["xy_index_map"], torch.int64))
    r = ekf_run(plan, t(x0), m["T"], ys=t(ys), ys_per_trajectory=True, minimal=True, **tk)
    idx = np.arange(4, 10)
    nll_g, g = ekf_grad_run(plan, t(x0), m["T"], idx, ys=t(ys), ys_per_trajectory=True, theta_shared=plan.default_params, **tk)
    for b in (0, 15, 16, 36):
        one = ekf_run(plan, t(x0[b:b + 1]), m["T"], ys=t(ys[:, b]), minimal=True, **tk)
        assert torch.equal(r.nll[b], one.nll[0]) and torch.equal(r.xT[b], one.xT[0]) and torch.equal(r.PT[b], one.PT[0])
        n1, g1 = ekf_grad_run(plan, t(x0[b:b + 1]), m["T"], idx, ys=t(ys[:, b]), theta_shared=plan.default_params, **tk)
        assert torch.equal(nll_g[b], n1[0]) and torch.equal(g[b], g1[0])
    assert torch.allclose(nll_g, r.nll, rtol=1e-12)
    o = RC.ekf_run(spec["ode"], spec["solver"], m["h"], x0[:3], m["T"], theta=plan.default_params, ys=ys[:, :3],
                   ys_per_trajectory=True, disable=True, guard="intended", **kw)
    np.testing.assert_allclose(r.nll[:3].cpu().numpy(), o["nll"], rtol=1e-9)


def test_dense_path_per_trajectory_observations():
    """odeu_ekf_dense_run (n = 128, DMMA path) with ys [T_obs, B, L]."""
    from ode_uncertainty_b200 import ekf_dense_run
    dev = torch.device("cuda:0")
    spec = cases.DENSE_CASES["c5_lcao64_rkf45_obs"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    rng = np.random.default_rng(9)
    B = 5
    ys = m["ys"].numpy()[:, None, :] + 0.05 * rng.normal(size=(m["ys"].shape[0], B, m["L"]))
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0) + 0.01 * rng.normal(size=(B, m["n"]))
    t = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
    kw = dict(P0_sqrt=m["P0s"].numpy(), H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), correct_flags=t(m["flags"], torch.uint8),
              xy_index_map=t(m["ymap"], torch.int64))
    r = ekf_dense_run(plan, t(x0), m["T"], ys=t(ys), ys_per_trajectory=True, **kw)
    for b in (0, 4):
        one = ekf_dense_run(plan, t(x0[b:b + 1]), m["T"], ys=t(ys[:, b]), **kw)
        assert torch.equal(r.nll[b], one.nll[0]) and torch.equal(r.xT[b], one.xT[0]) and torch.equal(r.PT[b], one.PT[0])
    # (Oracle-B holds n <= 16 states; the single-sequence reference-code fixture pins this path: tests/test_dense.py)


@pytest.mark.parametrize("guard", ["intended", "reference"])
def test_bulk_async_staged_observation_stream_equals_plain_loads(guard):
    """At benchmark batch sizes (B % 32 == 0, dynamic scheduler) the per-trajectory observation stream is
    staged through shared memory with cp.async.bulk + mbarrier (ekf_thread.cuh, OBS_CH steps per chunk).  A
    subset of the same trajectories run alone (B = 70: static launch, plain global loads) must give the
    same bits - with observations at every step, and with gaps (every third step, stopping early) so that
    chunks hold between 0 and OBS_CH lines and segment ends fall inside a chunk."""
    from ode_uncertainty_b200 import Plan, ekf_run, _native as N
    dev = torch.device("cuda:0")
    B, T, n = 32768, 333, 3
    rng = np.random.default_rng(12)
    x0 = torch.tensor(1.0 + rng.uniform(-1, 1, (B, n)), device=dev)
    plan = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
    assert N.lib().odeu_ekf_workspace_bytes(plan.handle, B, T) > 0
    for pattern in ("all", "gaps"):
        flags = torch.ones(T, dtype=torch.uint8, device=dev)
        if pattern == "gaps":
            flags.zero_()
            flags[2:int(0.7 * T):3] = 1
        n_obs = int(flags.sum())
        ymap = torch.zeros(T, dtype=torch.int64, device=dev)
        ymap[flags.bool()] = torch.arange(n_obs, device=dev)
        gen = torch.Generator(device=dev); gen.manual_seed(3)
        ys = 1.0 + 0.3 * torch.randn((n_obs, B, n), generator=gen, dtype=torch.float64, device=dev)
        kw = dict(P0_sqrt=np.eye(n) * 0.5, H=np.eye(n), R_sqrt=np.eye(n) * 0.1, correct_flags=flags, xy_index_map=ymap, guard=guard)
        big = ekf_run(plan, x0, T, ys=ys, ys_per_trajectory=True, **kw)
        idx = torch.tensor([0, 1, 31, 32, 33, 1000, 4097, 20000, B - 33, B - 1] + list(range(5000, 5060)), device=dev)
        small = ekf_run(plan, x0[idx], T, ys=ys[:, idx].contiguous(), ys_per_trajectory=True, **kw)
        for k in ("xT", "PT", "epsT", "yhatT", "ST"):
            assert torch.equal(getattr(big, k)[idx], getattr(small, k)), (pattern, k)
        assert torch.allclose(big.nll[idx], small.nll, rtol=1e-12, atol=1e-12)     # (log-determinant flushed per time segment)
        assert torch.isfinite(big.nll).all()
