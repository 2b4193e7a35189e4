"""GPU suite: the product path (Python host layer -> C ABI -> sm_100a kernels) against the
oracle, on the same seeded inputs as the CPU suite, plus size-independent properties at the
benchmark's full batch size."""
import numpy as np
import pytest
import torch

import cases
import util as U

pytestmark = pytest.mark.gpu

ALL = list(cases.CASES)


@pytest.mark.parametrize("name", ALL)
def test_cuda_path_matches_oracle_golden(name):
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    out = cases.run_product("gpu", spec, save_interval=1, batch=67)   # ragged: not a warp multiple
    for b in (0, 31, 32, 66):
        cases.compare(out, gold, spec, b=b)


@pytest.mark.parametrize("name", ["lorenz_rkf45_obs_full", "lv_rkf45_temper_q_only"])
def test_cuda_path_matches_live_oracle(name):
    spec = dict(cases.CASES[name])
    spec["T"] = 40
    gold = cases.run_oracle(spec)
    out = cases.run_product("gpu", spec, save_interval=1, batch=3)
    cases.compare(out, gold, spec, b=2)


@pytest.mark.parametrize("name", ["lorenz_rkf45_obs_full", "hh_r1_rkf45_temper", "c3_mhh_r1_rkf45_temper"])
def test_cuda_matches_host_compiled_source_bitwise_or_ulp(name):
    """Same source on CPU and GPU: differences can only come from libm (exp/log/sin) and FMA
    contraction, so they stay at the few-ulp level."""
    spec = cases.CASES[name]
    g = cases.run_product("gpu", spec, save_interval=1)
    h = cases.run_product("hostemu", spec, save_interval=1)
    assert U.rel_err(g["traj"]["x"], h["traj"]["x"]) < 1e-11
    assert abs(g["nll"][0] - h["nll"][0]) <= 1e-10 * max(1.0, abs(h["nll"][0]))


def test_save_interval_and_resume_on_gpu():
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    full = cases.run_product("gpu", spec, save_interval=1, batch=5)
    strided = cases.run_product("gpu", spec, save_interval=7, batch=5)
    for k in ("t", "x", "eps", "P", "y_hat", "S"):
        np.testing.assert_array_equal(strided["traj"][k], full["traj"][k][::7])
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    kw = dict(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy())
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), 5, axis=0)
    T, T1 = m["T"], 60
    a = U.run_ekf("gpu", plan, x0, T1, P0_sqrt=m["P0s"].numpy(), correct_flags=m["flags"],
                  xy_index_map=m["ymap"], **kw)
    b = U.run_ekf("gpu", plan, a["xT"], T - T1, t0=a["tT"], P0=a["PT"],
                  correct_flags=m["flags"][T1:], xy_index_map=m["ymap"][T1:], **kw)
    np.testing.assert_array_equal(b["xT"], full["xT"])
    np.testing.assert_array_equal(b["PT"], full["PT"])


def test_full_batch_properties_lorenz():
    """BASELINE config 2 batch size (65,536 random initial conditions), shortened horizon:
    (i) every trajectory equals the same trajectory run alone (batch independence / no
    cross-thread leakage), (ii) P stays symmetric positive semi-definite, (iii) permuting the
    batch permutes the outputs bit-for-bit."""
    from ode_uncertainty_b200 import Plan, ekf_run, _native as N
    dev = torch.device("cuda:0")
    B, T = 65536, 64
    rng = np.random.default_rng(7)
    x0 = torch.tensor(np.array([1., 1., 1.]) + rng.uniform(-5, 5, (B, 3)), device=dev)
    plan = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
    ys = torch.tensor(rng.normal(size=(T, 3)), device=dev)
    flags = torch.ones(T, dtype=torch.uint8, device=dev)
    ymap = torch.arange(T, device=dev)
    kw = dict(P0_sqrt=np.eye(3), H=np.eye(3), R_sqrt=np.eye(3) * 1e-3 ** 0.5, ys=ys,
              correct_flags=flags, xy_index_map=ymap)
    r = ekf_run(plan, x0, T, **kw)
    perm = torch.randperm(B, device=dev)
    rp = ekf_run(plan, x0[perm], T, **kw)
    assert torch.equal(rp.xT, r.xT[perm]) and torch.equal(rp.PT, r.PT[perm])
    assert torch.equal(rp.nll, r.nll[perm])
    idx = torch.tensor([0, 1, 31, 32, 4097, B - 1], device=dev)
    rs = ekf_run(plan, x0[idx], T, **kw)
    assert torch.equal(rs.xT, r.xT[idx]) and torch.equal(rs.nll, r.nll[idx])
    P = r.PT
    assert torch.equal(P, P.transpose(1, 2))
    assert torch.isfinite(P).all() and torch.isfinite(r.nll).all()
    assert (torch.linalg.eigvalsh(P.cpu()) > -1e-18).all()


def test_particle_zero_is_plain_rk_and_ensemble_statistics():
    """src/filters/particle_filter.py:104-105: particle 0 is noise-free, i.e. the plain RK
    trajectory (bit-equal to the EKF mean with P ignored); the others receive N(0, (scale eps)^2)
    per step -> standardised one-step increments have zero mean / unit variance."""
    from ode_uncertainty_b200 import Plan, ekf_run, pf_run, _native as N
    dev = torch.device("cuda:0")
    plan = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
    M, T = 200_000, 50
    r = pf_run(plan, M, T, x0_shared=[1., 1., 1.], seed=7, save_interval=1, device=dev)
    e = ekf_run(plan, torch.ones(1, 3, dtype=torch.float64, device=dev), T, save_interval=1)
    assert torch.equal(r.traj["x"][:, 0], e.traj["x"][:, 0])
    # one-step check: all particles share x after 1 step except for the noise
    r1 = pf_run(plan, M, 1, x0_shared=[1., 1., 1.], seed=11, device=dev)
    z = (r1.xT[1:] - r1.xT[0]) / r1.epsT[1:]
    assert abs(float(z.mean())) < 5 / (3 * M) ** 0.5
    assert abs(float(z.var()) - 1.0) < 0.02
    assert abs(float((z[:, 0] * z[:, 1]).mean())) < 5 / M ** 0.5
    # sharding invariance: particles [1000, 2000) computed alone equal the same slice of the full run
    part = pf_run(plan, 1000, T, x0_shared=[1., 1., 1.], seed=7, particle_offset=1000, device=dev)
    assert torch.equal(part.xT, r.xT[1000:2000])


def test_dynamic_scheduler_equals_static_launch_bitwise():
    """Persistent (block, time-segment) scheduling vs one static launch at the benchmark batch
    size: same arithmetic per trajectory, so the state is bit-identical; the NLL agrees to
    rounding (its log-determinant part is a pivot product flushed once per segment)."""
    from ode_uncertainty_b200 import Plan, ekf_run, _native as N
    dev = torch.device("cuda:0")
    for ode_id, n in ((N.ODE_LORENZ, 3), (N.ODE_VAN_DER_POL, 2)):
        B, T = 65536, 400
        rng = np.random.default_rng(7)
        x0 = torch.tensor(1.0 + rng.uniform(-1, 1, (B, n)), device=dev)
        plan = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01)
        assert N.lib().odeu_ekf_workspace_bytes(plan.handle, B, T) > 0
        ys = torch.tensor(1.0 + 0.1 * rng.normal(size=(T, n)), device=dev)
        kw = dict(P0_sqrt=np.eye(n), H=np.eye(n), R_sqrt=np.eye(n) * 0.1, ys=ys,
                  correct_flags=torch.ones(T, dtype=torch.uint8, device=dev),
                  xy_index_map=torch.arange(T, device=dev))
        a = ekf_run(plan, x0, T, dynamic=True, **kw)
        b = ekf_run(plan, x0, T, dynamic=False, **kw)
        for k in ("xT", "PT", "epsT", "yhatT", "ST"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        assert torch.allclose(a.nll, b.nll, rtol=1e-12, atol=1e-12)
        assert float(a.tT) == float(b.tT)
        # observations every third step that stop after 60 % of the run: the final y_hat / S come from
        # a step in an EARLIER time segment than the end of the run (only that step writes them)
        fl = torch.zeros(T, dtype=torch.uint8, device=dev)
        fl[2:int(0.6 * T):3] = 1
        kw2 = dict(kw, correct_flags=fl)
        a = ekf_run(plan, x0, T, dynamic=True, **kw2)
        b = ekf_run(plan, x0, T, dynamic=False, **kw2)
        for k in ("xT", "PT", "yhatT", "ST"):
            assert torch.equal(getattr(a, k), getattr(b, k)), k
        assert float(a.ST.abs().max()) > 0
        # prediction only as well
        a = ekf_run(plan, x0, T, dynamic=True)
        b = ekf_run(plan, x0, T, dynamic=False)
        assert torch.equal(a.xT, b.xT) and torch.equal(a.PT, b.PT)


@pytest.mark.parametrize("name", ["hh_r1_rkf45_temper", "hh_full_rkf45_small_h", "c3_mhh_r1_rkf45_temper"])
def test_row_kernel_minimal_and_full_output_runs_on_gpu(name):
    spec = cases.CASES[name]
    gold = cases.load_golden(name)
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    B = 75
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0)
    x0[1:, 0] += 0.02 * np.arange(1, B)
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
              H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
              xy_index_map=m["ymap"])
    coop = U.run_ekf("gpu", plan, x0, m["T"], minimal=True, **kw)
    thr = U.run_ekf("gpu", plan, x0, m["T"], **kw)
    np.testing.assert_allclose(coop["nll"], thr["nll"], rtol=1e-11)
    np.testing.assert_allclose(coop["xT"], thr["xT"], rtol=1e-11, atol=1e-300)
    np.testing.assert_allclose(coop["PT"], thr["PT"], rtol=1e-9, atol=1e-12 * np.abs(thr["PT"]).max())
    assert abs(coop["nll"][0] - float(gold["nll"])) <= 1e-9 * abs(float(gold["nll"]))


@pytest.mark.gpu
def test_shared_memory_kernels_are_run_to_run_deterministic():
    """compute-sanitizer is not available on the GPU pool, so a missing barrier in the row-parallel
    kernel (15+ barrier intervals per step) is hunted the poor man's way: a race makes results depend
    on warp timing, so repeated launches on a larger batch must agree BIT FOR BIT (and with the
    sequential host replay of the same source, tests/test_parity_cpu.py)."""
    import util as U
    name = "c3_mhh_r1_rkf45_temper"
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    B = 700                                   # several CTAs per SM, ragged tail
    rng = np.random.default_rng(0)
    x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), B, axis=0)
    x0[:, 0] += rng.uniform(-2, 2, B)
    x0[:, 7] += rng.uniform(-2, 2, B)
    kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
              H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
              xy_index_map=m["ymap"])
    runs = [U.run_ekf("gpu", plan, x0, m["T"], minimal=True, **kw) for _ in range(4)]
    for r in runs[1:]:
        for k in ("nll", "xT", "PT"):
            np.testing.assert_array_equal(r[k], runs[0][k])
    g = [U.run_grad("gpu", plan, x0[:64], m["T"], np.arange(4, 10), theta_shared=plan.default_params, **kw) for _ in range(3)]
    for gi in g[1:]:
        np.testing.assert_array_equal(gi[0], g[0][0])
        np.testing.assert_array_equal(gi[1], g[0][1])
