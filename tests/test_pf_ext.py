"""Bootstrap extension of the particle ensemble (no reference counterpart): systematic
resampling + collectives checked against a single-process numpy statement of the same
algorithm, on CPU with 2 gloo ranks; the full filter on the GPU."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ode_uncertainty_b200 import distributed as D
from ode_uncertainty_b200 import particle_filter_ext as PX


def _systematic_numpy(w, u0):
    M = len(w)
    cdf = np.cumsum(w) / np.sum(w)
    cdf[-1] = 1.0
    u = (np.arange(M) + u0) / M
    return np.searchsorted(cdf, u, side="left")      # ancestor of each slot: first i with cdf_i >= u_j


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M = 1001
        rng = np.random.default_rng(3)
        x = rng.normal(size=(M, 3))
        w = rng.gamma(0.3, size=M)
        w[rng.integers(0, M, 50)] *= 40.0           # a few heavy particles -> many copies, zero copies
        w /= w.sum()
        lo, hi = D.shard_bounds(M, rank, world)
        out = PX.systematic_resample(torch.from_numpy(x[lo:hi]), torch.log(torch.from_numpy(w[lo:hi])), M, 0.37)
        assert out.shape[0] == hi - lo
        full = D.gather_batch(out, M)
        lse = D.global_logsumexp(torch.log(torch.from_numpy(w[lo:hi])) + 5.0)
        if rank == 0:
            q.put(dict(full=full.numpy(), ref=x[_systematic_numpy(w, 0.37)], lse=float(lse)))
    finally:
        dist.destroy_process_group()


def test_distributed_systematic_resampling_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res["full"], res["ref"])
    assert abs(res["lse"] - 5.0) < 1e-12


def test_single_rank_resampling_and_offsets():
    rng = np.random.default_rng(5)
    M = 257
    x = rng.normal(size=(M, 2))
    w = rng.random(M) ** 4
    w /= w.sum()
    for u0 in (0.0, 0.5, 0.999):
        out = PX.systematic_resample(torch.from_numpy(x), torch.log(torch.from_numpy(w)), M, u0)
        np.testing.assert_array_equal(out.numpy(), x[_systematic_numpy(w, u0)])
    assert 0.0 <= PX._u0(7, 3) < 1.0 and PX._u0(7, 3) == PX._u0(7, 3) and PX._u0(7, 3) != PX._u0(7, 4)


@pytest.mark.gpu
def test_bootstrap_filter_tracks_truth_and_weight_kernel():
    from oracle import ref_cpp as RC
    from ode_uncertainty_b200 import Plan, _native as N
    dev = torch.device("cuda:0")
    plan = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01, cov_scale=1e3)   # inflate eps so the ensemble spreads
    T, every = 400, 10
    xs, _ = RC.rk_run("Lorenz", "RKF45", 0.01, [1.0, 1.0, 1.0], T, theta=[10.0, 8.0 / 3, 28.0])
    rng = np.random.default_rng(8)
    ys = xs[every::every] + rng.normal(0, 0.1, (T // every, 3))
    # weight kernel against torch
    x = torch.tensor(rng.normal(size=(1000, 3)), device=dev)
    logw = torch.zeros(1000, dtype=torch.float64, device=dev)
    R = np.diag([0.01, 0.02, 0.03]) + 0.001
    PX.weight_update(x, logw, ys[0], np.eye(3), R)
    mvn = torch.distributions.MultivariateNormal(torch.tensor(ys[0]), covariance_matrix=torch.tensor(R))
    np.testing.assert_allclose(logw.cpu().numpy(), mvn.log_prob(x.cpu()).numpy(), rtol=1e-12)
    out = PX.bootstrap_filter(plan, 20000, T, ys, every, np.eye(3), np.eye(3) * 0.01, x0_shared=[1.2, 0.8, 1.1],
                              seed=7, device=dev)
    w = torch.exp(out["logw"])
    assert abs(float(w.sum()) - 1.0) < 1e-12
    mean = (w[:, None] * out["x"]).sum(0).cpu().numpy()
    assert np.abs(mean - xs[-1]).max() < 0.5, (mean, xs[-1])
    assert len(out["resampled"]) > 0 and math.isfinite(out["loglik"])


@pytest.mark.gpu
def test_fused_device_path_equals_the_eager_formulation():
    """The sync-free path (odeu_pf_weight_reduce / odeu_pf_normalize / odeu_pf_resample, one host read
    at the end) against the round-1 eager formulation on the same seeds: same log-likelihood, same ESS
    history, same resampling events, same posterior."""
    from oracle import ref_cpp as RC
    from ode_uncertainty_b200 import Plan, _native as N
    dev = torch.device("cuda:0")
    plan = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01, cov_scale=1e3)
    T, every, M = 300, 10, 40000
    xs, _ = RC.rk_run("Lorenz", "RKF45", 0.01, [1.0, 1.0, 1.0], T, theta=[10.0, 8.0 / 3, 28.0])
    ys = xs[every::every] + np.random.default_rng(8).normal(0, 0.1, (T // every, 3))
    kw = dict(x0_shared=[1.2, 0.8, 1.1], seed=7, device=dev)
    a = PX.bootstrap_filter(plan, M, T, ys, every, np.eye(3), np.eye(3) * 0.01, fused=True, **kw)
    b = PX.bootstrap_filter(plan, M, T, ys, every, np.eye(3), np.eye(3) * 0.01, fused=False, **kw)
    assert a["resampled"] == b["resampled"] and len(a["resampled"]) > 0
    np.testing.assert_allclose(a["ess"].numpy(), np.asarray(b["ess"]), rtol=1e-9)
    assert abs(a["loglik"] - b["loglik"]) <= 1e-9 * abs(b["loglik"])
    wa, wb = torch.exp(a["logw"]), torch.exp(b["logw"])
    ma, mb = (wa[:, None] * a["x"]).sum(0), (wb[:, None] * b["x"]).sum(0)
    assert torch.allclose(ma, mb, rtol=1e-9, atol=1e-9)
    assert abs(float(wa.sum()) - 1.0) < 1e-12
    # forced resampling at every observation exercises the gather + binary search every time
    c = PX.bootstrap_filter(plan, M, T, ys, every, np.eye(3), np.eye(3) * 0.01, fused=True, ess_frac=2.0, **kw)
    d = PX.bootstrap_filter(plan, M, T, ys, every, np.eye(3), np.eye(3) * 0.01, fused=False, ess_frac=2.0, **kw)
    assert c["resampled"] == d["resampled"] == list(range(T // every))
    assert abs(c["loglik"] - d["loglik"]) <= 1e-9 * abs(d["loglik"])
    assert torch.allclose(c["x"].mean(0), d["x"].mean(0), rtol=1e-9, atol=1e-9)
