"""N > 1 host path on CPU: world_size-2 gloo processes shard the batch exactly like bench.py /
the GPU launcher do, run their shard through the host-compiled kernel source (test-only), and
the product's collectives (ode_uncertainty_b200.distributed) reassemble the global result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import util as U
from ode_uncertainty_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = cases.CASES["lv_rkf45_temper_q_only"]
        m = cases.materialize(spec)
        plan = cases.make_plan_for(spec)
        B = 7                                               # ragged: shards of 4 and 3
        rng = np.random.default_rng(11)
        x0 = m["x0"].reshape(1, -1).numpy() + 0.05 * rng.normal(size=(B, m["n"]))
        theta = plan.default_params[None, :] * (1 + 0.05 * rng.uniform(-1, 1, (B, plan.p)))
        kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
                  H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                  xy_index_map=m["ymap"])
        lo, hi = D.shard_bounds(B, rank, world)
        loc = U.run_ekf("hostemu", plan, x0[lo:hi], m["T"], theta=theta[lo:hi], **kw)
        nll = D.gather_batch(torch.from_numpy(loc["nll"]), B)
        xT = D.gather_batch(torch.from_numpy(loc["xT"]), B)
        total = D.allreduce_sum(torch.from_numpy(loc["nll"]).sum().reshape(1))
        lse = D.global_logsumexp(torch.from_numpy(-loc["nll"]))
        # particle ensemble: shards keyed by global particle index
        pplan = cases.make_plan_for(cases.CASES["c1_lorenz_rkf45_predict"])
        M = 9
        plo, phi = D.shard_bounds(M, rank, world)
        pl = U.run_pf("hostemu", pplan, phi - plo, 20, x0_shared=[1.0, 1.0, 1.0], seed=7, particle_offset=plo)
        px = D.gather_batch(torch.from_numpy(pl["xT"]), M)
        if rank == 0:
            full = U.run_ekf("hostemu", plan, x0, m["T"], theta=theta, **kw)
            pfull = U.run_pf("hostemu", pplan, M, 20, x0_shared=[1.0, 1.0, 1.0], seed=7)
            q.put(dict(nll=nll.numpy(), xT=xT.numpy(), total=float(total), lse=float(lse), full_nll=full["nll"],
                       full_xT=full["xT"], px=px.numpy(), pfull=pfull["xT"]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_reassembles_global_result():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(res["nll"], res["full_nll"])
    np.testing.assert_array_equal(res["xT"], res["full_xT"])
    assert abs(res["total"] - res["full_nll"].sum()) <= 1e-12 * abs(res["full_nll"].sum())
    ref_lse = float(torch.logsumexp(torch.from_numpy(-res["full_nll"]), 0))
    assert abs(res["lse"] - ref_lse) <= 1e-12 * abs(ref_lse)
    np.testing.assert_array_equal(res["px"], res["pfull"])        # sharding-invariant random stream


def test_shard_bounds_cover_and_balance():
    for total in (1, 7, 64, 65536, 1_000_003):
        for ws in (1, 2, 3, 8):
            b = [D.shard_bounds(total, r, ws) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
