"""Bootstrap particle filter on top of the perturbed-solver ensemble (EXTENSION).

The reference's `ParticleFilter` only predicts (no weights, no correct step, no resampling;
src/filters/particle_filter.py:24-118), so there is no reference oracle for this
module; BASELINE config 4 names its global steps.  Prediction is the reference-parity kernel
(`odeu_pf_run`); on top of it

    weights      logw += log N(y; H x, R), fused with the rank's      odeu_pf_weight_reduce (CUDA)
                 (max, sum exp, sum exp^2) triple
    normalise    logw -= logsumexp over ALL ranks, ESS and the        odeu_pf_normalize(_w): device scalars,
                 resampling decision                                  no host read
    resample     systematic, every rank resamples its own global      odeu_pf_scan_resample(_peer)
                 slots from the SAME global CDF
Between the ranks, per observation: N = 1 nothing; N > 1 over NVLink PEER MEMORY (default: triples stored into every
rank's symmetric buffer, weights pulled and ancestor rows fetched from their owners, two device-side barriers, no NCCL
on the data path) or, with ODEU_PF_NCCL=1 / without peer mapping, two NCCL all-gathers (triples; packed rows).  Both give
the same bits.  `fused=False` keeps the round-1 eager formulation (all-reduces, all-gather of partial sums, all-to-all
of the surviving particles with host-known split sizes) as the cross-check of the tests.

Particles are sharded by contiguous global slots (`distributed.shard_bounds`); the random stream
of the prediction is keyed by global slot and step, the resampling offset by (seed, event), so
results do not depend on the number of ranks (up to the rounding of the log-sum-exp reduction order).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _native as N
from . import distributed as D
from .engine import Plan, pf_run


def weight_update(x: torch.Tensor, logw: torch.Tensor, y, H, R) -> None:
    """In place: logw_m += log N(y; H x_m, R).  x [M, n] CUDA."""
    if not x.is_cuda:
        raise RuntimeError("weight_update runs on the GPU only")
    M, n = x.shape
    Hh = np.ascontiguousarray(np.asarray(H, dtype=np.float64))
    L = Hh.shape[0]
    yh = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(L))
    Rh = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(L, L))
    xk = x.to(torch.float64).t().contiguous()
    st = torch.cuda.current_stream(x.device)
    with torch.cuda.device(x.device):
        N.check(N.lib().odeu_pf_weight_update(M, n, L, C.c_void_p(xk.data_ptr()), yh.ctypes.data_as(C.c_void_p),
                                              Hh.ctypes.data_as(C.c_void_p), Rh.ctypes.data_as(C.c_void_p),
                                              C.c_void_p(logw.data_ptr()), C.c_void_p(st.cuda_stream)),
                "odeu_pf_weight_update")


def _count_below(c: torch.Tensor, M_total: int, u0: float) -> torch.Tensor:
    """Number of systematic positions u_j = (j + u0) / M, j = 0..M-1, with u_j <= c."""
    k = torch.floor(c * M_total - u0).to(torch.int64) + 1
    return torch.clamp(k, 0, M_total)


def systematic_resample(x: torch.Tensor, logw: torch.Tensor, M_total: int, u0: float) -> torch.Tensor:
    """Distributed systematic resampling.  x [M_local, n], logw [M_local] globally normalised
    (sum over all ranks of exp(logw) = 1).  Returns the new local particles [M_local, n] (global
    slots `shard_bounds(M_total, rank, world)`); weights become uniform."""
    rank, ws = D.world()
    w = torch.exp(logw)
    s_local = w.sum().reshape(1)
    if ws > 1:
        sums = [torch.empty_like(s_local) for _ in range(ws)]
        dist.all_gather(sums, s_local)
        sums = torch.cat(sums)
    else:
        sums = s_local
    total = sums.sum()
    c_lo = sums[:rank].sum() / total
    cdf = c_lo + torch.cumsum(w, 0) / total
    if rank == ws - 1:
        cdf[-1] = 1.0                       # the last global particle closes the CDF exactly
    hi = _count_below(cdf, M_total, u0)
    lo0 = _count_below(c_lo.reshape(1), M_total, u0)
    if rank == 0:
        lo0 = torch.zeros_like(lo0)         # positions u_j <= 0 (u0 = 0) belong to the first particle
    lo = torch.cat([lo0, hi[:-1]])
    counts = hi - lo                        # descendants per local particle
    first_slot = int(lo[0])
    desc = torch.repeat_interleave(torch.arange(x.shape[0], device=x.device), counts)
    x_desc = x[desc]                        # ordered by global slot, occupying [first_slot, first_slot + len)
    if ws == 1:
        return x_desc
    n_desc = x_desc.shape[0]
    send = []
    for r in range(ws):
        r_lo, r_hi = D.shard_bounds(M_total, r, ws)
        a, b = max(first_slot, r_lo), min(first_slot + n_desc, r_hi)
        send.append(max(0, b - a))
    send_t = torch.tensor(send, dtype=torch.int64, device=x.device)
    recv_t = torch.empty_like(send_t)
    dist.all_to_all_single(recv_t, send_t)
    recv = recv_t.tolist()
    out = torch.empty((sum(recv), x.shape[1]), dtype=x.dtype, device=x.device)
    dist.all_to_all_single(out, x_desc.contiguous(), output_split_sizes=recv, input_split_sizes=send)
    return out


def _u0(seed: int, event: int) -> float:
    """Resampling offset in [0, 1): deterministic in (seed, event), identical on all ranks."""
    g = np.random.Generator(np.random.Philox(key=int(seed) & (2 ** 64 - 1), counter=[event, 0, 0, 0]))
    return float(g.random())


_PEER_CACHE: Dict[tuple, dict] = {}


def _peer_buffers(dev: torch.device, M: int, n: int, ws: int) -> Optional[dict]:
    """Symmetric (peer-mapped) buffers of this rank for the peer-memory bootstrap path: triples [ws][3], packed
    rows [M][n + 1], weights [M] in ONE allocation that every rank maps over NVLink
    (torch.distributed._symmetric_memory); cached per shape.  None when the rendezvous is not available
    (the caller then uses the NCCL all-gather formulation) or when ODEU_PF_NCCL=1 asks for that path."""
    import os
    if os.environ.get("ODEU_PF_NCCL") or ws > 16:
        return None
    key = (dev.index, M, n, ws)
    if key in _PEER_CACHE:
        return _PEER_CACHE[key]
    try:
        import torch.distributed._symmetric_memory as symm_mem
        tri_d = 32                                            # doubles reserved for the triples (ws <= 10); 256 bytes
        while tri_d < 3 * ws:
            tri_d += 32
        total = tri_d + M * (n + 1) + M
        buf = symm_mem.empty((total,), dtype=torch.float64, device=dev)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
        ptrs = [int(q) for q in hdl.buffer_ptrs]
        tab = lambda off: (C.c_void_p * ws)(*[q + 8 * off for q in ptrs])
        out = {"buf": buf, "hdl": hdl, "triples": buf[:3 * ws].view(ws, 3), "pack": buf[tri_d:tri_d + M * (n + 1)].view(M, n + 1),
               "w": buf[tri_d + M * (n + 1):], "tri_tab": tab(0), "pack_tab": tab(tri_d), "w_tab": tab(tri_d + M * (n + 1))}
        torch.cuda.synchronize(dev)
        hdl.barrier(0, 30000)
    except Exception as exc:                                   # no peer mapping on this system: NCCL formulation
        import warnings
        warnings.warn(f"peer-memory bootstrap path unavailable ({type(exc).__name__}: {exc}); using NCCL all-gathers")
        out = None
    _PEER_CACHE[key] = out
    return out


def _u0_stream(seed: int, n_events: int) -> np.ndarray:
    """_u0(seed, k) for k = 0 .. n_events - 1 from ONE generator (constructing a Philox bit generator costs
    ~150 us; event k's block is the k-th block of the stream started at counter 0, bit-identical values)."""
    g = np.random.Generator(np.random.Philox(key=int(seed) & (2 ** 64 - 1), counter=[0, 0, 0, 0]))
    return g.random(4 * max(n_events, 1))[::4].copy()


def bootstrap_filter(plan: Plan, M_total: int, T: int, ys, obs_every: int, H, R, *, x0_shared,
                     t0: float = 0.0, seed: int = 7, ess_frac: float = 0.5, theta_shared=None,
                     device="cuda", fused: bool = True) -> Dict[str, torch.Tensor]:
    """Bootstrap particle filter: `obs_every` prediction steps between observations `ys[k]`.
    Returns local particles, normalised log-weights, per-observation ESS and the resample events,
    plus the log marginal likelihood estimate.

    fused (default): the sync-free device path (`_bootstrap_fused`: three CUDA kernels and two
    all-gathers per observation, no host read until the end); fused=False keeps the round-1 eager
    formulation (host-synchronised all-to-all), used as the cross-check in tests."""
    if fused:
        return _bootstrap_fused(plan, M_total, T, ys, obs_every, H, R, x0_shared=x0_shared, t0=t0, seed=seed,
                                ess_frac=ess_frac, theta_shared=theta_shared, device=device)
    rank, ws = D.world()
    lo, hi = D.shard_bounds(M_total, rank, ws)
    M = hi - lo
    dev = torch.device(device)
    n = plan.n
    x = torch.as_tensor(np.asarray(x0_shared, dtype=np.float64).reshape(1, n)).to(dev).repeat(M, 1)
    logw = torch.full((M,), -math.log(M_total), dtype=torch.float64, device=dev)
    t = float(t0)
    n_obs = T // obs_every
    ess_hist, resampled, loglik = [], [], 0.0
    for k in range(n_obs):
        r = pf_run(plan, M, obs_every, x0=x, t0=t, theta_shared=theta_shared, seed=seed,
                   particle_offset=lo, step_offset=k * obs_every)
        x, t = r.xT.contiguous(), float(r.tT)
        weight_update(x, logw, np.asarray(ys)[k], H, R)
        lse = D.global_logsumexp(logw)          # log sum_i w_i^{k-1} p(y_k | x_i): likelihood increment
        loglik += float(lse)
        logw = logw - lse
        ess = 1.0 / float(D.allreduce_sum(torch.exp(2.0 * logw).sum().reshape(1)))
        ess_hist.append(ess)
        if ess < ess_frac * M_total:
            x = systematic_resample(x, logw, M_total, _u0(seed, k))
            logw = torch.full((M,), -math.log(M_total), dtype=torch.float64, device=dev)
            resampled.append(k)
    return {"x": x, "logw": logw, "ess": torch.tensor(ess_hist, dtype=torch.float64), "resampled": resampled,
            "loglik": loglik, "t": t}


def _bootstrap_fused(plan: Plan, M_total: int, T: int, ys, obs_every: int, H, R, *, x0_shared, t0, seed, ess_frac,
                     theta_shared, device) -> Dict[str, torch.Tensor]:
    """Device-resident bootstrap filter.  Per observation, all on torch's current stream:
    odeu_pf_run (obs_every steps) -> odeu_pf_weight_reduce -> all-gather of G x 3 doubles ->
    odeu_pf_normalize -> all-gather of the packed (x, w) rows -> odeu_pf_scan_resample (CUB scan of the
    weight column + resampling): four C-ABI calls and no torch operator per observation.
    The ESS test, the resampling decision and the log-likelihood stay on the device; the host reads
    them once after the last observation.  Every rank resamples its own global slots from the SAME
    gathered CDF, so the result does not depend on the number of ranks.  Needs M_total % world == 0."""
    rank, ws = D.world()
    if M_total % ws:
        raise ValueError(f"the fused bootstrap filter shards evenly: M_total={M_total} is not a multiple of {ws} ranks")
    M = M_total // ws
    lo = rank * M
    dev = torch.device(device)
    n = plan.n
    f64 = dict(dtype=torch.float64, device=dev)
    Hh = np.ascontiguousarray(np.asarray(H, dtype=np.float64))
    L = Hh.shape[0]
    Rh = np.ascontiguousarray(np.asarray(R, dtype=np.float64).reshape(L, L))
    ysh = np.ascontiguousarray(np.asarray(ys, dtype=np.float64))
    n_obs = T // obs_every
    lib = N.lib()
    # two ensemble buffers [n][M]: the prediction reads A and writes B, the resampling reads B (and the gathered
    # rows) and writes A - also when the device-side decision is "keep" - so nothing is copied or allocated per
    # observation, and the argument block of odeu_pf_run is built once
    xa = torch.as_tensor(np.asarray(x0_shared, dtype=np.float64).reshape(n, 1)).to(dev).repeat(1, M).contiguous()
    xb = torch.empty_like(xa)
    logw = torch.full((M,), -math.log(M_total), **f64)
    triple = torch.zeros(3, **f64)
    peer = _peer_buffers(dev, M, n, ws) if ws > 1 else None
    if peer is not None:
        triples, pack, pack_all = peer["triples"], peer["pack"], peer["pack"]
    else:
        triples = torch.zeros(ws, 3, **f64) if ws > 1 else triple.reshape(1, 3)
        pack = torch.empty(M, n + 1, **f64)
        pack_all = pack if ws == 1 else torch.empty(M_total, n + 1, **f64)
    stats = torch.zeros(4, **f64)
    ess_hist = torch.zeros(max(n_obs, 1), **f64)
    flag_hist = torch.zeros(max(n_obs, 1), **f64)
    scratch = torch.zeros(int(lib.odeu_pf_reduce_scratch_bytes(M)), dtype=torch.uint8, device=dev)
    scan_bytes = int(lib.odeu_pf_scan_bytes(M_total))
    scan = torch.empty(scan_bytes, dtype=torch.uint8, device=dev)
    p = lambda t_: C.c_void_p(t_.data_ptr())
    hp = lambda a_: a_.ctypes.data_as(C.c_void_p)
    ths_h = None if theta_shared is None else np.ascontiguousarray(np.asarray(theta_shared, dtype=np.float64).reshape(plan.p))
    io = N.PfIO()
    io.M, io.T, io.seed, io.particle_offset = M, int(obs_every), int(seed), lo
    io.x0, io.xT = xa.data_ptr(), xb.data_ptr()
    if ths_h is not None:
        io.theta_shared = ths_h.ctypes.data
    io_ref = C.byref(io)
    t = float(t0)
    h = plan.step_size
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    pxa, pxb, plogw, ptriple, pscratch, ppack, ppack_all, pstats, pscan = (p(xa), p(xb), p(logw), p(triple), p(scratch), p(pack),
                                                                            p(pack_all), p(stats), p(scan))
    ptriples = p(triples)
    u0s = _u0_stream(seed, n_obs)
    tri_in = triple.reshape(1, 3)
    with torch.cuda.device(dev):
        for k in range(n_obs):
            io.t0, io.step_offset = t, k * obs_every
            N.check(lib.odeu_pf_run(plan.handle, io_ref, st), "odeu_pf_run")
            for _ in range(obs_every):
                t = t + h                               # the kernel accumulates t the same way (rksolver.py:145)
            yk = np.ascontiguousarray(ysh[k])
            N.check(lib.odeu_pf_weight_reduce(M, n, L, pxb, hp(yk), hp(Hh), hp(Rh), plogw, ptriple, pscratch, st),
                    "odeu_pf_weight_reduce")
            ess_k, flag_k = C.c_void_p(ess_hist.data_ptr() + 8 * k), C.c_void_p(flag_hist.data_ptr() + 8 * k)
            if peer is not None:
                # peer-memory path: triples stored straight into every rank's buffer, rows and weights stay with
                # their owner and are read over NVLink by whoever needs them; two device-side barriers, no NCCL
                N.check(lib.odeu_pf_publish_triple(ptriple, peer["tri_tab"], rank, ws, st), "odeu_pf_publish_triple")
                peer["hdl"].barrier(0, 30000)
                N.check(lib.odeu_pf_normalize_w(M, M_total, n, ws, ptriples, pxb, plogw, ppack, p(peer["w"]), pstats,
                                                ess_k, flag_k, float(ess_frac), st), "odeu_pf_normalize_w")
                peer["hdl"].barrier(0, 30000)
                N.check(lib.odeu_pf_scan_resample_peer(M, M_total, lo, n, ws, float(u0s[k]), pstats, peer["pack_tab"],
                                                       peer["w_tab"], pxb, pxa, plogw, pscan, scan_bytes, st),
                        "odeu_pf_scan_resample_peer")
                continue
            if ws > 1:
                dist.all_gather_into_tensor(triples, tri_in)
            N.check(lib.odeu_pf_normalize(M, M_total, n, ws, ptriples, pxb, plogw, ppack, pstats, ess_k, flag_k,
                                          float(ess_frac), st), "odeu_pf_normalize")
            if ws > 1:
                dist.all_gather_into_tensor(pack_all, pack)
            N.check(lib.odeu_pf_scan_resample(M, M_total, lo, n, float(u0s[k]), pstats, ppack_all, pxb, pxa, plogw,
                                              pscan, scan_bytes, st), "odeu_pf_scan_resample")
    xk = xa
    flags = flag_hist[:n_obs].cpu().numpy()             # the only device-to-host reads of the run
    return {"x": xk.t(), "logw": logw, "ess": ess_hist[:n_obs].cpu(), "resampled": [int(k) for k in np.nonzero(flags)[0]],
            "loglik": float(stats[3]), "t": t}
