"""Multi-GPU plumbing: one process per GPU, the batch sharded by trajectory / parameter set /
particle.  The data path has NO collective - every unit is independent - so the
only communication is for the naturally global steps of the callers:

    gather_batch      all-gather of per-unit results (NLL[B], final states) -> every rank
    allreduce_sum     sum of batch log-likelihoods (and gradients) for an optimiser iteration
                      (scripts/run_parameter_estimation.py:793-794 sums over time; a batched
                      objective also sums over the batch)
    global_logsumexp  particle-weight normalisation (C4 extension; max + sum all-reduce)

`torch.distributed` carries them: NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU
tests.  Messages are tiny (<= B*8 bytes), i.e. latency-bound; they are issued once per run.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_batch(local: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenate per-rank shards (leading axis) produced with `shard_bounds` on every rank."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(total, r, ws)[1] - shard_bounds(total, r, ws)[0] for r in range(ws)]
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def allreduce_sum(t: torch.Tensor) -> torch.Tensor:
    rank, ws = world()
    out = t.clone()
    if ws > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out


def allreduce_max(t: torch.Tensor) -> torch.Tensor:
    rank, ws = world()
    out = t.clone()
    if ws > 1:
        dist.all_reduce(out, op=dist.ReduceOp.MAX)
    return out


def global_logsumexp(local_logw: torch.Tensor) -> torch.Tensor:
    """log sum_i exp(logw_i) over all ranks' particles, stable (global max first)."""
    m = allreduce_max(local_logw.max().reshape(1))
    s = allreduce_sum(torch.exp(local_logw - m).sum().reshape(1))
    return (m + torch.log(s))[0]
