"""torchrun, >= 2 ranks: the peer-memory bootstrap path (symmetric buffers read over NVLink, no NCCL on the data
path) against the NCCL all-gather formulation - bit-identical particles, weights, ESS history and log-likelihood -
and the time of both (M = 1e6 particles, T = 1000, an observation with FORCED resampling every 10 steps)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ode_uncertainty_b200 import Plan, _native as N, runners
from ode_uncertainty_b200.particle_filter_ext import bootstrap_filter
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
plan = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01)
M, T, every = (int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000), 1000, 10
xs = runners.solve_trajectory(plan, [1.0, 1.0, 1.0], T, device=dev)
ys = xs[every::every] + np.random.default_rng(8).normal(0.0, 0.1, xs[every::every].shape)
def run(nccl, ess_frac):
    if nccl:
        os.environ["ODEU_PF_NCCL"] = "1"
    else:
        os.environ.pop("ODEU_PF_NCCL", None)
    return bootstrap_filter(plan, M, T, ys, every, np.eye(3), np.eye(3) * 1e-2, x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev,
                            ess_frac=ess_frac)
def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps + 1):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        best = min(best, float(tt))
    return best, out
for ess_frac in (2.0, 0.5):
    tp, a = timed(lambda: run(False, ess_frac))
    tn, b = timed(lambda: run(True, ess_frac))
    same = bool(torch.equal(a["x"], b["x"]) and torch.equal(a["logw"], b["logw"]) and torch.equal(a["ess"], b["ess"])
                and a["loglik"] == b["loglik"] and a["resampled"] == b["resampled"])
    ok = torch.tensor([1.0 if same else 0.0], device=dev); dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ess_frac={ess_frac}: peer-memory {tp:.2f} ms, NCCL all-gather {tn:.2f} ms, {len(a['resampled'])} resampling events, "
              f"loglik={a['loglik']:.9f}, bit-identical on all ranks: {bool(ok.item())}", flush=True)
    assert bool(ok.item())
if os.environ.get("PF_PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize(); dist.barrier()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        run(False, 2.0); torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70), flush=True)
dist.destroy_process_group()
