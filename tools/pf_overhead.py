"""Per-observation cost of the fused bootstrap filter at the per-rank shard size of an 8-GPU run (M = 125,000),
on one GPU (no NCCL): wall time, device time, and a torch-profiler breakdown of the launch train."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ode_uncertainty_b200 import Plan, _native as N, pf_run, runners
from ode_uncertainty_b200.particle_filter_ext import bootstrap_filter
dev = torch.device("cuda:0")
planp = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
T, every = 1000, 10
xs = runners.solve_trajectory(planp, [1.0, 1.0, 1.0], T, device=dev)
ys = xs[every::every] + np.random.default_rng(8).normal(0.0, 0.1, xs[every::every].shape)
def run():
    return bootstrap_filter(planp, M, T, ys, every, np.eye(3), np.eye(3) * 1e-2, x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev, ess_frac=2.0)
def pred():
    return pf_run(planp, M, T, x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev)
for fn, name in ((pred, "predict"), (run, "bootstrap")):
    fn(); torch.cuda.synchronize()
    for _ in range(3):
        t0 = time.perf_counter(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); t_host = time.perf_counter() - t0; e1.record(); torch.cuda.synchronize()
        print(f"{name}: M={M} host-issue {1e3 * t_host:.2f} ms, device {e0.elapsed_time(e1):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
