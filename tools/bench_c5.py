"""BASELINE config 5: large-state stress - oscillator chain with D = 128
(n = 256 states), B = 1,024 trajectories, RKF45 h = 0.01, P0 = 1e-6 I, the first 16 positions
observed at every step with R = 1e-2; dense J P J^T on FP64 tensor-core MMAs.

    python tools/bench_c5.py [B] [T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5.py [B] [T]
(weak scaling: B trajectories per GPU, no data-path collective; time = max over ranks).
Algorithmic flops per trajectory-step: 77.6 M (4 n^3 for the covariance
propagation + stage tangents + measurement update)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ode_uncertainty_b200 import Plan, _native as N, ekf_dense_run  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 50
D, n, L = 128, 256, 16
import torch.distributed as dist  # noqa: E402
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
plan = Plan(N.ODE_LCAO, N.SOLVER_RKF45, 0.01, ode_variant=D)
rng = np.random.default_rng(7 + 1000 * rank)
x0 = np.concatenate([rng.normal(0, 1, (B, D)), np.zeros((B, D))], axis=1)
ys = torch.from_numpy(rng.normal(0, 1, (T, L))).to(dev)
H = np.eye(n)[:L]
kw = dict(P0_sqrt=np.eye(n) * 1e-3, H=H, R_sqrt=np.eye(L) * 0.1, ys=ys,
          correct_flags=torch.ones(T, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T, device=dev))
x0d = torch.from_numpy(x0).to(dev)
ws = torch.empty((int(N.lib().odeu_ekf_dense_workspace_bytes(plan.handle, B)) + 7) // 8, dtype=torch.float64, device=dev)
out = ekf_dense_run(plan, x0d, 2, workspace=ws, **kw)
torch.cuda.synchronize()
best = 1e9
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(); out = ekf_dense_run(plan, x0d, T, workspace=ws, **kw); e1.record(); torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    best = min(best, float(tt.item()))
units = B * T * world
if rank != 0:
    sys.stdout = open(os.devnull, "w")
print(f"C5 dense EKF [{world} GPU(s)]: B/GPU={B} n={n} T={T} L={L}  {best*1e3:.1f} ms  {best/T*1e3:.3f} ms/step  {units/best/1e3:.1f} k trajectory-steps/s  "
      f"{units/best*77.6e6/1e12:.2f} TFLOP/s alg (77.6 MFLOP/unit)  finite={bool(torch.isfinite(out.nll).all())}")
if world > 1:
    dist.destroy_process_group()
