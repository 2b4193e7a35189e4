"""BASELINE config 4, reference-parity part: perturbed-solver particle ensemble,
M particles, Lorenz-63, RKF45 h=0.01, T steps, Diagonal scale 1, particle 0 noise-free.

    python tools/bench_c4.py [M] [T]          (single GPU; shard M/G per GPU for G GPUs)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ode_uncertainty_b200 import Plan, _native as N, pf_run  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
dev = torch.device("cuda:0")
plan = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01)
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = pf_run(plan, M, T, x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) * 1e-3)
units = M * T
x = r.xT
print(f"C4 particle ensemble: M={M} T={T} {best*1e3:.1f} ms  {units/best/1e9:.2f} G particle-steps/s  "
      f"{units/best*210/1e12:.2f} TFLOP/s alg (~210 flops/unit)  finite={bool(torch.isfinite(x).all())}  "
      f"spread={x.std(0).tolist()}")
