"""One launch of each C2 kernel (Lorenz / Van der Pol, predict + correct) for ncu captures.

    python tools/profile_c2.py [T] [system ...] [--guard=reference|intended]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ode_uncertainty_b200 import Plan, _native as N, ekf_run  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
guard = "reference"
args = [a for a in sys.argv[2:] if not a.startswith("--")]
for a in sys.argv[2:]:
    if a.startswith("--guard="):
        guard = a.split("=", 1)[1]
systems = args or ["Lorenz", "VanDerPol"]
dev = torch.device("cuda:0")
B = 65536
for system in systems:
    ode_id = {"Lorenz": N.ODE_LORENZ, "VanDerPol": N.ODE_VAN_DER_POL}[system]
    plan = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01)
    w = bench.workload_inputs(system, B, T, 0)
    ys = torch.from_numpy(bench.observations(system, T, w)).to(dev)
    x0 = torch.from_numpy(w["x0"]).to(dev)
    flags = torch.ones(T, dtype=torch.uint8, device=dev)
    ymap = torch.arange(T, dtype=torch.int64, device=dev)
    for _ in range(2):
        r = ekf_run(plan, x0, T, t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"], R_sqrt=w["R_sqrt"],
                    ys=ys, correct_flags=flags, xy_index_map=ymap, guard=guard)
    torch.cuda.synchronize()
    print(system, "nll mean", float(r.nll.mean()), "finite", bool(torch.isfinite(r.nll).all()))
