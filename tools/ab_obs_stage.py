"""A/B of the bulk-async (cp.async.bulk + mbarrier) staging of the per-trajectory observation stream against plain
coalesced loads (ODEU_NO_OBS_STAGE=1), SURVEY 8(d) C2(2b) shape.  Run once per setting (the switch is read once per
process):  python tools/ab_obs_stage.py [--system Lorenz|VanDerPol] [--guard reference|intended]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ode_uncertainty_b200 import Plan, ekf_run, _native as N  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--system", default="Lorenz")
ap.add_argument("--guard", default="reference")
ap.add_argument("--B", type=int, default=65536)
ap.add_argument("--T", type=int, default=10000)
a = ap.parse_args()
dev = torch.device("cuda:0")
n = 3 if a.system == "Lorenz" else 2
ode = N.ODE_LORENZ if a.system == "Lorenz" else N.ODE_VAN_DER_POL
plan = Plan(ode_id=ode, solver_id=N.SOLVER_RKF45, step_size=0.01)
rng = np.random.default_rng(0)
x0 = torch.tensor((np.array([1.0, 1.0, 1.0]) if n == 3 else np.array([2.0, 10.0])) + rng.uniform(-1, 1, (a.B, n)) * (5.0 if n == 3 else 1.0), device=dev)
B, T = a.B, a.T
t0 = 0.0 if n == 3 else 10.0
pred = ekf_run(plan, x0, T, t0=t0, P0_sqrt=np.eye(n) * 1e-12, save_interval=1, save_keys=("x",), want_final=False)
ys = pred.traj["x"][1:]
del pred
gen = torch.Generator(device=dev); gen.manual_seed(8)
ys = ys + (1e-3 ** 0.5) * torch.randn(ys.shape, generator=gen, dtype=torch.float64, device=dev)
flags = torch.ones(T, dtype=torch.uint8, device=dev)
ymap = torch.arange(T, dtype=torch.int64, device=dev)
kw = dict(t0=t0, P0_sqrt=np.eye(n) * 1e-12, H=np.eye(n), R_sqrt=np.eye(n) * 1e-3 ** 0.5, ys=ys, ys_per_trajectory=True,
          correct_flags=flags, xy_index_map=ymap, guard=a.guard)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ekf_run(plan, x0, T, **kw)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = ekf_run(plan, x0, T, **kw)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = min(ts) * 1e-3
print(json.dumps({"system": a.system, "guard": a.guard, "staged": os.environ.get("ODEU_NO_OBS_STAGE") is None,
                  "ms": [round(x, 3) for x in ts], "traj_steps_per_s": B * T / t, "obs_GBps": ys.numel() * 8 / t / 1e9,
                  "nll_sum": float(r.nll.sum()), "finite": bool(torch.isfinite(r.nll).all())}))
