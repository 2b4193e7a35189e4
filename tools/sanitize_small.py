"""Small runs of every kernel family for compute-sanitizer (racecheck / memcheck; the tool is closed on the
shared GPU pool since round 1g - run it on a private box):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import util as U  # noqa: E402

name = "c3_mhh_r1_rkf45_temper"
spec = dict(cases.CASES[name]); spec["T"] = 4
m = cases.materialize(spec)
plan = cases.make_plan_for(spec)
x0 = np.repeat(m["x0"].reshape(1, -1).numpy(), 20, axis=0)
kw = dict(t0=m["t0"], P0_sqrt=m["P0s"].numpy(), Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
          R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
r = U.run_ekf("gpu", plan, x0, m["T"], minimal=True, **kw)
print("rows nll", r["nll"][:2])
nll, g = U.run_grad("gpu", plan, x0[:3], m["T"], np.arange(4, 8), t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                    theta_shared=plan.default_params, Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
                    R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
print("rows grad", g[0])
from ode_uncertainty_b200 import ekf_dense_run  # noqa: E402
spec = dict(cases.DENSE_CASES["c5_lcao64_rkf45_obs"]); spec["T"] = 2
m = cases.materialize(spec)
dev = torch.device("cuda:0")
d = ekf_dense_run(cases.make_plan_for(spec), torch.as_tensor(np.repeat(m["x0"].reshape(1, -1).numpy(), 2, 0)).to(dev), m["T"],
                  P0_sqrt=m["P0s"].numpy(), H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].to(dev),
                  correct_flags=torch.as_tensor(m["flags"]).to(dev), xy_index_map=torch.as_tensor(m["ymap"]).to(dev))
torch.cuda.synchronize()
print("dense nll", d.nll.tolist())
# parameter_sensitivity path: one-step nested-dual kernel + per-parameter-set diagonal Q in the row kernel
spec = dict(cases.CASES[name]); spec["T"] = 4
m = cases.materialize(spec)
idx = np.arange(4, 8)
w, wt = U.run_sens("gpu", plan, x0[:3], idx, t0=m["t0"], theta_shared=plan.default_params)
nll, g = U.run_grad("gpu", plan, x0[:3], m["T"], idx, t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                    theta_shared=plan.default_params, gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(),
                    R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"],
                    Q_sqrt_diag=w, Q_sqrt_diag_tangent=wt)
print("sens w", w[0][:3], "grad", g[0])
# thread kernels: dynamic scheduler hand-over (workspace) and the particle ensemble on an odd step offset
from ode_uncertainty_b200 import Plan, _native as N, ekf_run, pf_run  # noqa: E402
lp = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
B, T = 148 * 4 * 32, 260
ys = torch.ones(T, 3, dtype=torch.float64, device=dev)
r = ekf_run(lp, torch.ones(B, 3, dtype=torch.float64, device=dev), T, P0_sqrt=np.eye(3) * 1e-2, H=np.eye(3), R_sqrt=np.eye(3) * 0.1,
            ys=ys, correct_flags=torch.ones(T, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T, device=dev), dynamic=True)
p = pf_run(lp, 1000, 7, x0_shared=[1.0, 1.0, 1.0], seed=3, step_offset=3, device=dev)
torch.cuda.synchronize()
print("sched nll", float(r.nll[0]), "yhatT", r.yhatT[0].tolist(), "pf", p.xT[1].tolist())
