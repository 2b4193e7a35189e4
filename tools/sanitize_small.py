"""Small runs of the shared-memory kernels for compute-sanitizer (racecheck / memcheck):
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
