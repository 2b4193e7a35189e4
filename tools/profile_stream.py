"""One launch each of the Lorenz prediction-only filter kernel with (a) no trajectory output, (b) x saved
every step, (c) x, eps, P saved every step - for ncu (streaming-variant analysis)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ode_uncertainty_b200 import Plan, _native as N, ekf_run
dev = torch.device("cuda:0"); B, Ts = 65536, 512
plan = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
x0 = torch.from_numpy(1.0 + np.random.default_rng(7).uniform(-1, 1, (B, 3))).to(dev)
for rep in range(2):
    ekf_run(plan, x0, Ts, P0_sqrt=np.eye(3) * 1e-12, want_final=False)
    ekf_run(plan, x0, Ts, P0_sqrt=np.eye(3) * 1e-12, save_interval=1, save_keys=("x",), want_final=False)
    ekf_run(plan, x0, Ts, P0_sqrt=np.eye(3) * 1e-12, save_interval=1, save_keys=("x", "eps", "P"), want_final=False)
torch.cuda.synchronize()
