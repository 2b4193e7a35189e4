"""Dev tool: time the streaming variant (every step saved) of the Lorenz / VdP filter kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ode_uncertainty_b200 import Plan, _native as N, ekf_run
dev = torch.device("cuda:0"); B = 65536
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for system, ode_id, n, x0v in (("Lorenz", N.ODE_LORENZ, 3, [1.0, 1.0, 1.0]), ("VanDerPol", N.ODE_VAN_DER_POL, 2, [2.0, 10.0])):
    plan = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01)
    rng = np.random.default_rng(7)
    x0 = torch.from_numpy(np.asarray(x0v) + rng.uniform(-1, 1, (B, n))).to(dev)
    for Ts in (256, 512):
        for keys in (("x",), ("x", "eps", "P")):
            fn = lambda: ekf_run(plan, x0, Ts, P0_sqrt=np.eye(n) * 1e-12, save_interval=1, save_keys=keys, want_final=False)
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            t = min(ts)
            nd = sum({"x": n, "eps": n, "P": n * n}[k] for k in keys)
            out[f"{system}/T{Ts}/{'+'.join(keys)}"] = (round(B * Ts / t / 1e9, 2), round(B * (Ts + 1) * nd * 8 / t / 1e9, 1))
print({k: f"{v[0]} G steps/s, {v[1]} GB/s" for k, v in out.items()})
