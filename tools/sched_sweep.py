"""Dev tool: time the C2 Lorenz kernel (predict+correct and predict-only), static vs dynamic."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ode_uncertainty_b200 import Plan, _native as N, ekf_run
dev = torch.device("cuda:0"); B, T = 65536, 10000
res = {}
for system, ode_id in (("Lorenz", N.ODE_LORENZ), ("VanDerPol", N.ODE_VAN_DER_POL)):
    plan = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01)
    w = bench.workload_inputs(system, B, T, 0)
    ys = torch.from_numpy(bench.observations(system, T, w)).to(dev)
    x0 = torch.from_numpy(w["x0"]).to(dev)
    kw = dict(t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"], R_sqrt=w["R_sqrt"], ys=ys,
              correct_flags=torch.ones(T, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T, device=dev))
    for tag, k in (("obs", kw), ("pred", dict(P0_sqrt=np.eye(w["n"]) * 1e-12))):
        for dyn in (False, True):
            ts = []
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); ekf_run(plan, x0, T, dynamic=dyn, want_final=not os.environ.get('SWEEP_NOFINAL'), **k); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[(system, tag, dyn)] = B * T / (min(ts[1:]) * 1e-3) / 1e9
print(os.environ.get("ODEU_SCHED_NSEG"), os.environ.get("ODEU_SCHED_CAP"),
      {f"{a}/{b}/{'dyn' if c else 'static'}": round(v, 2) for (a, b, c), v in res.items()})
