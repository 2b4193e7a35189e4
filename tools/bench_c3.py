"""BASELINE config 3: Hodgkin-Huxley parameter estimation with process-noise
tempering - batched EKF loss (+ forward-mode gradient) over B parameter sets.

    python tools/bench_c3.py [B] [T] [--grad] [--single]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c3.py [B] [T] [--grad]
Under torchrun every rank evaluates its own B parameter sets (weak scaling, no data-path
collective); the batch objective - the sum of the log-likelihoods and of their gradients - is
formed with ONE NCCL all-reduce of [1 + p] doubles per evaluation, inside the timed
region; the time reported is the maximum over ranks.
2-compartment reduced-1 model (n=14, L=2, p=12 optimised scalars), RKF45 h=0.01,
disable_cov_update, Q_sqrt = I, gamma = 1e-2, R = 0.1, observation every step.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ode_uncertainty_b200 import Plan, _native as N, ekf_grad_run, ekf_run, runners  # noqa: E402
from ode_uncertainty_b200 import ode as O  # noqa: E402

import torch.distributed as dist  # noqa: E402
from ode_uncertainty_b200 import distributed as D  # noqa: E402

pos = [a for a in sys.argv[1:] if not a.startswith("--")]
B = int(pos[0]) if len(pos) > 0 else 4096
T = int(pos[1]) if len(pos) > 1 else 1000
want_grad = "--grad" in sys.argv
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
single = "--single" in sys.argv     # C3': single-compartment reduced-1 (n = 7, L = 1, p = 9), hodgkinhuxley9_r1.yaml
if single:
    ob = O.HodgkinHuxley(model="reduced-1")
    plan = Plan(N.ODE_HODGKIN_HUXLEY, N.SOLVER_RKF45, 0.01, ode_variant=1, disable_cov_update=True)
    nc, obs_cols, v0 = 1, [0], np.array([[-70.0]])
    opt = ["g_Na", "E_Na", "g_K", "E_K", "g_leak", "E_leak", "V_T", "g_M", "g_L"]      # params_optimized, :60-75
    flops_nll, flops_grad = 10.0e3, 0.18e6                                             # SURVEY 8(d) C3'
else:
    ob = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)
    plan = Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=2, disable_cov_update=True)
    nc, obs_cols, v0 = 2, [0, 7], np.array([[-70.0, -70.0]])
    opt = ["g_Na", "g_K", "g_leak", "V_T", "g_M", "g_L"]       # configs/params/hodgkinhuxley6_c2_r1.yaml:42-58
    flops_nll, flops_grad = 40.7e3, 0.99e6
n, L = plan.n, len(obs_cols)
th0 = ob.flat_params(ob.params)
x0 = ob.build_initial_value(v0, ob.params).reshape(-1)
xs = runners.solve_trajectory(plan, x0, T, theta_shared=th0, device=dev)
rng = np.random.default_rng(621)
ys = xs[1:][:, obs_cols] + rng.normal(0, 0.1 ** 0.5, (T, L))
names = list(ob.params)
off, o = {}, 0
for k in names:
    off[k] = o
    o += ob.params[k].size
idx = np.concatenate([np.arange(off[k], off[k] + nc) for k in opt])
rng = np.random.default_rng(7 + 1000 * rank)
theta = np.repeat(th0[None, :], B, 0)
for k in opt:
    sl = slice(off[k], off[k] + nc)
    # stay near the defaults so every parameter set integrates stably with the explicit solver
    theta[:, sl] = (th0[sl] + rng.uniform(-3, 3, (B, nc))) if k.startswith(("V_", "E_")) \
        else th0[sl] * (1 + 0.2 * rng.uniform(-1, 1, (B, nc)))
H = np.zeros((L, n))
for l, c in enumerate(obs_cols):
    H[l, c] = 1
kw = dict(P0_sqrt=np.eye(n) * 1e-12, theta=torch.from_numpy(theta).to(dev), Q_sqrt=np.eye(n), gamma_sqrt=0.1,
          H=H, R_sqrt=np.eye(L) * 0.1 ** 0.5, ys=torch.from_numpy(ys).to(dev),
          correct_flags=torch.ones(T, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T, device=dev))
x0b = torch.from_numpy(np.repeat(x0[None, :], B, 0)).to(dev)


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        tt = D.allreduce_max(torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev))
        best = min(best, float(tt.item()))
    return best, out


def loss_only():
    res = ekf_run(plan, x0b, T, want_final=False, minimal=True, **kw)
    res.total = D.allreduce_sum(res.nll.sum().reshape(1))      # batch objective over all ranks
    return res


def loss_and_grad():
    nll, g = ekf_grad_run(plan, x0b, T, idx, **kw)
    tot = D.allreduce_sum(torch.cat([nll.sum().reshape(1), g.sum(dim=0)]))   # [1 + p] doubles
    return nll, g, tot


t, r = timed(loss_only)
units = B * T * world
if rank != 0:
    sys.stdout = open(os.devnull, "w")
print(f"[{world} GPU(s), B={B} parameter sets per GPU, n={n}, L={L}, p={idx.size}]")
print(f"C3 nll only : B={B} T={T} {t*1e3:.1f} ms  {units/t/1e6:.2f} M param-set-steps/s  "
      f"{units/t*flops_nll/1e12:.3f} TFLOP/s alg ({flops_nll/1e3:.1f}k flops/unit)  finite={bool(torch.isfinite(r.nll).all())}")
if want_grad:
    t, (nll, g, tot) = timed(loss_and_grad, reps=1)
    print(f"C3 nll+grad : B={B} T={T} p={idx.size} {t*1e3:.1f} ms  {units/t/1e6:.3f} M param-set-steps/s  "
          f"{units/t*flops_grad/1e12:.3f} TFLOP/s alg ({flops_grad/1e6:.2f}M flops/unit)  finite={bool(torch.isfinite(g).all())}")
    print("nll agreement grad-kernel vs filter kernel:", float((nll - r.nll).abs().max() / r.nll.abs().max()))
if world > 1:
    dist.destroy_process_group()
