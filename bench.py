#!/usr/bin/env python
"""Benchmark of the batched EKF-over-embedded-RK path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload = "C2"): BASELINE config 2 - batched EKF over 65,536 random initial
conditions per GPU, Lorenz-63 and Van der Pol, RKF45 h=0.01, T=10,000 steps, predict + correct +
log-likelihood at every step (H = I, R = 1e-3 I, one shared noisy observation sequence of the
true trajectory).  One bench "step" = one whole-trajectory pass over both batches
(2 x 65,536 x 10,000 trajectory-steps per GPU).  Weak scaling: every rank runs its own 65,536
trajectories; there is no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic flops per trajectory-step, SURVEY 8(d) / BASELINE.md section 3 (full-covariance
# count; FMA = 2): predict + correct
F_STEP = {"Lorenz": 1155.0, "VanDerPol": 501.0}
F_STEP_PREDICT = {"Lorenz": 777.0, "VanDerPol": 380.0}
METRIC = "ekf_rk_trajectory_steps_per_sec"
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--B", type=int, default=65536, help="trajectories per GPU (config: 65536)")
    ap.add_argument("--T", type=int, default=10000, help="steps per trajectory (config: 10000)")
    ap.add_argument("--cpu-sample-B", type=int, default=0, help="CPU baseline sample size (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--guard", default="reference", choices=["reference", "intended"],
                    help="zero-gain guard of the measurement update: 'reference' = the predicate exactly as "
                         "src/filters/sqrt_ekf.py:351 writes it (factor-form kernels, results identical to the "
                         "reference's also where its sign quirk drops observations); 'intended' = all(|S_sqrt| < 1e-16) "
                         "on the full-covariance kernels")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def workload_inputs(system: str, B: int, T: int, rank: int):
    """Synthetic inputs of SURVEY 8(d) C2 (host numpy)."""
    rng = np.random.default_rng(7 + 1000 * rank)
    if system == "Lorenz":
        n, x_true0, spread, t0 = 3, np.array([1.0, 1.0, 1.0]), 5.0, 0.0
    else:
        n, x_true0, spread, t0 = 2, np.array([2.0, 10.0]), 1.0, 10.0
    x0 = x_true0 + rng.uniform(-spread, spread, (B, n))
    P0_sqrt = np.eye(n) * (spread / 3 ** 0.5)          # std of U(-spread, spread)
    H = np.eye(n)
    R_sqrt = np.eye(n) * 1e-3 ** 0.5
    return dict(n=n, x_true0=x_true0, x0=x0, P0_sqrt=P0_sqrt, H=H, R_sqrt=R_sqrt, t0=t0)


def observations(system: str, T: int, w, plan=None, dev=None):
    """Shared observation sequence: the true trajectory (noise-free RK from x_true0) + N(0, 1e-3),
    default_rng(8).  The GPU arm integrates it with the product path (`plan` given); the CPU legs
    (`--impl reference`, cpu_baseline) with the oracle's RK, the only place they may use it."""
    if plan is not None:
        from ode_uncertainty_b200 import runners
        xs = runners.solve_trajectory(plan, w["x_true0"], T, t0=w["t0"], device=dev)
    else:
        from oracle import ref_cpp as RC
        th = {"Lorenz": [10.0, 8.0 / 3, 28.0], "VanDerPol": [5.0]}[system]
        xs, _ = RC.rk_run(system, "RKF45", 0.01, w["x_true0"], T, t0=w["t0"], theta=th)
    rng = np.random.default_rng(8)
    return xs[1:] + rng.normal(0.0, 1e-3 ** 0.5, (T, w["n"]))


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region (one long-lived
    `nvidia-smi -lms 50` process; B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.3)   # let the first samples arrive before the timed region starts
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        num = lambda v: v.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in rows if r and num(r[0])]
        mx = [float(r[1]) for r in rows if len(r) > 1 and num(r[1])]
        pw = [float(r[2]) for r in rows if len(r) > 2 and num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 7 for i in range(4)
                          if r[3 + i].lower().startswith("active")})
        # "under load" = samples drawing clearly more than idle power
        load = [s_ for s_, p_ in zip(sm, pw) if p_ > 250.0] or sm
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(rows), "samples_under_load": len(load)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(B_s: int, T_s: int, nthreads: int = 0):
    """Times the reference's CPU algorithm (Oracle-B: C++ restatement of the square-root EKF,
    all host threads) on a bounded sample of the same workload; returns (traj-steps/s, cores,
    sample description, seconds)."""
    from oracle import ref_cpp as RC
    cores = RC.num_threads() if nthreads <= 0 else nthreads
    units, secs = 0, 0.0
    for system in ("Lorenz", "VanDerPol"):
        w = workload_inputs(system, B_s, T_s, 0)
        ys = observations(system, T_s, w)
        th = {"Lorenz": [10.0, 8.0 / 3, 28.0], "VanDerPol": [5.0]}[system]
        flags, ymap = np.ones(T_s, np.uint8), np.arange(T_s, dtype=np.int64)
        t = time.perf_counter()
        RC.ekf_run(system, "RKF45", 0.01, w["x0"], T_s, t0=w["t0"], P0_sqrt=w["P0_sqrt"], theta=th,
                   H=w["H"], R_sqrt=w["R_sqrt"], ys=ys, correct_flags=flags, xy_index_map=ymap,
                   nthreads=cores, count_fragile=False)
        secs += time.perf_counter() - t
        units += B_s * T_s
    return units / secs, cores, f"Lorenz+VanDerPol, B={B_s} trajectories x T={T_s} steps each, Oracle-B sqrt-EKF", secs


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU formulation of the path (Oracle-B; the JAX
    reference cannot be installed offline: jax/jaxlib/diffrax/jaxopt are absent)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B_s = args.cpu_sample_B or 512
    T_s = 1000
    rate0, cores, sample, secs = cpu_reference_rate(B_s, T_s)        # sizing probe (also warm-up)
    target = 4.0                                                     # seconds per timed step
    B_s = int(max(64, min(65536, B_s * target / max(secs, 1e-3))))
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_rate(min(B_s, 256), 200)
    tot_units, tot_secs = 0.0, 0.0
    for _ in range(args.steps):
        rate, cores, sample, secs = cpu_reference_rate(B_s, T_s)
        tot_units += rate * secs
        tot_secs += secs
    value = tot_units / tot_secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "C2", "systems": ["Lorenz", "VanDerPol"], "solver": "RKF45", "h": 0.01,
                   "observations": "every step, H=I, R=1e-3 I, shared sequence",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _other_configs(dev, timed):
    """Bounded samples of BASELINE configs 3 and 4 for the `extras` block."""
    import torch
    from ode_uncertainty_b200 import Plan, _native as N, ekf_grad_run, ekf_run, pf_run, runners
    from ode_uncertainty_b200 import ode as O
    out = {}
    # C3: 2-compartment reduced-1 HH, n=14, L=2, p=12, B=4096 parameter sets
    ob = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)
    plan = Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=2, disable_cov_update=True)
    B3, T3 = 4096, 1000
    th0 = ob.flat_params(ob.params)
    x0 = ob.build_initial_value(np.array([[-70.0, -70.0]]), ob.params).reshape(-1)
    xs = runners.solve_trajectory(plan, x0, T3, theta_shared=th0, device=dev)
    rng = np.random.default_rng(621)
    ys = xs[1:][:, [0, 7]] + rng.normal(0, 0.1 ** 0.5, (T3, 2))
    names, off, o = list(ob.params), {}, 0
    for k in names:
        off[k] = o
        o += ob.params[k].size
    opt = ["g_Na", "g_K", "g_leak", "V_T", "g_M", "g_L"]
    idx = np.concatenate([np.arange(off[k], off[k] + 2) for k in opt])
    rng = np.random.default_rng(7)
    theta = np.repeat(th0[None, :], B3, 0)
    for k in opt:
        sl = slice(off[k], off[k] + 2)
        theta[:, sl] = th0[sl] + rng.uniform(-3, 3, (B3, 2)) if k == "V_T" else th0[sl] * (1 + 0.2 * rng.uniform(-1, 1, (B3, 2)))
    H = np.zeros((2, 14)); H[0, 0] = 1; H[1, 7] = 1
    kw = dict(P0_sqrt=np.eye(14) * 1e-12, theta=torch.from_numpy(theta).to(dev), Q_sqrt=np.eye(14), gamma_sqrt=0.1,
              H=H, R_sqrt=np.eye(2) * 0.1 ** 0.5, ys=torch.from_numpy(ys).to(dev),
              correct_flags=torch.ones(T3, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T3, device=dev))
    x0b = torch.from_numpy(np.repeat(x0[None, :], B3, 0)).to(dev)
    t_nll = timed(lambda: ekf_run(plan, x0b, T3, want_final=False, minimal=True, **kw), reps=2)
    out["c3_hh_loss"] = {"param_set_steps_per_s": B3 * T3 / t_nll, "sample": f"B={B3} x T={T3} (config: T=10000), n=14, L=2",
                         "alg_tflops": B3 * T3 / t_nll * 40.7e3 / 1e12}
    T3g = 200
    kwg = dict(kw, ys=kw["ys"][:T3g], correct_flags=kw["correct_flags"][:T3g], xy_index_map=kw["xy_index_map"][:T3g])
    t_g = timed(lambda: ekf_grad_run(plan, x0b, T3g, idx, **kwg), reps=2)
    out["c3_hh_loss_and_grad"] = {"param_set_steps_per_s": B3 * T3g / t_g, "sample": f"B={B3} x T={T3g}, p=12 forward-mode",
                                  "alg_tflops": B3 * T3g / t_g * 0.99e6 / 1e12}
    # C4: particle ensemble (reference-parity part: predict only)
    planp = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01)
    M4, T4 = 1_000_000, 1000
    t_pf = timed(lambda: pf_run(planp, M4, T4, x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev), reps=2)
    out["c4_particle_ensemble"] = {"particle_steps_per_s": M4 * T4 / t_pf, "sample": f"M={M4} x T={T4} (config: T=5000) on 1 GPU"}
    # C5: large-state oscillator chain, n = 256, dense J P J^T on FP64 tensor-core MMAs (DMMA)
    from ode_uncertainty_b200 import ekf_dense_run
    D5, n5, L5, B5, T5 = 128, 256, 16, 1024, 20
    plan5 = Plan(N.ODE_LCAO, N.SOLVER_RKF45, 0.01, ode_variant=D5)
    rng = np.random.default_rng(7)
    x05 = torch.from_numpy(np.concatenate([rng.normal(0, 1, (B5, D5)), np.zeros((B5, D5))], axis=1)).to(dev)
    ys5 = torch.from_numpy(rng.normal(0, 1, (T5, L5))).to(dev)
    ws5 = torch.empty((int(N.lib().odeu_ekf_dense_workspace_bytes(plan5.handle, B5)) + 7) // 8, dtype=torch.float64, device=dev)
    kw5 = dict(P0_sqrt=np.eye(n5) * 1e-3, H=np.eye(n5)[:L5], R_sqrt=np.eye(L5) * 0.1, ys=ys5, workspace=ws5,
               correct_flags=torch.ones(T5, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T5, device=dev))
    t5 = timed(lambda: ekf_dense_run(plan5, x05, T5, **kw5), reps=1)
    out["c5_large_state_dmma"] = {"traj_steps_per_s": B5 * T5 / t5, "ms_per_step": 1e3 * t5 / T5,
                                  "sample": f"B={B5} x T={T5} (config: T=1000), n=256, L=16",
                                  "alg_tflops": B5 * T5 / t5 * 77.6e6 / 1e12}
    del ws5
    return out


def _scale_extras(dev, world, rank):
    """The collective-bearing configs under the driver's clock at EVERY world size:
    C3 loss + gradient with the all-reduce of the batch loss/gradient (weak scaling: 4,096 parameter sets
    per GPU) and C4, the 1M-particle bootstrap filter sharded over the ranks (strong scaling) with its
    global weight normalisation and resampling exchange.  Timed on the device, max over ranks."""
    import torch
    import torch.distributed as dist
    from ode_uncertainty_b200 import Plan, _native as N, ekf_grad_run, pf_run, runners
    from ode_uncertainty_b200 import distributed as D
    from ode_uncertainty_b200 import ode as O
    from ode_uncertainty_b200.particle_filter_ext import bootstrap_filter

    def timed(fn, reps=2):
        best, out = 1e30, None
        for _ in range(reps + 1):                      # first repetition = warm-up
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            best = min(best, float(tt.item()))
        return best, out

    out = {}
    # ---- C3: loss + forward-mode gradient over 12 parameters, batch loss/gradient summed over all ranks
    ob = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)
    plan = Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=2, disable_cov_update=True)
    B3, T3 = 4096, 200
    th0 = ob.flat_params(ob.params)
    x0 = ob.build_initial_value(np.array([[-70.0, -70.0]]), ob.params).reshape(-1)
    xs = runners.solve_trajectory(plan, x0, T3, theta_shared=th0, device=dev)
    ys = xs[1:][:, [0, 7]] + np.random.default_rng(621).normal(0, 0.1 ** 0.5, (T3, 2))
    off, o = {}, 0
    for k in ob.params:
        off[k] = o
        o += ob.params[k].size
    opt = ["g_Na", "g_K", "g_leak", "V_T", "g_M", "g_L"]
    idx = np.concatenate([np.arange(off[k], off[k] + 2) for k in opt])
    rng = np.random.default_rng(7 + 1000 * rank)
    theta = np.repeat(th0[None, :], B3, 0)
    for k in opt:
        sl = slice(off[k], off[k] + 2)
        theta[:, sl] = th0[sl] + rng.uniform(-3, 3, (B3, 2)) if k == "V_T" else th0[sl] * (1 + 0.2 * rng.uniform(-1, 1, (B3, 2)))
    H = np.zeros((2, 14)); H[0, 0] = 1; H[1, 7] = 1
    kw = dict(P0_sqrt=np.eye(14) * 1e-12, theta=torch.from_numpy(theta).to(dev), Q_sqrt=np.eye(14), gamma_sqrt=0.1, H=H,
              R_sqrt=np.eye(2) * 0.1 ** 0.5, ys=torch.from_numpy(ys).to(dev),
              correct_flags=torch.ones(T3, dtype=torch.uint8, device=dev), xy_index_map=torch.arange(T3, device=dev))
    x0b = torch.from_numpy(np.repeat(x0[None, :], B3, 0)).to(dev)

    def c3():
        nll, g = ekf_grad_run(plan, x0b, T3, idx, **kw)
        msg = torch.cat([nll.sum().reshape(1), g.sum(0)])          # 1 + 12 doubles
        if world > 1:
            dist.all_reduce(msg)                                    # NCCL, same stream
        return msg

    t3, msg = timed(c3)
    out["c3_scale"] = {"param_set_steps_per_s": world * B3 * T3 / t3, "n_gpus": world, "scaling": "weak",
                       "sample": f"B={B3} parameter sets per GPU x T={T3}, p=12, all-reduce of 13 doubles per evaluation",
                       "ms": 1e3 * t3, "finite": bool(torch.isfinite(msg).all())}
    # ---- C4: 1M particles sharded over the ranks (strong scaling); bootstrap filter with a resampling
    # exchange at every observation (ess_frac = 2 forces it: with solver-error-sized noise the ESS test never fires)
    planp = Plan(N.ODE_LORENZ, N.SOLVER_RKF45, 0.01)
    M4, T4, every = 1_000_000, 1000, 10
    if M4 % world == 0:
        lo, hi = D.shard_bounds(M4, rank, world)
        tp, _ = timed(lambda: pf_run(planp, hi - lo, T4, x0_shared=[1.0, 1.0, 1.0], seed=7, particle_offset=lo, device=dev))
        xs4 = runners.solve_trajectory(planp, [1.0, 1.0, 1.0], T4, device=dev)
        ys4 = xs4[every::every] + np.random.default_rng(8).normal(0.0, 0.1, xs4[every::every].shape)
        tb, res = timed(lambda: bootstrap_filter(planp, M4, T4, ys4, every, np.eye(3), np.eye(3) * 1e-2,
                                                 x0_shared=[1.0, 1.0, 1.0], seed=7, device=dev, ess_frac=2.0))
        out["c4_scale"] = {"predict_particle_steps_per_s": M4 * T4 / tp, "bootstrap_particle_steps_per_s": M4 * T4 / tb,
                           "bootstrap_over_predict_time": tb / tp, "n_gpus": world, "scaling": "strong",
                           "sample": f"M={M4} particles over {world} GPU(s) x T={T4} (config: T=5000), weights every {every} steps, "
                                     f"{len(res['resampled'])} resampling exchanges (forced), no host read; N > 1: peer-memory path (rows and weights stay "
                                     f"with their owner in symmetric buffers and are read over NVLink, 2 device-side barriers per observation, "
                                     f"no NCCL on the data path; ODEU_PF_NCCL=1 selects the all-gather formulation)",
                           "predict_ms": 1e3 * tp, "bootstrap_ms": 1e3 * tb}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    # stdout carries exactly ONE line (the JSON record): while the bench runs, file descriptor 1
    # points at stderr so that library banners (NCCL prints its version to stdout) cannot
    # interleave with it; it is restored just before the record is printed.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from ode_uncertainty_b200 import Plan, _native as N, ekf_run, launch_count

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL's own banner/debug lines go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B, T = args.B, args.T

    plans, inp = {}, {}
    for system, ode_id in (("Lorenz", N.ODE_LORENZ), ("VanDerPol", N.ODE_VAN_DER_POL)):
        plans[system] = Plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01)
        w = workload_inputs(system, B, T, rank)
        ys = observations(system, T, w, plans[system], dev)
        w["x0_host"] = torch.from_numpy(w["x0"]).pin_memory()
        w["ys_host"] = torch.from_numpy(ys).pin_memory()
        w["x0_dev"] = w["x0_host"].to(dev)
        w["ys_dev"] = w["ys_host"].to(dev)
        w["flags"] = torch.ones(T, dtype=torch.uint8, device=dev)
        w["ymap"] = torch.arange(T, dtype=torch.int64, device=dev)
        inp[system] = w

    def one(system, x0_dev, ys_dev):
        w = inp[system]
        return ekf_run(plans[system], x0_dev, T, t0=w["t0"], P0_sqrt=w["P0_sqrt"], H=w["H"],
                       R_sqrt=w["R_sqrt"], ys=ys_dev, correct_flags=w["flags"], xy_index_map=w["ymap"],
                       guard=args.guard)

    def step_resident():
        return [one(s, inp[s]["x0_dev"], inp[s]["ys_dev"]) for s in ("Lorenz", "VanDerPol")]

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 pipe peak (roofline denominator), measured here with the library's DFMA kernel
    import ctypes as C
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)
    flops = C.c_double(0.0)
    st = torch.cuda.current_stream(dev)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(N.lib().odeu_bench_dfma(200000, 148 * 8, 256, C.c_void_p(scratch.data_ptr()),
                                        C.byref(flops), C.c_void_p(st.cuda_stream)), "odeu_bench_dfma")
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3))
    dfma_peak_tflops = best / 1e12

    def run_steps(k):
        """k steps enqueued back to back (launches are asynchronous on the stream, as a production
        caller issues them), each bracketed by its own events and preceded by the L2 flush; ONE
        synchronize at the end, so a host-side stall cannot idle the GPU inside the region."""
        evs, last = [], None
        for _ in range(k):
            flush.fill_(1.0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            r_l = one("Lorenz", inp["Lorenz"]["x0_dev"], inp["Lorenz"]["ys_dev"])
            ev[1].record()
            r_v = one("VanDerPol", inp["VanDerPol"]["x0_dev"], inp["VanDerPol"]["ys_dev"])
            ev[2].record()
            evs.append(ev)
            last = (r_l, r_v)
        torch.cuda.synchronize()
        return evs, last

    # ---- warm-up: the SAME loop as the timed region (flush kernel, events, allocator state).  The
    # clock sampler starts BEFORE it (its start-up sleeps 0.3 s; an idle gap right before the timed
    # region lets the GPU drop to idle clocks), so its samples cover warm-up + timed region.
    sampler = ClockSampler(local_rank)
    if not os.environ.get("ODEU_BENCH_NO_SAMPLER"):      # A/B knob: is a slow launch caused by the nvidia-smi poll?
        sampler.start()
    run_steps(max(args.warmup, 3))
    barrier()

    # ---- timed region (device-resident inputs): K steps, L2 flushed between steps
    launches0 = launch_count()
    step_ms, lorenz_ms, vdp_ms = [], [], []
    barrier()
    wall0 = time.perf_counter()
    evs, (r_l, r_v) = run_steps(args.steps)
    for ev in evs:
        lorenz_ms.append(ev[0].elapsed_time(ev[1]))
        vdp_ms.append(ev[1].elapsed_time(ev[2]))
        step_ms.append(ev[0].elapsed_time(ev[2]))
    barrier()
    if os.environ.get("ODEU_BENCH_TRACE") and rank == 0:
        print("per-step ms (lorenz, vdp):", [(round(a_, 2), round(b_, 2)) for a_, b_ in zip(lorenz_ms, vdp_ms)], file=sys.stderr)
    wall = time.perf_counter() - wall0
    launches = launch_count() - launches0
    clocks = sampler.stop()
    t_local = sum(step_ms) * 1e-3
    tt = torch.tensor([t_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    units_per_step = 2.0 * B * T * world
    value = units_per_step * args.steps / t_max
    finite = bool(torch.isfinite(r_l.nll).all() and torch.isfinite(r_v.nll).all())
    # what the zero-gain guard did on this rank's batch in the last timed step (reference mode only):
    # measurement updates dropped by the predicate as the reference writes it, and updates on which that
    # predicate and its intended meaning differ (all of them the reference's sign quirk)
    guard_stats = None
    if args.guard == "reference":
        guard_stats = {s_: {"guard_fired_steps": int(r_.guard_fired.sum()), "guard_mismatch_steps": int(r_.guard_mismatch.sum())}
                       for s_, r_ in (("Lorenz", r_l), ("VanDerPol", r_v))}

    # ---- e2e: host (pinned) buffers through the public API, H2D + D2H inside the timed region
    out_host = {s: (torch.empty(B, inp[s]["n"], dtype=torch.float64).pin_memory(),
                    torch.empty(B, dtype=torch.float64).pin_memory(),
                    torch.empty(B, inp[s]["n"], inp[s]["n"], dtype=torch.float64).pin_memory()) for s in inp}
    h2d = sum(inp[s]["x0_host"].numel() * 8 + inp[s]["ys_host"].numel() * 8 for s in inp)
    d2h = sum(sum(t_.numel() * 8 for t_ in out_host[s]) for s in inp)     # means, NLL and covariances

    def step_e2e():
        for s in ("Lorenz", "VanDerPol"):
            x0d = inp[s]["x0_host"].to(dev, non_blocking=True)
            ysd = inp[s]["ys_host"].to(dev, non_blocking=True)
            r = one(s, x0d, ysd)
            out_host[s][0].copy_(r.xT, non_blocking=True)
            out_host[s][1].copy_(r.nll, non_blocking=True)
            out_host[s][2].copy_(r.PT, non_blocking=True)
        torch.cuda.synchronize()

    step_e2e()
    barrier()
    e2e_t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_local = time.perf_counter() - e2e_t0
    et = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = units_per_step * args.steps / float(et.item())

    # ---- extras (rank 0, N=1): predict-only and streaming variants of the dominant kernel
    extras = {}
    if rank == 0:
        def timed(fn, reps=2):
            fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            return min(ts)
        w = inp["Lorenz"]
        t_pred = timed(lambda: ekf_run(plans["Lorenz"], w["x0_dev"], T, P0_sqrt=np.eye(3) * 1e-12))
        extras["lorenz_predict_only_traj_steps_per_s"] = B * T / t_pred
        extras["lorenz_predict_only_frac_fp64_peak"] = F_STEP_PREDICT["Lorenz"] * B * T / t_pred / 1e12 / dfma_peak_tflops
        # streaming variant: save t,x,eps,P every step (128 B per trajectory-step at n=3) for a
        # shorter horizon so the output (B * Ts * 21 * 8 B) stays bounded
        Ts = min(T, 512)
        t_str = timed(lambda: ekf_run(plans["Lorenz"], w["x0_dev"], Ts, P0_sqrt=np.eye(3) * 1e-12,
                                      save_interval=1, save_keys=("x", "eps", "P")), reps=1)
        bytes_str = B * (Ts + 1) * (3 + 3 + 9) * 8
        extras["lorenz_streaming_save_every_step"] = {"traj_steps_per_s": B * Ts / t_str,
                                                      "hbm_write_GBps": bytes_str / t_str / 1e9}
        # SURVEY 8(d) C2(2b) as specified: PER-TRAJECTORY observations y_t = x_t^(2a) + N(0, 1e-3) of each
        # trajectory's own prediction-only solution, P0_sqrt = 1e-12 I; the 15.7 GB observation stream
        # ys [T][L][B] is read once per launch (coalesced 256-byte rows per warp and component)
        if B == 65536 and T == 10000:
            try:
                pred = ekf_run(plans["Lorenz"], w["x0_dev"], T, P0_sqrt=np.eye(3) * 1e-12, save_interval=1, save_keys=("x",),
                               want_final=False)
                ys_pt = pred.traj["x"][1:]                                   # [T, B, n] view of the kernel's [T][n][B]
                del pred
                gen = torch.Generator(device=dev); gen.manual_seed(8)
                ys_pt = ys_pt + (1e-3 ** 0.5) * torch.randn(ys_pt.shape, generator=gen, dtype=torch.float64, device=dev)
                kw_pt = dict(t0=w["t0"], P0_sqrt=np.eye(3) * 1e-12, H=w["H"], R_sqrt=w["R_sqrt"], ys=ys_pt, ys_per_trajectory=True,
                             correct_flags=w["flags"], xy_index_map=w["ymap"], guard=args.guard)
                t_pt = timed(lambda: ekf_run(plans["Lorenz"], w["x0_dev"], T, **kw_pt), reps=2)
                r_pt = ekf_run(plans["Lorenz"], w["x0_dev"], T, **kw_pt)
                extras["c2_per_trajectory_obs"] = {
                    "lorenz_traj_steps_per_s": B * T / t_pt, "ms": 1e3 * t_pt,
                    "obs_stream_GB": ys_pt.numel() * 8 / 1e9, "obs_stream_GBps": ys_pt.numel() * 8 / t_pt / 1e9,
                    "frac_fp64_peak": F_STEP["Lorenz"] * B * T / t_pt / 1e12 / dfma_peak_tflops,
                    "all_finite": bool(torch.isfinite(r_pt.nll).all()),
                    "guard_fired_steps": None if r_pt.guard_fired is None else int(r_pt.guard_fired.sum()),
                    "sample": "Lorenz, B=65536 x T=10000, y_t = own prediction-only trajectory + N(0, 1e-3), P0_sqrt = 1e-12 I"}
                del ys_pt, r_pt
            except Exception as exc:
                extras["c2_per_trajectory_obs_error"] = f"{type(exc).__name__}: {exc}"
        # C1 (BASELINE config 1, the reference's own CPU-runnable case): ONE Lorenz trajectory,
        # T = 5,000 steps, prediction only, every step saved - a latency case, not a throughput case
        x1 = torch.ones(1, 3, dtype=torch.float64, device=dev)
        t_c1 = timed(lambda: ekf_run(plans["Lorenz"], x1, 5000, P0_sqrt=np.eye(3) * 1e-12, save_interval=1), reps=2)
        extras["c1_single_trajectory"] = {"ms": 1e3 * t_c1, "traj_steps_per_s": 5000 / t_c1,
                                          "sample": "B=1, T=5000, save_interval=1 (configs/ekf_trajectory_conrad_baseline/rkf45/lorenz.yaml)"}
        extras["launch_ms_min_median_max"] = {
            "lorenz": [float(np.min(lorenz_ms)), float(np.median(lorenz_ms)), float(np.max(lorenz_ms))],
            "vdp": [float(np.min(vdp_ms)), float(np.median(vdp_ms)), float(np.max(vdp_ms))]}
        extras["vdp_traj_steps_per_s"] = B * T / (np.mean(vdp_ms) * 1e-3)
        extras["lorenz_traj_steps_per_s"] = B * T / (np.mean(lorenz_ms) * 1e-3)
        # other BASELINE configs, bounded samples (not part of `value`): C3 Hodgkin-Huxley
        # parameter-estimation loss / loss+gradient, C4 particle ensemble
        if world == 1 and B == 65536 and T == 10000:
            try:
                extras.update(_other_configs(dev, timed))
            except Exception as exc:  # never lose the headline line to an extra
                extras["other_configs_error"] = f"{type(exc).__name__}: {exc}"

    # ---- collective-bearing configs at every world size (all ranks take part)
    if B == 65536 and T == 10000:
        try:
            sc = _scale_extras(dev, world, rank)
            if rank == 0:
                extras.update(sc)
        except Exception as exc:
            if rank == 0:
                extras["scale_extras_error"] = f"{type(exc).__name__}: {exc}"

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        B_s = args.cpu_sample_B or 1024
        rate, cores, sample, secs = cpu_reference_rate(B_s, 1000)
        if secs < 5.0:   # scale the sample to ~10-20 s of CPU work
            B_s = int(B_s * 12.0 / max(secs, 1e-3))
            rate, cores, sample, secs = cpu_reference_rate(B_s, 1000)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "seconds": secs}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        lor_s = float(np.mean(lorenz_ms)) * 1e-3
        achieved = F_STEP["Lorenz"] * B * T / lor_s / 1e12
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tr.get("B") == B and tr.get("T") == T:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C2", "systems": ["Lorenz", "VanDerPol"], "B_per_gpu": B, "T": T,
                       "solver": "RKF45", "h": 0.01,
                       "observations": "every step, H=I, R=1e-3 I, shared sequence",
                       "guard": args.guard,
                       "units_per_step": units_per_step, "l2": "flushed (256 MB write) between timed steps",
                       "parallelism": f"trajectory-sharded x{world}, no data-path collective"},
            "roofline": {"bound": "fp64", "kernel": "ekf_thread_sched_kernel<OdeLorenz,TabRKF45,3,3,64,6,%d>" % (1 if args.guard == "reference" else 0),
                         "achieved": achieved, "peak": dfma_peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved / dfma_peak_tflops,
                         "peak_source": "measured in this run: odeu_bench_dfma (8 DFMA chains/thread, 148x8x256 threads)",
                         "algorithmic_flops_per_unit": F_STEP["Lorenz"], "units_per_launch": B * T,
                         "avg_launch_ms": lor_s * 1e3, "traffic": traffic,
                         "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "wall_s_timed_region": wall,
            "all_finite": finite,
            "guard": {"mode": args.guard, "last_step": guard_stats},
            "extras": extras,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
