"""Shared helpers for the parity tests.

`run_ekf(backend, ...)` drives the SAME kernel source through either
  * backend="gpu":     the product path  (ode_uncertainty_b200.engine -> C ABI -> CUDA), or
  * backend="hostemu": the test-only host compilation of that source (tests/host_emu.cu),
and returns numpy arrays in the reference's orientation, so one comparison routine serves the
CPU suite and the GPU suite.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ode_uncertainty_b200 import _native as N  # noqa: E402

HOSTEMU = os.path.join(ROOT, "build", "libodeu_hostemu.so")
_emu = None


def hostemu():
    global _emu
    if _emu is None:
        src = os.path.join(ROOT, "tests", "host_emu.cu")
        deps = [src] + [os.path.join(ROOT, "ode_uncertainty_b200", "csrc", f)
                        for f in os.listdir(os.path.join(ROOT, "ode_uncertainty_b200", "csrc"))
                        if f.endswith((".cuh", ".h"))]
        stale = (not os.path.exists(HOSTEMU)
                 or os.path.getmtime(HOSTEMU) < max(os.path.getmtime(d) for d in deps))
        if stale:
            os.makedirs(os.path.dirname(HOSTEMU), exist_ok=True)
            subprocess.check_call(
                ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-DODEU_HOSTEMU",
                 "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared", src, "-o", HOSTEMU])
        _emu = C.CDLL(HOSTEMU)
        _emu.hostemu_ekf_run.argtypes = [C.POINTER(N.PlanDesc), C.POINTER(C.c_double), C.c_int,
                                         C.POINTER(N.EkfIO)]
        _emu.hostemu_pf_run.argtypes = [C.POINTER(N.PlanDesc), C.POINTER(C.c_double), C.c_int,
                                        C.POINTER(N.PfIO)]
        _emu.hostemu_grad_run.argtypes = [C.POINTER(N.PlanDesc), C.POINTER(C.c_double), C.c_int,
                                          C.POINTER(N.EkfIO), C.POINTER(N.GradIO)]
        _emu.hostemu_sens_run.argtypes = [C.POINTER(N.PlanDesc), C.POINTER(C.c_double), C.c_int,
                                          C.POINTER(N.SensIO)]
        _emu.hostemu_last_error.restype = C.c_char_p
    return _emu


def _np(a, dtype=np.float64):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=dtype))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_plan(**kw):
    """Plan creation works without a GPU (it only selects launchers)."""
    from ode_uncertainty_b200 import Plan
    return Plan(**kw)


def run_ekf(backend, plan, x0, T, *, t0=0.0, P0_sqrt=None, P0=None, theta=None, theta_shared=None,
            Q_sqrt=None, gamma_sqrt=0.0, H=None, R_sqrt=None, ys=None, ys_per_trajectory=False,
            correct_flags=None, xy_index_map=None, save_interval=0, segmented=False, dynamic=True,
            minimal=False, cov_scale_batch=None, nll_nan_to_num=False, guard="intended",
            P0_sqrt_batch=None):
    """Returns dict(xT [B,n], epsT, PT [B,n,n], nll [B], tT, traj{t,x,eps,P,y_hat,S}).
    guard != "intended" (factor form) adds PT_sqrt [B,n,n], guard_fired [B], guard_mismatch [B].
    segmented (hostemu): replay the dynamic scheduler's (block, time-segment) items sequentially."""
    x0 = _np(x0)
    B, n = x0.shape
    if backend == "gpu":
        from ode_uncertainty_b200 import ekf_run
        dev = torch.device("cuda:0")
        tt = lambda a, dt=torch.float64: None if a is None else torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
        r = ekf_run(plan, tt(x0), T, t0=t0, P0_sqrt=P0_sqrt, P0=tt(P0), theta=tt(theta),
                    theta_shared=theta_shared, Q_sqrt=Q_sqrt, gamma_sqrt=gamma_sqrt, H=H,
                    R_sqrt=R_sqrt, ys=tt(ys), ys_per_trajectory=ys_per_trajectory,
                    correct_flags=tt(correct_flags, torch.uint8),
                    xy_index_map=tt(xy_index_map, torch.int64), save_interval=save_interval,
                    dynamic=dynamic, minimal=minimal, cov_scale_batch=tt(cov_scale_batch),
                    nll_nan_to_num=nll_nan_to_num, guard=guard, P0_sqrt_batch=tt(P0_sqrt_batch))
        torch.cuda.synchronize()
        c = lambda v: None if v is None else v.cpu().numpy()
        out = dict(xT=c(r.xT), epsT=c(r.epsT), PT=c(r.PT), nll=c(r.nll),
                   tT=None if r.tT is None else float(r.tT), yhatT=c(r.yhatT), ST=c(r.ST),
                   PT_sqrt=c(r.PT_sqrt), guard_fired=c(r.guard_fired), guard_mismatch=c(r.guard_mismatch))
        out["traj"] = None if r.traj is None else {k: c(v) for k, v in r.traj.items()}
        return out
    assert backend == "hostemu"
    emu = hostemu()
    keep = []

    def K(a):
        keep.append(a)
        return a

    io = N.EkfIO()
    L = 0
    io.B, io.T, io.t0 = B, int(T), float(t0)
    io.x0 = _p(K(np.ascontiguousarray(x0.T)))
    from ode_uncertainty_b200.engine import GUARD_MODES
    io.guard_mode = GUARD_MODES[guard]
    if P0_sqrt_batch is not None:
        io.P0_sqrt_batch = _p(K(np.ascontiguousarray(_np(P0_sqrt_batch).reshape(B, n * n).T)))
    elif P0 is not None:
        io.P0 = _p(K(np.ascontiguousarray(_np(P0).reshape(B, n * n).T)))
    else:
        P0s = np.eye(n) * 1e-12 if P0_sqrt is None else _np(P0_sqrt).reshape(n, n)
        io.P0_sqrt = _p(K(P0s))
    if theta is not None:
        io.theta = _p(K(np.ascontiguousarray(_np(theta).T)))
    if theta_shared is not None:
        io.theta_shared = _p(K(_np(theta_shared)))
    if Q_sqrt is not None:
        io.Q_sqrt = _p(K(_np(Q_sqrt).reshape(n, n)))
    io.gamma_sqrt = float(gamma_sqrt)
    if H is not None and ys is not None:
        Hn = K(_np(H))
        L = Hn.shape[0]
        io.H = _p(Hn)
        io.R_sqrt = _p(K(_np(R_sqrt).reshape(L, L)))
        ysn = _np(ys)
        if ys_per_trajectory:
            ysn = np.ascontiguousarray(ysn.transpose(0, 2, 1))
        io.ys = _p(K(ysn))
        io.ys_per_trajectory = int(bool(ys_per_trajectory))
        io.correct_flags = _p(K(_np(correct_flags, np.uint8)))
        io.xy_index_map = _p(K(_np(xy_index_map, np.int64)))
    io.L = L
    io.save_interval = int(save_interval)
    if cov_scale_batch is not None:
        io.cov_scale_batch = _p(K(_np(cov_scale_batch).reshape(B)))
    io.nll_nan_to_num = int(bool(nll_nan_to_num))
    if segmented:
        wsbuf = K(np.zeros((n + n * n + 2) * B + (B + 31) // 32 + 8))
        io.workspace, io.workspace_bytes = _p(wsbuf), wsbuf.nbytes
    xT, epsT, PT = np.zeros((n, B)), np.zeros((n, B)), np.zeros((n * n, B))
    yT, ST = np.zeros((max(L, 1), B)), np.zeros((max(L * L, 1), B))
    nll, tT = np.zeros(B), np.zeros(1)
    PsT, gcount = np.zeros((n * n, B)), np.zeros((2, B), dtype=np.int64)
    if guard != "intended":
        io.PT_sqrt, io.guard_counts = _p(PsT), _p(gcount)
    if minimal:      # only nll, xT, PT requested
        io.xT, io.PT, io.nll = map(_p, (xT, PT, nll))
    else:
        io.xT, io.epsT, io.PT, io.yhatT, io.ST, io.nll, io.tT = map(_p, (xT, epsT, PT, yT, ST, nll, tT))
    tr = None
    if save_interval > 0:
        Ts = T // save_interval + 1
        tr = dict(t=np.zeros(Ts), x=np.zeros((Ts, n, B)), eps=np.zeros((Ts, n, B)),
                  P=np.zeros((Ts, n * n, B)), y_hat=np.zeros((Ts, max(L, 1), B)),
                  S=np.zeros((Ts, max(L * L, 1), B)))
        io.out_t, io.out_x, io.out_eps, io.out_P = map(_p, (tr["t"], tr["x"], tr["eps"], tr["P"]))
        if L > 0:
            io.out_yhat, io.out_S = _p(tr["y_hat"]), _p(tr["S"])
        if guard != "intended":
            tr["P_sqrt"] = np.zeros((Ts, n * n, B))
            io.out_P_sqrt = _p(tr["P_sqrt"])
    th = (C.c_double * plan.p)(*plan.default_params)
    rc = emu.hostemu_ekf_run(C.byref(plan.desc), th, plan.p, C.byref(io))
    if rc != 0:
        raise ValueError(f"hostemu: {emu.hostemu_last_error().decode()} ({rc})")
    out = dict(xT=xT.T.copy(), epsT=epsT.T.copy(), PT=PT.T.reshape(B, n, n).copy(), nll=nll,
               tT=float(tT[0]), yhatT=yT[:L].T.copy(), ST=ST[:L * L].T.reshape(B, L, L).copy())
    if guard != "intended":
        out.update(PT_sqrt=PsT.T.reshape(B, n, n).copy(), guard_fired=gcount[0].copy(), guard_mismatch=gcount[1].copy())
    out["traj"] = None
    if tr is not None:
        Ts = tr["t"].shape[0]
        out["traj"] = dict(t=tr["t"], x=tr["x"].transpose(0, 2, 1), eps=tr["eps"].transpose(0, 2, 1),
                           P=tr["P"].transpose(0, 2, 1).reshape(Ts, B, n, n),
                           y_hat=tr["y_hat"][:, :L].transpose(0, 2, 1),
                           S=tr["S"][:, :L * L].transpose(0, 2, 1).reshape(Ts, B, L, L))
        if "P_sqrt" in tr:
            out["traj"]["P_sqrt"] = tr["P_sqrt"].transpose(0, 2, 1).reshape(Ts, B, n, n)
    return out


def run_pf(backend, plan, M, T, *, x0_shared, t0=0.0, seed=7, particle_offset=0, step_offset=0,
           save_interval=0, x0=None):
    n = plan.n
    if backend == "gpu":
        from ode_uncertainty_b200 import pf_run
        x0t = None if x0 is None else torch.as_tensor(np.asarray(x0)).to("cuda:0")
        r = pf_run(plan, M, T, x0_shared=x0_shared, x0=x0t, t0=t0, seed=seed,
                   particle_offset=particle_offset, step_offset=step_offset,
                   save_interval=save_interval, device="cuda:0")
        torch.cuda.synchronize()
        out = dict(xT=r.xT.cpu().numpy(), epsT=r.epsT.cpu().numpy(), tT=float(r.tT))
        out["traj"] = None if r.traj is None else {k: v.cpu().numpy() for k, v in r.traj.items()}
        return out
    emu = hostemu()
    io = N.PfIO()
    io.M, io.T, io.t0 = int(M), int(T), float(t0)
    x0s = _np(x0_shared).reshape(n)
    io.x0_shared = _p(x0s)
    x0k = None
    if x0 is not None:
        x0k = np.ascontiguousarray(_np(x0).T)
        io.x0 = _p(x0k)
    io.seed, io.particle_offset, io.step_offset = int(seed), int(particle_offset), int(step_offset)
    io.save_interval = int(save_interval)
    xT, epsT, tT = np.zeros((n, M)), np.zeros((n, M)), np.zeros(1)
    io.xT, io.epsT, io.tT = _p(xT), _p(epsT), _p(tT)
    tr = None
    if save_interval > 0:
        Ts = T // save_interval + 1
        tr = dict(t=np.zeros(Ts), x=np.zeros((Ts, n, M)), eps=np.zeros((Ts, n, M)))
        io.out_t, io.out_x, io.out_eps = _p(tr["t"]), _p(tr["x"]), _p(tr["eps"])
    th = (C.c_double * plan.p)(*plan.default_params)
    rc = emu.hostemu_pf_run(C.byref(plan.desc), th, plan.p, C.byref(io))
    if rc != 0:
        raise ValueError(f"hostemu: {emu.hostemu_last_error().decode()} ({rc})")
    out = dict(xT=xT.T.copy(), epsT=epsT.T.copy(), tT=float(tT[0]), traj=None)
    if tr is not None:
        out["traj"] = dict(t=tr["t"], x=tr["x"].transpose(0, 2, 1), eps=tr["eps"].transpose(0, 2, 1))
    return out


def sync_times_all(T):
    """Observation at every step: flags all True, map = arange (all shipped configs)."""
    return np.ones(T, dtype=np.uint8), np.arange(T, dtype=np.int64)


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.max(np.abs(b)), floor)
    if scale == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)


def run_grad(backend, plan, x0, T, grad_idx, *, t0=0.0, P0_sqrt=None, theta=None, theta_shared=None,
             Q_sqrt=None, gamma_sqrt=0.0, H=None, R_sqrt=None, ys=None, correct_flags=None,
             xy_index_map=None, x0_tangent=None, Q_sqrt_diag=None, Q_sqrt_diag_tangent=None):
    """Returns (nll [B], grad [B, p_opt]) from the product path (gpu) or the host-compiled source.
    x0_tangent [B, p_opt, n]: d x0 / d theta_j (initial_state_parametrized).
    Q_sqrt_diag [B, n], Q_sqrt_diag_tangent [B, p_opt, n]: parameter_sensitivity weights."""
    x0 = _np(x0)
    B, n = x0.shape
    if backend == "gpu":
        from ode_uncertainty_b200 import ekf_grad_run
        dev = torch.device("cuda:0")
        tt = lambda a, dt=torch.float64: None if a is None else torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
        nll, g = ekf_grad_run(plan, tt(x0), T, grad_idx, t0=t0, P0_sqrt=P0_sqrt, theta=tt(theta),
                              theta_shared=theta_shared, Q_sqrt=Q_sqrt, gamma_sqrt=gamma_sqrt, H=H,
                              R_sqrt=R_sqrt, ys=tt(ys), correct_flags=tt(correct_flags, torch.uint8),
                              xy_index_map=tt(xy_index_map, torch.int64), x0_tangent=tt(x0_tangent),
                              Q_sqrt_diag=tt(Q_sqrt_diag), Q_sqrt_diag_tangent=tt(Q_sqrt_diag_tangent))
        torch.cuda.synchronize()
        return nll.cpu().numpy(), g.cpu().numpy()
    emu = hostemu()
    keep = []

    def K(a):
        keep.append(a)
        return a

    io = N.EkfIO()
    io.B, io.T, io.t0 = B, int(T), float(t0)
    io.x0 = _p(K(np.ascontiguousarray(x0.T)))
    io.P0_sqrt = _p(K(np.eye(n) * 1e-12 if P0_sqrt is None else _np(P0_sqrt).reshape(n, n)))
    if theta is not None:
        io.theta = _p(K(np.ascontiguousarray(_np(theta).T)))
    if theta_shared is not None:
        io.theta_shared = _p(K(_np(theta_shared)))
    if Q_sqrt is not None:
        io.Q_sqrt = _p(K(_np(Q_sqrt).reshape(n, n)))
    io.gamma_sqrt = float(gamma_sqrt)
    L = 0
    if H is not None and ys is not None:
        Hn = K(_np(H))
        L = Hn.shape[0]
        io.H = _p(Hn)
        io.R_sqrt = _p(K(_np(R_sqrt).reshape(L, L)))
        io.ys = _p(K(_np(ys)))
        io.correct_flags = _p(K(_np(correct_flags, np.uint8)))
        io.xy_index_map = _p(K(_np(xy_index_map, np.int64)))
    io.L = L
    idx = K(np.ascontiguousarray(np.asarray(grad_idx, dtype=np.int32)))
    nll, grad = np.zeros(B), np.zeros((idx.size, B))
    io.nll = _p(nll)
    g = N.GradIO()
    g.p_opt, g.idx, g.grad = int(idx.size), _p(idx), _p(grad)
    if x0_tangent is not None:
        g.x0_tangent = _p(K(np.ascontiguousarray(_np(x0_tangent).transpose(1, 2, 0))))     # [p_opt][n][B]
    if Q_sqrt_diag is not None:
        io.Q_sqrt_diag_batch = _p(K(np.ascontiguousarray(_np(Q_sqrt_diag).T)))                 # [n][B]
        if Q_sqrt_diag_tangent is not None:
            g.Q_sqrt_diag_tangent = _p(K(np.ascontiguousarray(_np(Q_sqrt_diag_tangent).transpose(1, 2, 0))))
    th = (C.c_double * plan.p)(*plan.default_params)
    rc = emu.hostemu_grad_run(C.byref(plan.desc), th, plan.p, C.byref(io), C.byref(g))
    if rc != 0:
        raise ValueError(f"hostemu: {emu.hostemu_last_error().decode()} ({rc})")
    return nll, grad.T.copy()


def run_sens(backend, plan, x0, grad_idx, *, t0=0.0, theta=None, theta_shared=None, x0_tangent=None):
    """parameter_sensitivity weights: (w [B, n], d w / d theta_j [B, p_opt, n]) from the product path
    (gpu, odeu_param_sensitivity) or the host-compiled source of the same kernel body."""
    x0 = _np(x0)
    B, n = x0.shape
    if backend == "gpu":
        from ode_uncertainty_b200.engine import param_sensitivity
        dev = torch.device("cuda:0")
        tt = lambda a: None if a is None else torch.as_tensor(np.asarray(a), dtype=torch.float64).to(dev)
        w, wt = param_sensitivity(plan, tt(x0), grad_idx, t0=t0, theta=tt(theta), theta_shared=theta_shared,
                                  x0_tangent=tt(x0_tangent))
        torch.cuda.synchronize()
        return w.cpu().numpy(), wt.cpu().numpy()
    emu = hostemu()
    keep = []

    def K(a):
        keep.append(a)
        return a

    idx = K(np.ascontiguousarray(np.asarray(grad_idx, dtype=np.int32)))
    w, wt = np.zeros((n, B)), np.zeros((idx.size, n, B))
    s = N.SensIO()
    s.B, s.t0, s.x0 = B, float(t0), _p(K(np.ascontiguousarray(x0.T)))
    if theta is not None:
        s.theta = _p(K(np.ascontiguousarray(_np(theta).T)))
    if theta_shared is not None:
        s.theta_shared = _p(K(_np(theta_shared)))
    s.p_opt, s.idx = int(idx.size), _p(idx)
    if x0_tangent is not None:
        s.x0_tangent = _p(K(np.ascontiguousarray(_np(x0_tangent).transpose(1, 2, 0))))
    s.w, s.w_tangent = _p(w), _p(wt)
    th = (C.c_double * plan.p)(*plan.default_params)
    rc = emu.hostemu_sens_run(C.byref(plan.desc), th, plan.p, C.byref(s))
    if rc != 0:
        raise ValueError(f"hostemu: {emu.hostemu_last_error().decode()} ({rc})")
    return w.T.copy(), wt.transpose(2, 0, 1).copy()
