"""Host side of the C ABI: plans and batched whole-trajectory runs on CUDA tensors.

This layer owns nothing numerical: it allocates torch CUDA buffers in the kernel's layout
(component-major, trajectory-minor), fills an `odeu_ekf_io` / `odeu_pf_io` with raw pointers,
and calls `odeu_ekf_run` / `odeu_pf_run` on torch's current stream.  It replaces the loop
owners of the reference: `unroll` (scripts/run_filter.py:166-224) and the scan inside `nll`
(scripts/run_parameter_estimation.py:771-794).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _native as N


def _host(a, shape=None) -> Optional[np.ndarray]:
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _hp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _dev(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class Plan:
    """Compiled plugin selection (ODE x RK tableau x covariance-update function)."""

    def __init__(self, ode_id: int, solver_id: int, step_size: float, *, ode_variant: int = 0,
                 num_compartments: int = 0, cov_fn_id: int = N.COV_DIAGONAL, cov_scale: float = 1.0,
                 disable_cov_update: bool = False) -> None:
        self._h = C.c_void_p()
        self.desc = N.PlanDesc(ode_id, ode_variant, num_compartments, solver_id, float(step_size),
                               cov_fn_id, float(cov_scale), int(bool(disable_cov_update)))
        N.check(N.lib().odeu_plan_create(C.byref(self.desc), C.byref(self._h)), "odeu_plan_create")
        self.n = N.lib().odeu_plan_state_dim(self._h)
        self.p = N.lib().odeu_plan_num_params(self._h)
        buf = (C.c_double * self.p)()
        N.check(N.lib().odeu_plan_default_params(self._h, buf), "odeu_plan_default_params")
        self.default_params = np.array(buf[:], dtype=np.float64)
        self.step_size = float(step_size)

    def __del__(self):
        try:
            if self._h:
                N.lib().odeu_plan_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        return self._h


@dataclass
class EkfResult:
    """Outputs in the reference's orientation (trajectory axis first)."""
    xT: torch.Tensor            # [B, n]
    epsT: torch.Tensor          # [B, n]
    PT: torch.Tensor            # [B, n, n]
    nll: torch.Tensor           # [B]
    tT: torch.Tensor            # []
    yhatT: Optional[torch.Tensor] = None   # [B, L]
    ST: Optional[torch.Tensor] = None      # [B, L, L]
    traj: Optional[Dict[str, torch.Tensor]] = None  # t [Ts], x/eps [Ts,B,n], P [Ts,B,n,n], y_hat [Ts,B,L], S [Ts,B,L,L]
    # guard="reference" only (factor form): the final factor with the reference's Householder signs, the
    # number of measurement updates on which the zero-gain guard fired, and on which the verbatim and the
    # intended predicate differ
    PT_sqrt: Optional[torch.Tensor] = None        # [B, n, n]
    guard_fired: Optional[torch.Tensor] = None    # [B] int64
    guard_mismatch: Optional[torch.Tensor] = None  # [B] int64


GUARD_MODES = {"intended": N.GUARD_INTENDED, "reference": N.GUARD_REFERENCE,
               "intended_factor": N.GUARD_INTENDED_FACTOR}


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the EKF-RK path has no CPU fallback")


def ekf_run(plan: Plan, x0: torch.Tensor, T: int, *, t0: float = 0.0, P0_sqrt=None,
            P0: Optional[torch.Tensor] = None, theta: Optional[torch.Tensor] = None,
            theta_shared=None, Q_sqrt=None, gamma_sqrt: float = 0.0, H=None, R_sqrt=None,
            ys: Optional[torch.Tensor] = None, ys_per_trajectory: bool = False,
            correct_flags: Optional[torch.Tensor] = None,
            xy_index_map: Optional[torch.Tensor] = None, save_interval: int = 0,
            save_keys=("t", "x", "eps", "P", "y_hat", "S"), want_final: bool = True,
            skip_predict: bool = False, dynamic: bool = True, minimal: bool = False,
            cov_scale_batch: Optional[torch.Tensor] = None, nll_nan_to_num: bool = False,
            stream: Optional[torch.cuda.Stream] = None, guard: str = "intended",
            P0_sqrt_batch: Optional[torch.Tensor] = None) -> EkfResult:
    """Run T EKF steps for a batch of trajectories.

    x0 [B, n] (CUDA, float64); P0 [B, n, n] per-trajectory covariance or P0_sqrt [n, n] shared
    factor (host); theta [B, p] per-trajectory parameters; ys [T_obs, L] shared or
    [T_obs, B, L] per trajectory; correct_flags [T] bool/uint8; xy_index_map [T] int64.

    guard: "intended" (zero gain iff all(|S_sqrt| < 1e-16); full-covariance kernels) or "reference"
    (the predicate exactly as src/filters/sqrt_ekf.py:351 writes it, all(S_sqrt < 1e-16), on a factor
    with LAPACK's Householder signs; factor-form kernel, n <= 4; resume with P0_sqrt_batch = a previous
    result's PT_sqrt).
    """
    _require_cuda(x0, "x0")
    dev = x0.device
    B, n = x0.shape
    if n != plan.n:
        raise ValueError(f"x0 has state dimension {n}, plan expects {plan.n}")
    f64 = dict(dtype=torch.float64, device=dev)
    x0_k = x0.to(torch.float64).t().contiguous()                       # [n][B]
    P0_k = None
    if P0 is not None:
        _require_cuda(P0, "P0")
        P0_k = P0.to(torch.float64).reshape(B, n * n).t().contiguous()  # [n*n][B]
    if guard not in GUARD_MODES:
        raise ValueError(f"guard must be one of {sorted(GUARD_MODES)}, got {guard!r}")
    factor = guard != "intended"
    P0f_k = None
    if P0_sqrt_batch is not None:
        _require_cuda(P0_sqrt_batch, "P0_sqrt_batch")
        P0f_k = P0_sqrt_batch.to(torch.float64).reshape(B, n * n).t().contiguous()  # [n*n][B]
    P0s_h = _host(P0_sqrt, (n, n)) if P0_sqrt is not None else None
    if P0_k is None and P0s_h is None and P0f_k is None:
        P0s_h = np.eye(n) * 1e-12   # default of scripts/run_filter.py:74-78
    th_k = None
    if theta is not None:
        _require_cuda(theta, "theta")
        if theta.shape != (B, plan.p):
            raise ValueError(f"theta must be [B={B}, p={plan.p}], got {tuple(theta.shape)}")
        th_k = theta.to(torch.float64).t().contiguous()
    ths_h = _host(theta_shared, (plan.p,)) if theta_shared is not None else None
    Q_h = _host(Q_sqrt, (n, n)) if Q_sqrt is not None else None

    L = 0
    H_h = R_h = None
    ys_k = flags_k = map_k = None
    if H is not None and ys is not None:
        H_h = _host(H)
        L = H_h.shape[0]
        if H_h.shape != (L, n):
            raise ValueError("Invalid measurement matrix!")  # scripts/run_filter.py:109
        R_h = _host(R_sqrt, (L, L))
        _require_cuda(ys, "ys")
        if ys_per_trajectory:
            ys_k = ys.to(torch.float64).permute(0, 2, 1).contiguous()   # [T_obs][L][B]
        else:
            ys_k = ys.to(torch.float64).reshape(-1, L).contiguous()
        flags_k = correct_flags.to(device=dev, dtype=torch.uint8).contiguous()
        map_k = xy_index_map.to(device=dev, dtype=torch.int64).contiguous()
        if flags_k.numel() < T or map_k.numel() < T:
            raise ValueError("correct_flags / xy_index_map shorter than T")

    io = N.EkfIO()
    io.B, io.T, io.t0, io.L = B, int(T), float(t0), L
    io.x0, io.P0, io.P0_sqrt = _dev(x0_k), _dev(P0_k), _hp(P0s_h)
    io.theta, io.theta_shared = _dev(th_k), _hp(ths_h)
    io.Q_sqrt, io.gamma_sqrt = _hp(Q_h), float(gamma_sqrt)
    io.H, io.R_sqrt, io.ys = _hp(H_h), _hp(R_h), _dev(ys_k)
    io.ys_per_trajectory = int(bool(ys_per_trajectory))
    io.correct_flags, io.xy_index_map = _dev(flags_k), _dev(map_k)
    io.save_interval = int(save_interval)
    io.skip_predict = int(bool(skip_predict))
    scale_k = None
    if cov_scale_batch is not None:      # calibration sweep: one covariance-update scale per trajectory
        _require_cuda(cov_scale_batch, "cov_scale_batch")
        scale_k = cov_scale_batch.to(torch.float64).reshape(B).contiguous()
    io.cov_scale_batch, io.nll_nan_to_num = _dev(scale_k), int(bool(nll_nan_to_num))
    io.guard_mode, io.P0_sqrt_batch = GUARD_MODES[guard], _dev(P0f_k)
    ws = None
    if dynamic and save_interval == 0 and not skip_predict and P0_k is None and P0f_k is None:
        wsb = int(N.lib().odeu_ekf_workspace_bytes(plan.handle, B, int(T)))
        if wsb > 0:   # large throughput runs: dynamic (block, time-segment) scheduling
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            io.workspace, io.workspace_bytes = _dev(ws), wsb

    # minimal=True requests only nll (+ xT, PT with want_final): no eps / y_hat / S / t outputs are written
    xT = epsT = PT = yT = ST = None
    nll = torch.zeros(B, **f64)
    tT = None if minimal else torch.zeros(1, **f64)
    if want_final:
        xT = torch.empty(n, B, **f64)
        epsT = None if minimal else torch.empty(n, B, **f64)
        PT = torch.empty(n * n, B, **f64)
        if L > 0 and not minimal:
            yT = torch.zeros(L, B, **f64)
            ST = torch.zeros(L * L, B, **f64)
    io.xT, io.epsT, io.PT, io.yhatT, io.ST = _dev(xT), _dev(epsT), _dev(PT), _dev(yT), _dev(ST)
    io.nll, io.tT = _dev(nll), _dev(tT)
    PsT = gcount = None
    if factor:
        if want_final:
            PsT = torch.empty(n * n, B, **f64)
        gcount = torch.zeros(2, B, dtype=torch.int64, device=dev)
        io.PT_sqrt, io.guard_counts = _dev(PsT), _dev(gcount)

    tr = {}
    if save_interval > 0:
        Ts = T // save_interval + 1
        if "t" in save_keys:
            tr["t"] = torch.empty(Ts, **f64)
        if "x" in save_keys:
            tr["x"] = torch.empty(Ts, n, B, **f64)
        if "eps" in save_keys:
            tr["eps"] = torch.empty(Ts, n, B, **f64)
        if "P" in save_keys:
            tr["P"] = torch.empty(Ts, n * n, B, **f64)
        if L > 0 and "y_hat" in save_keys:
            tr["y_hat"] = torch.empty(Ts, L, B, **f64)
        if L > 0 and "S" in save_keys:
            tr["S"] = torch.empty(Ts, L * L, B, **f64)
        if factor and "P" in save_keys:
            tr["P_sqrt"] = torch.empty(Ts, n * n, B, **f64)
            io.out_P_sqrt = _dev(tr["P_sqrt"])
        io.out_t, io.out_x, io.out_eps = _dev(tr.get("t")), _dev(tr.get("x")), _dev(tr.get("eps"))
        io.out_P, io.out_yhat, io.out_S = _dev(tr.get("P")), _dev(tr.get("y_hat")), _dev(tr.get("S"))

    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().odeu_ekf_run(plan.handle, C.byref(io), C.c_void_p(st.cuda_stream)),
                "odeu_ekf_run")

    traj = None
    if save_interval > 0:
        traj = {}
        for k, v in tr.items():
            if k == "t":
                traj[k] = v
            elif k in ("P", "P_sqrt"):
                traj[k] = v.permute(0, 2, 1).reshape(v.shape[0], B, n, n)
            elif k == "S":
                traj[k] = v.permute(0, 2, 1).reshape(v.shape[0], B, L, L)
            else:
                traj[k] = v.permute(0, 2, 1)
    return EkfResult(
        xT=None if xT is None else xT.t(), epsT=None if epsT is None else epsT.t(),
        PT=None if PT is None else PT.t().reshape(B, n, n), nll=nll,
        tT=None if tT is None else tT[0],
        yhatT=None if yT is None else yT.t(),
        ST=None if ST is None else ST.t().reshape(B, L, L), traj=traj,
        PT_sqrt=None if PsT is None else PsT.t().reshape(B, n, n),
        guard_fired=None if gcount is None else gcount[0],
        guard_mismatch=None if gcount is None else gcount[1])


def ekf_grad_run(plan: Plan, x0: torch.Tensor, T: int, grad_idx, *, t0: float = 0.0, P0_sqrt=None,
                 theta: Optional[torch.Tensor] = None, theta_shared=None, Q_sqrt=None,
                 gamma_sqrt: float = 0.0, H=None, R_sqrt=None, ys: Optional[torch.Tensor] = None,
                 ys_per_trajectory: bool = False, correct_flags: Optional[torch.Tensor] = None,
                 xy_index_map: Optional[torch.Tensor] = None, x0_tangent: Optional[torch.Tensor] = None,
                 Q_sqrt_diag: Optional[torch.Tensor] = None, Q_sqrt_diag_tangent: Optional[torch.Tensor] = None,
                 stream: Optional[torch.cuda.Stream] = None):
    """NLL [B] and d NLL / d theta_j [B, p_opt] for the flat parameter indices `grad_idx`
    (builder order) in one launch (`odeu_ekf_grad_run`).  x0_tangent [B, p_opt, n] optional.
    Q_sqrt_diag [B, n] (+ Q_sqrt_diag_tangent [B, p_opt, n]): per-parameter-set Q_sqrt = diag(w) of
    `parameter_sensitivity` (see `param_sensitivity`); replaces Q_sqrt."""
    _require_cuda(x0, "x0")
    dev = x0.device
    B, n = x0.shape
    if n != plan.n:
        raise ValueError(f"x0 has state dimension {n}, plan expects {plan.n}")
    f64 = dict(dtype=torch.float64, device=dev)
    idx = np.ascontiguousarray(np.asarray(grad_idx, dtype=np.int32))
    p_opt = idx.size
    x0_k = x0.to(torch.float64).t().contiguous()
    P0s_h = _host(P0_sqrt, (n, n)) if P0_sqrt is not None else np.eye(n) * 1e-12
    th_k = None
    if theta is not None:
        _require_cuda(theta, "theta")
        th_k = theta.to(torch.float64).t().contiguous()
    ths_h = _host(theta_shared, (plan.p,)) if theta_shared is not None else None
    Q_h = _host(Q_sqrt, (n, n)) if Q_sqrt is not None else None
    L = 0
    H_h = R_h = None
    ys_k = flags_k = map_k = None
    if H is not None and ys is not None:
        H_h = _host(H)
        L = H_h.shape[0]
        R_h = _host(R_sqrt, (L, L))
        _require_cuda(ys, "ys")
        ys_k = (ys.to(torch.float64).permute(0, 2, 1).contiguous() if ys_per_trajectory
                else ys.to(torch.float64).reshape(-1, L).contiguous())
        flags_k = correct_flags.to(device=dev, dtype=torch.uint8).contiguous()
        map_k = xy_index_map.to(device=dev, dtype=torch.int64).contiguous()
    x0t_k = None
    if x0_tangent is not None:
        x0t_k = x0_tangent.to(torch.float64).permute(1, 2, 0).contiguous()    # [p_opt][n][B]
    qd_k = qdt_k = None
    if Q_sqrt_diag is not None:
        _require_cuda(Q_sqrt_diag, "Q_sqrt_diag")
        qd_k = Q_sqrt_diag.to(torch.float64).t().contiguous()                 # [n][B]
        if Q_sqrt_diag_tangent is not None:
            qdt_k = Q_sqrt_diag_tangent.to(torch.float64).permute(1, 2, 0).contiguous()   # [p_opt][n][B]
    io = N.EkfIO()
    io.B, io.T, io.t0, io.L = B, int(T), float(t0), L
    io.x0, io.P0_sqrt = _dev(x0_k), _hp(P0s_h)
    io.theta, io.theta_shared = _dev(th_k), _hp(ths_h)
    io.Q_sqrt, io.gamma_sqrt = _hp(Q_h), float(gamma_sqrt)
    io.Q_sqrt_diag_batch = _dev(qd_k)
    io.H, io.R_sqrt, io.ys = _hp(H_h), _hp(R_h), _dev(ys_k)
    io.ys_per_trajectory = int(bool(ys_per_trajectory))
    io.correct_flags, io.xy_index_map = _dev(flags_k), _dev(map_k)
    nll = torch.zeros(B, **f64)
    grad = torch.zeros(p_opt, B, **f64)
    io.nll = _dev(nll)
    g = N.GradIO()
    g.p_opt, g.idx, g.x0_tangent, g.grad = int(p_opt), idx.ctypes.data_as(C.c_void_p), _dev(x0t_k), _dev(grad)
    g.Q_sqrt_diag_tangent = _dev(qdt_k)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().odeu_ekf_grad_run(plan.handle, C.byref(io), C.byref(g), C.c_void_p(st.cuda_stream)),
                "odeu_ekf_grad_run")
    return nll, grad.t()


def param_sensitivity(plan: Plan, x0: torch.Tensor, grad_idx, *, t0: float = 0.0,
                      theta: Optional[torch.Tensor] = None, theta_shared=None,
                      x0_tangent: Optional[torch.Tensor] = None, want_tangent: bool = True,
                      stream: Optional[torch.cuda.Stream] = None):
    """The `parameter_sensitivity` weights of nll() (scripts/run_parameter_estimation.py:750-769):
    w [B, n] with Q_sqrt = diag(w), and d w / d theta_j [B, p_opt, n] (physical parameters), from one
    solver step at (t0, x0) (`odeu_param_sensitivity`)."""
    _require_cuda(x0, "x0")
    dev = x0.device
    B, n = x0.shape
    if n != plan.n:
        raise ValueError(f"x0 has state dimension {n}, plan expects {plan.n}")
    f64 = dict(dtype=torch.float64, device=dev)
    idx = np.ascontiguousarray(np.asarray(grad_idx, dtype=np.int32))
    p_opt = idx.size
    x0_k = x0.to(torch.float64).t().contiguous()
    th_k = None
    if theta is not None:
        _require_cuda(theta, "theta")
        th_k = theta.to(torch.float64).t().contiguous()
    ths_h = _host(theta_shared, (plan.p,)) if theta_shared is not None else None
    x0t_k = None if x0_tangent is None else x0_tangent.to(torch.float64).permute(1, 2, 0).contiguous()
    w = torch.zeros(n, B, **f64)
    wt = torch.zeros(p_opt, n, B, **f64) if want_tangent else None
    s = N.SensIO()
    s.B, s.t0, s.x0, s.theta, s.theta_shared = B, float(t0), _dev(x0_k), _dev(th_k), _hp(ths_h)
    s.p_opt, s.idx, s.x0_tangent = int(p_opt), idx.ctypes.data_as(C.c_void_p), _dev(x0t_k)
    s.w, s.w_tangent = _dev(w), _dev(wt)
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().odeu_param_sensitivity(plan.handle, C.byref(s), C.c_void_p(st.cuda_stream)),
                "odeu_param_sensitivity")
    return w.t(), (None if wt is None else wt.permute(2, 0, 1))


@dataclass
class DenseResult:
    """Large-state run (odeu_ekf_dense_run): per-trajectory row-major matrices."""
    xT: torch.Tensor            # [B, n]
    PT: torch.Tensor            # [B, n, n]
    epsT: torch.Tensor          # [B, n]
    nll: torch.Tensor           # [B]
    tT: float


def ekf_dense_run(plan: Plan, x0: torch.Tensor, T: int, *, t0: float = 0.0, P0_sqrt=None,
                  P: Optional[torch.Tensor] = None, theta_shared=None, Q_sqrt=None, gamma_sqrt: float = 0.0,
                  H=None, R_sqrt=None, ys: Optional[torch.Tensor] = None, ys_per_trajectory: bool = False,
                  correct_flags: Optional[torch.Tensor] = None, xy_index_map: Optional[torch.Tensor] = None,
                  workspace: Optional[torch.Tensor] = None, inplace: bool = False) -> DenseResult:
    """T EKF steps of the large-state oscillator chain (BASELINE config 5) on the DMMA path.

    x0 [B, n] CUDA float64.  The initial covariance is either the shared factor P0_sqrt [n, n]
    (host; default 1e-12 I like scripts/run_filter.py:74-78) or a per-trajectory P [B, n, n]
    (CUDA; updated IN PLACE when `inplace`, so a run can be resumed without copies).
    H [L, n] must select state components (rows of the identity); Q_sqrt must be diagonal."""
    _require_cuda(x0, "x0")
    dev = x0.device
    B, n = x0.shape
    if n != plan.n:
        raise ValueError(f"x0 has state dimension {n}, plan expects {plan.n}")
    x = x0.to(torch.float64).contiguous() if inplace else x0.to(torch.float64).clone().contiguous()
    P0s_h = None
    if P is None:
        P0s_h = _host(P0_sqrt, (n, n)) if P0_sqrt is not None else np.eye(n) * 1e-12
        Pd = torch.empty((B, n, n), dtype=torch.float64, device=dev)
    else:
        _require_cuda(P, "P")
        Pd = P if (inplace and P.dtype == torch.float64 and P.is_contiguous()) else P.to(torch.float64).clone().contiguous()
    ths_h = _host(theta_shared, (plan.p,)) if theta_shared is not None else None
    qd_h = None
    if Q_sqrt is not None:
        Q_h = _host(Q_sqrt, (n, n))
        if np.any(Q_h - np.diag(np.diag(Q_h)) != 0.0):
            raise ValueError("the large-state path takes a diagonal Q_sqrt")
        qd_h = np.ascontiguousarray(np.diag(Q_h))
    L = 0
    idx_h = R_h = ys_k = flags_k = map_k = None
    if H is not None and ys is not None:
        H_h = _host(H)
        L = H_h.shape[0]
        if H_h.shape != (L, n):
            raise ValueError("Invalid measurement matrix!")  # scripts/run_filter.py:109
        idx = H_h.argmax(axis=1)
        if not np.array_equal(H_h, np.eye(n)[idx]):
            raise ValueError("the large-state path takes a measurement matrix that selects state components")
        idx_h = np.ascontiguousarray(idx.astype(np.int32))
        R_h = _host(R_sqrt, (L, L))
        _require_cuda(ys, "ys")
        ys_k = ys.to(torch.float64).contiguous()          # [T_obs, L] or [T_obs, B, L]
        flags_k = correct_flags.to(device=dev, dtype=torch.uint8).contiguous()
        map_k = xy_index_map.to(device=dev, dtype=torch.int64).contiguous()
        if flags_k.numel() < T or map_k.numel() < T:
            raise ValueError("correct_flags / xy_index_map shorter than T")
    need = int(N.lib().odeu_ekf_dense_workspace_bytes(plan.handle, B))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need + 7) // 8, dtype=torch.float64, device=dev)
    eps = torch.zeros((B, n), dtype=torch.float64, device=dev)
    nll = torch.zeros(B, dtype=torch.float64, device=dev)
    tT = C.c_double(t0)
    io = N.DenseIO()
    io.B, io.T, io.t0, io.L = B, int(T), float(t0), L
    io.x0, io.x, io.P, io.P0_sqrt = _dev(x), _dev(x), _dev(Pd), _hp(P0s_h)
    io.theta_shared, io.Q_sqrt_diag, io.gamma_sqrt = _hp(ths_h), _hp(qd_h), float(gamma_sqrt)
    io.obs_index = None if idx_h is None else idx_h.ctypes.data_as(C.c_void_p)
    io.R_sqrt, io.ys, io.ys_per_trajectory = _hp(R_h), _dev(ys_k), int(bool(ys_per_trajectory))
    io.correct_flags, io.xy_index_map = _dev(flags_k), _dev(map_k)
    io.eps, io.nll, io.tT = _dev(eps), _dev(nll), C.cast(C.pointer(tT), C.c_void_p)
    io.workspace, io.workspace_bytes = _dev(workspace), workspace.numel() * workspace.element_size()
    stream = torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().odeu_ekf_dense_run(plan.handle, C.byref(io), C.c_void_p(stream.cuda_stream)),
                "odeu_ekf_dense_run")
    return DenseResult(xT=x, PT=Pd, epsT=eps, nll=nll, tT=float(tT.value))


@dataclass
class PfResult:
    xT: torch.Tensor       # [M, n]
    epsT: torch.Tensor     # [M, n]
    tT: torch.Tensor
    traj: Optional[Dict[str, torch.Tensor]] = None  # t [Ts], x/eps [Ts, M, n]


def pf_run(plan: Plan, M: int, T: int, *, x0_shared=None, x0: Optional[torch.Tensor] = None,
           t0: float = 0.0, theta_shared=None, seed: int = 7, particle_offset: int = 0,
           step_offset: int = 0, save_interval: int = 0, device=None, noise_free: bool = False,
           stream: Optional[torch.cuda.Stream] = None) -> PfResult:
    """Perturbed-solver particle ensemble (src/filters/particle_filter.py:73-118), predict only."""
    n = plan.n
    if x0 is not None:
        _require_cuda(x0, "x0")
        dev = x0.device
        x0_k = x0.to(torch.float64).t().contiguous()
    else:
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise RuntimeError("pf_run needs a CUDA device: there is no CPU fallback")
        x0_k = None
    f64 = dict(dtype=torch.float64, device=dev)
    x0s_h = _host(x0_shared, (n,)) if x0_shared is not None else None
    ths_h = _host(theta_shared, (plan.p,)) if theta_shared is not None else None
    io = N.PfIO()
    io.M, io.T, io.t0 = int(M), int(T), float(t0)
    io.x0, io.x0_shared, io.theta_shared = _dev(x0_k), _hp(x0s_h), _hp(ths_h)
    io.seed, io.particle_offset, io.step_offset = int(seed), int(particle_offset), int(step_offset)
    io.save_interval = int(save_interval)
    io.noise_free = int(bool(noise_free))
    xT = torch.empty(n, M, **f64)
    epsT = torch.empty(n, M, **f64)
    tT = torch.zeros(1, **f64)
    io.xT, io.epsT, io.tT = _dev(xT), _dev(epsT), _dev(tT)
    tr = {}
    if save_interval > 0:
        Ts = T // save_interval + 1
        tr = {"t": torch.empty(Ts, **f64), "x": torch.empty(Ts, n, M, **f64),
              "eps": torch.empty(Ts, n, M, **f64)}
        io.out_t, io.out_x, io.out_eps = _dev(tr["t"]), _dev(tr["x"]), _dev(tr["eps"])
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().odeu_pf_run(plan.handle, C.byref(io), C.c_void_p(st.cuda_stream)),
                "odeu_pf_run")
    traj = None
    if save_interval > 0:
        traj = {"t": tr["t"], "x": tr["x"].permute(0, 2, 1), "eps": tr["eps"].permute(0, 2, 1)}
    return PfResult(xT=xT.t(), epsT=epsT.t(), tT=tT[0], traj=traj)


def launch_count() -> int:
    return int(N.lib().odeu_launch_count())
