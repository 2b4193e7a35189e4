"""Loop owners: GPU replacements of the reference's scripts-level time loops.

    ekf_unroll / run_filter   <- scripts/run_filter.py: main :31-163, unroll :166-224
    ekf_nll                   <- scripts/run_parameter_estimation.py: nll :685-796 (batched over
                                 parameter sets instead of one process per random run, :265-272)
    pf_unroll                 <- unroll driven by ParticleFilter (particle_filter.py:73-118)
    sync_times                <- src/utils.py:181-215 (host preprocessing, numpy)

Everything numerical happens in ONE kernel launch per call (`odeu_ekf_run` / `odeu_pf_run`);
this module only prepares inputs the reference also prepares on the host and reshapes the
outputs into the reference's `traj_states` layout.
"""
from __future__ import annotations

import math
from ast import literal_eval
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _native as N
from .engine import Plan, ekf_grad_run, ekf_run, param_sensitivity, pf_run


# ------------------------------------------------------------------------------------------------
def isin_tolerance(elements: np.ndarray, test_elements: np.ndarray, tol: float) -> np.ndarray:
    """src/utils.py:190-215 (sorted inputs)."""
    elements = np.asarray(elements, dtype=np.float64)
    test_elements = np.asarray(test_elements, dtype=np.float64)
    idx = np.searchsorted(test_elements, elements)
    linvalid = idx == len(test_elements)
    idx = np.where(linvalid, len(test_elements) - 1, idx)
    lval = test_elements[idx] - elements
    lval = np.where(linvalid, -lval, lval)
    rinvalid = idx == 0
    idx1 = np.where(rinvalid, 0, idx - 1)
    rval = elements - test_elements[idx1]
    rval = np.where(rinvalid, -rval, rval)
    return np.minimum(lval, rval) <= tol


def sync_times(ts_x: np.ndarray, ts_y: np.ndarray):
    """src/utils.py:181-187."""
    x_indices = np.nonzero(isin_tolerance(ts_x, ts_y, 1e-8))[0]
    y_indices = np.nonzero(isin_tolerance(ts_y, np.asarray(ts_x)[x_indices], 1e-8))[0]
    assert len(x_indices) == len(y_indices), f"{len(x_indices)} != {len(y_indices)}"
    return x_indices, y_indices


def observation_schedule(t0: float, tN: float, step_size: float, ts_y: np.ndarray):
    """scripts/run_filter.py:95-105: num_steps, correct_flags [T], xy_index_map [T]."""
    num_steps = int(math.ceil((tN - t0) / step_size))
    ts_x = np.arange(t0 + step_size, tN + step_size, step_size)
    x_idx, y_idx = sync_times(ts_x, np.asarray(ts_y, dtype=np.float64))
    flags = np.zeros(ts_x.shape, dtype=bool)
    flags[x_idx] = True
    ymap = np.zeros(ts_x.shape, dtype=np.int64)
    ymap[x_idx] = y_idx
    return num_steps, flags, ymap


def _arr(v, dtype=np.float64):
    if isinstance(v, str):
        v = literal_eval(v)
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=dtype)


def _plan_for(filter_builder, solver_builder, ode_builder, use_static_cov_fn: bool = False) -> Plan:
    if use_static_cov_fn:
        cov_id, scale = filter_builder.static_cov_update_fn_builder.cov_fn_id, filter_builder.static_cov_update_fn_builder.scale
    else:
        cov_id, scale = filter_builder.cov_update_fn_builder.cov_fn_id, filter_builder.cov_update_fn_builder.scale
    return Plan(ode_builder.ode_id, solver_builder.solver_id, solver_builder.h,
                ode_variant=ode_builder.ode_variant, num_compartments=ode_builder.num_compartments_abi,
                cov_fn_id=cov_id, cov_scale=float(scale),
                disable_cov_update=bool(getattr(filter_builder, "disable_cov_update", False)))


def _chol_or_zero(P: torch.Tensor) -> torch.Tensor:
    from .filters import _factor
    return _factor(P)


# ------------------------------------------------------------------------------------------------
def ekf_unroll(filter_builder, solver_builder, ode_builder, *, x0, P0_sqrt, t0: float, num_steps: int,
               measurement_matrix=None, ys=None, correct_flags=None, xy_index_map=None,
               R_sqrt=None, Q_sqrt=None, gamma_sqrt: float = 0.0, save_interval: int = 1,
               use_static_cov_fn: bool = False, params: Optional[Dict[str, np.ndarray]] = None,
               theta: Optional[torch.Tensor] = None, ys_per_trajectory: bool = False,
               device="cuda", reference_layout: bool = True) -> Dict[str, torch.Tensor]:
    """Whole-trajectory EKF for a batch: `unroll()` of scripts/run_filter.py:166-224.

    x0 [N, D] (one trajectory, like the reference) or [B, N, D] / [B, n] (batch).
    Returns the reference's traj_states keys; with `reference_layout` and a single trajectory the
    shapes are the reference's (`t [Ts,1]`, `x [Ts,1,N,D]`, `P_sqrt [Ts,1,n,n]`, ...), otherwise the
    trajectory axis B replaces the singleton axis.  Extra keys: `P`, `S` (full matrices), `nll`.
    """
    dev = torch.device(device)
    plan = _plan_for(filter_builder, solver_builder, ode_builder, use_static_cov_fn)
    Nn, D = ode_builder.shape
    n = plan.n
    x0 = _arr(x0)
    xb = x0.reshape(-1, n)
    B = xb.shape[0]
    ths = ode_builder.flat_params(params if params is not None else ode_builder.params)
    kw = {}
    L = 0
    if measurement_matrix is not None and ys is not None:
        H = _arr(measurement_matrix)
        L = H.shape[0]
        if H.shape[1] != n:
            raise AssertionError("Invalid measurement matrix!")        # run_filter.py:109
        ys_t = torch.as_tensor(_arr(ys)).to(dev)
        kw.update(H=H, R_sqrt=_arr(R_sqrt).reshape(L, L), ys=ys_t, ys_per_trajectory=ys_per_trajectory,
                  correct_flags=torch.as_tensor(_arr(correct_flags, np.uint8)).to(dev),
                  xy_index_map=torch.as_tensor(_arr(xy_index_map, np.int64)).to(dev))
    r = ekf_run(plan, torch.as_tensor(xb).to(dev), int(num_steps), t0=float(t0), P0_sqrt=_arr(P0_sqrt),
                theta=theta, theta_shared=None if theta is not None else ths,
                Q_sqrt=None if Q_sqrt is None else _arr(Q_sqrt), gamma_sqrt=float(gamma_sqrt),
                save_interval=int(save_interval), **kw)
    tr = r.traj
    Ts = tr["t"].shape[0]
    out = {"t": tr["t"].reshape(Ts, 1).expand(Ts, B) if B > 1 else tr["t"].reshape(Ts, 1),
           "x": tr["x"].reshape(Ts, B, Nn, D), "eps": tr["eps"].reshape(Ts, B, Nn, D),
           "P": tr["P"], "P_sqrt": _chol_or_zero(tr["P"])}
    if L > 0:
        out["y_hat"] = tr["y_hat"]
        out["S"] = tr["S"]
        out["S_sqrt"] = _chol_or_zero(tr["S"])
    else:
        out["y_hat"] = torch.zeros(Ts, B, 0, dtype=torch.float64, device=dev)
        out["S_sqrt"] = torch.zeros(Ts, B, 0, 0, dtype=torch.float64, device=dev)
    out["nll"] = r.nll
    return out


def run_filter(output: Optional[str] = None, filter_builder=None, solver_builder=None, ode_builder=None,
               x0="[[1.0, 1.0]]", P0=None, t0: float = 0.0, tN: float = 80.0, ts_y=None, ys_x=None,
               y_path: Optional[str] = None, measurement_matrix=None, obs_noise_var: float = 1e-3,
               seed: int = 7, save_interval: int = 1, use_static_cov_fn: bool = False,
               device="cuda") -> Dict[str, np.ndarray]:
    """`main()` of scripts/run_filter.py:31-163 with the same keyword meaning.  Observations come
    either from `y_path` (an .npz with datasets `t`, `x`, standing in for the reference's H5 file)
    or directly as `ts_y`, `ys_x`.  Writes `output` as .npz with the reference's dataset names."""
    from .filters import SQRT_EKF, ParticleFilter
    from .ode import LotkaVolterra
    from .solvers import Dopri65
    filter_builder = filter_builder if filter_builder is not None else SQRT_EKF()
    solver_builder = solver_builder if solver_builder is not None else Dopri65()
    ode_builder = ode_builder if ode_builder is not None else LotkaVolterra()
    x0_arr = _arr(x0)
    x0_built = ode_builder.build_initial_value(x0_arr, ode_builder.params)
    n = x0_built.size
    P0_sqrt = np.eye(n) * 1e-12 if P0 is None else np.linalg.cholesky(_arr(P0))   # run_filter.py:74-78
    h = solver_builder.h
    num_steps = int(math.ceil((tN - t0) / h))
    if y_path is not None:
        with np.load(y_path) as f:
            ts_y, ys_x = f["t"], f["x"]
    if isinstance(filter_builder, ParticleFilter):
        if ts_y is not None and measurement_matrix is not None:
            raise NotImplementedError          # the reference's PF has no correct step (filter.py:122-133)
        traj = pf_unroll(filter_builder, solver_builder, ode_builder, x0=x0_built, t0=t0,
                         num_steps=num_steps, seed=seed, save_interval=save_interval,
                         use_static_cov_fn=use_static_cov_fn, device=device)
    elif isinstance(filter_builder, SQRT_EKF):
        kw = {}
        if ts_y is not None and measurement_matrix is not None:
            _, flags, ymap = observation_schedule(t0, tN, h, ts_y)
            H = _arr(measurement_matrix)
            ys = np.einsum("ij,tj->ti", H, _arr(ys_x).reshape(-1, H.shape[1]))       # run_filter.py:111
            L = H.shape[0]
            kw = dict(measurement_matrix=H, ys=ys, correct_flags=flags, xy_index_map=ymap,
                      R_sqrt=np.eye(L) * obs_noise_var ** 0.5)
        else:
            print("Prediction only")                                                  # run_filter.py:115
        traj = ekf_unroll(filter_builder, solver_builder, ode_builder, x0=x0_built, P0_sqrt=P0_sqrt, t0=t0,
                          num_steps=num_steps, save_interval=save_interval,
                          use_static_cov_fn=use_static_cov_fn, device=device, **kw)
    else:
        raise ValueError("Unsupported filter builder:", type(filter_builder))        # run_filter.py:146
    res = {k: v.cpu().numpy() for k, v in traj.items() if isinstance(v, torch.Tensor)}
    if output is not None:
        np.savez(output, **{k: v for k, v in res.items() if k != "prng_key"})         # utils.py:90-106
    return res


# ------------------------------------------------------------------------------------------------
def calibration(output: Optional[str] = None, filter_builder=None, solver_builder=None, ode_builder=None,
                x0="[[1.0, 1.0]]", P0=None, t0: float = 0.0, tN: float = 80.0, ts_y=None, ys_x=None,
                y_path: Optional[str] = None, measurement_matrix=None, obs_noise_var: float = 1e-3,
                min_noise_log: float = -8.0, max_noise_log: float = 0.0, num_noise_levels: int = 100,
                device="cuda") -> Dict[str, np.ndarray]:
    """`main()` of scripts/run_calibration_conrad_baseline_calibration.py:31-160: the filter's
    log-likelihood (mean over ALL steps of the nan_to_num'ed per-step terms, :216-220) for
    `num_noise_levels` static process-noise levels (`StaticDiagonalCovarianceUpdate`, scale =
    noise level, :126-142) and for the configured error-driven covariance update (:143-154).
    The reference scans the noise levels one after the other; here they are the batch axis of ONE
    launch (one trajectory per level, per-trajectory scale).  Returns the reference's datasets
    `noise_levels`, `nll_conrad`, `nll_ours` (:156-158)."""
    from .covariance_update_functions import StaticDiagonalCovarianceUpdate
    if measurement_matrix is None:
        raise ValueError("Measurement matrix is required!")
    dev = torch.device(device)
    n = ode_builder.state_dim
    x0_built = ode_builder.build_initial_value(_arr(x0), ode_builder.params).reshape(-1)
    P0_sqrt = np.eye(n) * 1e-12 if P0 is None else np.linalg.cholesky(_arr(P0))
    if y_path is not None:
        dat = np.load(y_path)
        ts_y, ys_x = dat["t"], dat["x"]
    h = solver_builder.h
    num_steps, flags, ymap = observation_schedule(t0, tN, h, ts_y)
    H = _arr(measurement_matrix)
    L = H.shape[0]
    assert H.shape[1] == n, "Invalid measurement matrix!"
    ys = np.einsum("ij,tj->ti", H, _arr(ys_x).reshape(-1, n))
    levels = np.logspace(min_noise_log, max_noise_log, num_noise_levels, endpoint=True)
    kw = dict(t0=float(t0), P0_sqrt=P0_sqrt, theta_shared=ode_builder.flat_params(ode_builder.params), H=H,
              R_sqrt=np.eye(L) * obs_noise_var ** 0.5, ys=torch.as_tensor(ys).to(dev),
              correct_flags=torch.as_tensor(flags.astype(np.uint8)).to(dev),
              xy_index_map=torch.as_tensor(ymap).to(dev), want_final=False, nll_nan_to_num=True)
    static_fb = type(filter_builder)(static_cov_update_fn_builder=StaticDiagonalCovarianceUpdate(1.0),
                                     disable_cov_update=filter_builder.disable_cov_update)
    plan_s = _plan_for(static_fb, solver_builder, ode_builder, use_static_cov_fn=True)
    xb = torch.as_tensor(np.repeat(x0_built[None, :], num_noise_levels, axis=0)).to(dev)
    r = ekf_run(plan_s, xb, num_steps, cov_scale_batch=torch.as_tensor(levels).to(dev), **kw)
    plan_o = _plan_for(filter_builder, solver_builder, ode_builder)
    r_o = ekf_run(plan_o, xb[:1], num_steps, **kw)
    out = {"noise_levels": levels, "nll_conrad": r.nll.cpu().numpy() / num_steps,
           "nll_ours": float(r_o.nll[0]) / num_steps}
    if output is not None:
        np.savez(output, **out)
    return out


def initial_value_and_tangent(ode_builder, x0_raw, flat_sorted: np.ndarray, opt_idx_sorted: np.ndarray,
                              rel_step: float = 1e-6):
    """`initial_state_parametrized` (scripts/run_parameter_estimation.py:744-748): x0(theta) =
    build_initial_value(x0_raw, theta) for every row of `flat_sorted` [B, P] (all parameters, JAX's
    sorted-key order) and its derivative w.r.t. the optimised entries, d x0 / d theta_j [B, p_opt, n]
    - the `x0_tangent` of odeu_ekf_grad_run.  build_initial_value is a cheap host function
    (Hodgkin-Huxley steady-state gates, src/ode/hodgkin_huxley.py:251-281); its derivative is taken
    by central differences with a relative step of 1e-6 (truncation error ~1e-12 relative)."""
    keys_s, sizes, _ = param_layout(ode_builder)

    def build(row):
        pd, o = {}, 0
        for k in keys_s:
            pd[k] = row[o:o + sizes[k]].reshape(np.asarray(ode_builder.params[k]).shape)
            o += sizes[k]
        return ode_builder.build_initial_value(x0_raw, pd).reshape(-1)

    B = flat_sorted.shape[0]
    xb = np.stack([build(flat_sorted[b]) for b in range(B)])
    tan = np.zeros((B, len(opt_idx_sorted), xb.shape[1]))
    for b in range(B):
        for j, k in enumerate(opt_idx_sorted):
            hstep = rel_step * max(1.0, abs(flat_sorted[b, k]))
            up, dn = flat_sorted[b].copy(), flat_sorted[b].copy()
            up[k] += hstep
            dn[k] -= hstep
            tan[b, j] = (build(up) - build(dn)) / (2.0 * hstep)
    return xb, tan


def solve_trajectory(plan: Plan, x0, num_steps: int, *, t0: float = 0.0, theta_shared=None, device="cuda") -> np.ndarray:
    """Plain fixed-step solution x(t0), x(t0 + h), ... [num_steps + 1, n] of ONE initial value with the
    plan's solver (what scripts/run_ode_solver.py:56-74 produces as ground truth): the noise-free
    particle of the ensemble kernel (a plain RK step, no tangents), every step saved.  Used to
    synthesise observation sequences on the product path (bench.py, tools/)."""
    n = plan.n
    r = pf_run(plan, 1, int(num_steps), x0_shared=_arr(x0).reshape(n), t0=float(t0), theta_shared=theta_shared,
               save_interval=1, device=device, noise_free=True)
    return r.traj["x"][:, 0, :].cpu().numpy()


def param_layout(ode_builder):
    """Index bookkeeping between JAX's flattening of the parameter dict (sorted keys,
    `ravel_pytree`) and the builder-order `theta` of the C ABI.
    Returns (sorted_keys, sizes, perm) with theta_builder = flat_sorted[perm]."""
    keys_b = list(ode_builder.params)
    keys_s = sorted(keys_b)
    sizes = {k: int(np.asarray(ode_builder.params[k]).size) for k in keys_b}
    off_s, o = {}, 0
    for k in keys_s:
        off_s[k] = o
        o += sizes[k]
    perm = np.concatenate([np.arange(off_s[k], off_s[k] + sizes[k]) for k in keys_b])
    return keys_s, sizes, perm


def ekf_nll(filter_builder, solver_builder, ode_builder, *, params_norm, params_min, params_max,
            params_optimized: Optional[Dict[str, bool]] = None, x0, P0_sqrt, t0: float, num_steps: int,
            measurement_matrix, ys, correct_flags, xy_index_map, R_sqrt, Q_sqrt, gamma_sqrt: float,
            initial_state_parametrized: bool = False, parameter_sensitivity: bool = False,
            device="cuda") -> torch.Tensor:
    """Batched `nll()` of scripts/run_parameter_estimation.py:685-796.

    params_norm: dict key -> [B, size] normalised values in [0, 1] for the optimised parameters
    (the reference's `params_norms`, one row per random run, :174-201), or a flat array
    [B, p_opt] in JAX's sorted-key order.  Returns NLL [B] (one launch for all B)."""
    dev = torch.device(device)
    plan = _plan_for(filter_builder, solver_builder, ode_builder)
    keys_s, sizes, perm = param_layout(ode_builder)
    opt = {k: True for k in keys_s} if params_optimized is None else params_optimized
    default_flat = np.concatenate([np.asarray(ode_builder.params[k], dtype=np.float64).reshape(-1) for k in keys_s])
    opt_keys = [k for k in keys_s if opt[k]]
    lo = np.concatenate([np.full(sizes[k], params_min[k][0] if np.ndim(params_min[k]) else params_min[k]) for k in opt_keys])
    hi = np.concatenate([np.full(sizes[k], params_max[k][0] if np.ndim(params_max[k]) else params_max[k]) for k in opt_keys])
    if isinstance(params_norm, dict):
        pn = np.concatenate([_arr(params_norm[k]).reshape(-1, sizes[k]) for k in opt_keys], axis=1)
    else:
        pn = _arr(params_norm).reshape(-1, lo.size)
    B = pn.shape[0]
    vals = pn * (hi - lo) + lo                                          # inv_normalize, utils.py:156-178
    opt_idx = np.concatenate([np.arange(sizes[k]) + sum(sizes[q] for q in keys_s[:keys_s.index(k)]) for k in opt_keys])
    flat = np.repeat(default_flat[None, :], B, axis=0)
    flat[:, opt_idx] = vals                                             # :735-742
    theta = flat[:, perm]
    x0_arr = _arr(x0)
    if initial_state_parametrized:                                      # :744-748
        xs = []
        for b in range(B):
            pd, o = {}, 0
            for k in keys_s:
                pd[k] = flat[b, o:o + sizes[k]].reshape(np.asarray(ode_builder.params[k]).shape)
                o += sizes[k]
            xs.append(ode_builder.build_initial_value(x0_arr, pd).reshape(-1))
        xb = np.stack(xs)
    else:
        xb = np.repeat(ode_builder.build_initial_value(x0_arr, ode_builder.params).reshape(1, -1)
                       if x0_arr.size != plan.n else x0_arr.reshape(1, -1), B, axis=0)
    H = _arr(measurement_matrix)
    L = H.shape[0]
    if parameter_sensitivity:                                           # :750-769: Q_sqrt = diag(w(theta))
        gidx = np.argsort(perm)[opt_idx]                                # builder positions of the optimised entries
        xb_d, th_d = torch.as_tensor(xb).to(dev), torch.as_tensor(theta).to(dev)
        w, _ = param_sensitivity(plan, xb_d, gidx, t0=float(t0), theta=th_d, want_tangent=False)
        nll, _ = ekf_grad_run(plan, xb_d, int(num_steps), gidx[:1], t0=float(t0), P0_sqrt=_arr(P0_sqrt), theta=th_d,
                              gamma_sqrt=float(gamma_sqrt), H=H, R_sqrt=_arr(R_sqrt).reshape(L, L),
                              ys=torch.as_tensor(_arr(ys)).to(dev),
                              correct_flags=torch.as_tensor(_arr(correct_flags, np.uint8)).to(dev),
                              xy_index_map=torch.as_tensor(_arr(xy_index_map, np.int64)).to(dev), Q_sqrt_diag=w)
        return nll
    r = ekf_run(plan, torch.as_tensor(xb).to(dev), int(num_steps), t0=float(t0), P0_sqrt=_arr(P0_sqrt),
                theta=torch.as_tensor(theta).to(dev), Q_sqrt=_arr(Q_sqrt), gamma_sqrt=float(gamma_sqrt),
                H=H, R_sqrt=_arr(R_sqrt).reshape(L, L), ys=torch.as_tensor(_arr(ys)).to(dev),
                correct_flags=torch.as_tensor(_arr(correct_flags, np.uint8)).to(dev),
                xy_index_map=torch.as_tensor(_arr(xy_index_map, np.int64)).to(dev), want_final=False,
                minimal=True)
    return r.nll


# ------------------------------------------------------------------------------------------------
def baseline_nll(solver_builder, ode_builder, **kw) -> torch.Tensor:
    """Batched plain-RK least-squares loss, `nll()` of scripts/run_parameter_estimation_baseline.py:552-632,
    with the keyword meaning of `ekf_nll` (no P0 / Q / gamma): the degenerate filter P0 = 0, Q = 0,
    disable_cov_update has zero gain, so each term is negative_log_gaussian_sqrt(y, H x_RK, R_sqrt)."""
    from .filters import SQRT_EKF

    n = ode_builder.build_initial_value(_arr(kw["x0"]), ode_builder.params).size
    return ekf_nll(SQRT_EKF(disable_cov_update=True), solver_builder, ode_builder, P0_sqrt=np.zeros((n, n)),
                   Q_sqrt=np.zeros((n, n)), gamma_sqrt=0.0, **kw)


def pf_unroll(filter_builder, solver_builder, ode_builder, *, x0, t0: float, num_steps: int, seed: int = 7,
              save_interval: int = 1, use_static_cov_fn: bool = False, particle_offset: int = 0,
              num_particles: Optional[int] = None, device="cuda") -> Dict[str, torch.Tensor]:
    """`unroll()` with the particle ensemble: traj_states `t [Ts,M]`, `x [Ts,M,N,D]`, `eps`."""
    plan = _plan_for(filter_builder, solver_builder, ode_builder, use_static_cov_fn)
    M = int(num_particles if num_particles is not None else filter_builder.M)
    Nn, D = ode_builder.shape
    r = pf_run(plan, M, int(num_steps), x0_shared=_arr(x0).reshape(-1), t0=float(t0),
               theta_shared=ode_builder.flat_params(ode_builder.params), seed=int(seed),
               particle_offset=int(particle_offset), save_interval=int(save_interval), device=device)
    tr = r.traj
    Ts = tr["t"].shape[0]
    return {"t": tr["t"].reshape(Ts, 1).expand(Ts, M), "x": tr["x"].reshape(Ts, M, Nn, D),
            "eps": tr["eps"].reshape(Ts, M, Nn, D)}
