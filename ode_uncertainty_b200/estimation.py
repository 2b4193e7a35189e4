"""Parameter estimation with process-noise tempering: the caller of the hot path.

Mirror of `optimize()` / `optimize_run()` of scripts/run_parameter_estimation.py:49-308, 540-682:
for every tempering stage gamma_s (noise schedule; last stage 0, :621-623) each random run
minimises the EKF negative log-likelihood over the normalised parameters in [0, 1]^p with
SciPy's L-BFGS-B - the same optimiser the reference reaches through
jaxopt.ScipyBoundedMinimize (:599).

What changes is where the (loss, gradient) pairs come from: the reference runs one process per
random run (`p_umap`, :265-272) and differentiates in reverse mode on the CPU; here all runs
advance in lock step - every optimiser lives in its own host thread, its requests are collected
by `BatchedObjective`, and ONE `odeu_ekf_grad_run` launch (forward-mode gradient fused with the
filter) serves all pending runs.  SciPy sees exactly the function it would see alone, so the
iterates of a run do not depend on the batching.

Differences to the reference: initial parameters come from NumPy's `default_rng(seed)` instead
of `jax.random.uniform(key(seed))` (:174-201); results are returned as a dict of NumPy arrays
with the reference's dataset names (:297-306) instead of being written to H5.
"""
from __future__ import annotations

import math
import threading
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .engine import ekf_grad_run, param_sensitivity
from .noise_schedules import ExponentialDecaySchedule, NoiseSchedule
from .runners import _arr, _plan_for, initial_value_and_tangent, observation_schedule, param_layout


class BatchedObjective:
    """Lock-step evaluation service for R concurrent optimisers.

    `evaluate(r, z)` blocks until every optimiser that is still running has posted its next
    point, then one batched kernel launch computes all of them."""

    def __init__(self, num_runs: int, batch_fn) -> None:
        self.R, self.batch_fn = num_runs, batch_fn
        self.cv = threading.Condition()
        self.pending: Dict[int, np.ndarray] = {}
        self.results: Dict[int, Tuple[float, np.ndarray]] = {}
        self.active = num_runs
        self.launches = 0
        self.error: Optional[BaseException] = None

    def _flush_locked(self) -> None:
        runs = sorted(self.pending)
        Z = np.stack([self.pending[r] for r in runs])
        try:
            f, g = self.batch_fn(np.array(runs), Z)
            for k, r in enumerate(runs):
                self.results[r] = (float(f[k]), np.array(g[k], dtype=np.float64))
        except BaseException as exc:  # propagate to every waiting optimiser
            self.error = exc
            for r in runs:
                self.results[r] = (float("nan"), np.zeros(Z.shape[1]))
        self.launches += 1
        self.pending.clear()
        self.cv.notify_all()

    def evaluate(self, r: int, z: np.ndarray) -> Tuple[float, np.ndarray]:
        with self.cv:
            self.pending[r] = np.array(z, dtype=np.float64)
            if len(self.pending) == self.active:
                self._flush_locked()
            while r not in self.results:
                self.cv.wait()
            out = self.results.pop(r)
            if self.error is not None:
                raise RuntimeError(str(self.error))
            return out

    def finish(self, r: int) -> None:
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) == self.active:
                self._flush_locked()


def optimize(filter_builder, solver_builder, ode_builder, *, x0, ts_y, ys_x, measurement_matrix,
             params_range: Dict[str, Tuple[float, float]], gamma_noise_weights,
             params_optimized: Optional[Dict[str, bool]] = None, P0=None, t0: float = 0.0, tN: float = 80.0,
             num_tempering_stages: int = 10, final_gamma_zero: bool = True, obs_noise_var: float = 0.1,
             gamma_noise_schedule: NoiseSchedule = ExponentialDecaySchedule(), lbfgs_maxiter: int = 200,
             num_random_runs: int = 0, seed: int = 7, initial_state_parametrized: bool = False,
             parameter_sensitivity: bool = False, device="cuda", verbose: bool = False,
             _P0_sqrt: Optional[np.ndarray] = None, optimizer: str = "scipy", check_every: int = 16,
             _z0: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """scripts/run_parameter_estimation.py:49-308 with the same keyword meaning (observations
    are passed as arrays `ts_y`, `ys_x` instead of an H5 path).

    optimizer: "scipy" - SciPy's L-BFGS-B per random run, like the reference (through jaxopt, :599), the
    runs advancing in lock step (one host thread each; fine for tens of runs);  "device" - the
    lock-step projected L-BFGS of csrc/lbfgs.cu: iterates, curvature history and line search of ALL
    runs stay on the GPU, one `odeu_lbfgs_step` launch per batched objective evaluation, no host thread
    per run and no device-to-host read inside a tempering stage (`check_every` = 0; a positive value
    reads ONE flag every that many evaluations to stop early once every run has converged).  The device
    optimiser is not SciPy's algorithm (projected path + Armijo instead of Cauchy point + More-Thuente),
    so its iterates differ while the optima agree (tests/test_estimation.py)."""
    from scipy.optimize import minimize

    if measurement_matrix is None:
        raise ValueError("Measurement matrix is required!")              # :119-120
    if gamma_noise_weights is None:
        raise ValueError("Gamma noise weight vector is required!")       # :121-122
    if params_range is None:
        raise ValueError("Parameter ranges are required!")               # :123-124
    dev = torch.device(device)
    plan = _plan_for(filter_builder, solver_builder, ode_builder)
    n = plan.n
    x0_built = ode_builder.build_initial_value(_arr(x0), ode_builder.params).reshape(-1)
    P0_sqrt = np.eye(n) * 1e-12 if P0 is None else np.linalg.cholesky(_arr(P0))
    if _P0_sqrt is not None:
        P0_sqrt = _arr(_P0_sqrt).reshape(n, n)
    h = solver_builder.h
    num_steps, flags, ymap = observation_schedule(t0, tN, h, ts_y)
    H = _arr(measurement_matrix)
    L = H.shape[0]
    assert H.shape[1] == n, "Invalid measurement matrix!"
    ys = np.einsum("ij,tj->ti", H, _arr(ys_x).reshape(-1, n))            # :147
    w = _arr(gamma_noise_weights)
    assert w.shape[0] == n, "Invalid gamma noise weight vector!"
    keys_s, sizes, perm = param_layout(ode_builder)
    opt = {k: True for k in keys_s} if params_optimized is None else params_optimized
    assert len(params_range) == len(ode_builder.params), "Invalid parameter ranges!"
    opt_keys = [k for k in keys_s if opt[k]]
    lo = np.concatenate([np.full(sizes[k], params_range[k][0]) for k in opt_keys])
    hi = np.concatenate([np.full(sizes[k], params_range[k][1]) for k in opt_keys])
    default_sorted = np.concatenate([np.asarray(ode_builder.params[k], dtype=np.float64).reshape(-1) for k in keys_s])
    off = np.cumsum([0] + [sizes[k] for k in keys_s])
    opt_idx_sorted = np.concatenate([np.arange(off[keys_s.index(k)], off[keys_s.index(k) + 1]) for k in opt_keys])
    inv_perm = np.argsort(perm)                      # builder position of each sorted entry
    grad_idx_builder = inv_perm[opt_idx_sorted]
    p = lo.size
    if num_random_runs > 0:
        rng = np.random.default_rng(seed)
        z0 = rng.uniform(0.0, 1.0, (num_random_runs, p))
        R = num_random_runs
    else:                                            # single run from the defaults (:203-220)
        z0 = ((default_sorted[opt_idx_sorted] - lo) / (hi - lo))[None, :]
        R = 1
    ys_d = torch.as_tensor(ys).to(dev)
    flags_d = torch.as_tensor(flags.astype(np.uint8)).to(dev)
    ymap_d = torch.as_tensor(ymap).to(dev)
    x0_all = torch.as_tensor(np.repeat(x0_built[None, :], R, axis=0)).to(dev)
    R_sqrt = np.eye(L) * obs_noise_var ** 0.5
    Q_sqrt = np.diag(w)

    def batch_fn_factory(gamma):
        def batch_fn(runs, Z):
            flat = np.repeat(default_sorted[None, :], len(runs), axis=0)
            flat[:, opt_idx_sorted] = Z * (hi - lo) + lo                 # inv_normalize (:735-742)
            theta = torch.as_tensor(flat[:, perm]).to(dev)
            x0_b, x0_tan = x0_all[: len(runs)], None
            if initial_state_parametrized:               # x0 = build_initial_value(V0, theta), :744-748
                xb, tan = initial_value_and_tangent(ode_builder, _arr(x0), flat, opt_idx_sorted)
                x0_b, x0_tan = torch.as_tensor(xb).to(dev), torch.as_tensor(tan).to(dev)
            qd = qd_tan = None
            if parameter_sensitivity:                    # Q_sqrt = diag(w(theta)) inside the loss, :750-769
                qd, qd_tan = param_sensitivity(plan, x0_b, grad_idx_builder, t0=t0, theta=theta, x0_tangent=x0_tan)
            nll, g = ekf_grad_run(plan, x0_b, num_steps, grad_idx_builder, t0=t0,
                                  P0_sqrt=P0_sqrt, theta=theta, Q_sqrt=Q_sqrt, gamma_sqrt=gamma ** 0.5,
                                  H=H, R_sqrt=R_sqrt, ys=ys_d, correct_flags=flags_d, xy_index_map=ymap_d,
                                  x0_tangent=x0_tan, Q_sqrt_diag=qd, Q_sqrt_diag_tangent=qd_tan)
            return nll.cpu().numpy(), g.cpu().numpy() * (hi - lo)        # d/d theta_norm
        return batch_fn

    if optimizer not in ("scipy", "device"):
        raise ValueError(f"optimizer must be 'scipy' or 'device', got {optimizer!r}")
    if _z0 is not None:
        z0 = np.asarray(_z0, dtype=np.float64).reshape(R, p)
    if optimizer == "device":
        if initial_state_parametrized:
            raise ValueError("optimizer='device' does not serve initial_state_parametrized (x0(theta) is a host "
                             "function of the ODE builder); use optimizer='scipy'")
        return _optimize_device(plan, R, p, z0, lo, hi, default_sorted, opt_idx_sorted, perm, grad_idx_builder, x0_all,
                                num_steps, dict(t0=t0, P0_sqrt=P0_sqrt, H=H, R_sqrt=R_sqrt, ys=ys_d, correct_flags=flags_d,
                                                xy_index_map=ymap_d), Q_sqrt, parameter_sensitivity, gamma_noise_schedule,
                                num_tempering_stages, final_gamma_zero, lbfgs_maxiter, check_every, opt_keys, sizes, verbose)
    params_optims = np.zeros((R, num_tempering_stages, p))
    nll_optims = np.zeros((R, num_tempering_stages))
    iters = np.zeros((R, num_tempering_stages), dtype=np.int64)
    nfev = np.zeros((R, num_tempering_stages), dtype=np.int64)
    z = z0.copy()
    gammas, launches = [], 0
    for stage in range(num_tempering_stages):
        gamma = float(gamma_noise_schedule.step(stage))
        if final_gamma_zero and stage + 1 == num_tempering_stages:
            gamma = 0.0                                                   # :621-623
        gammas.append(gamma)
        svc = BatchedObjective(R, batch_fn_factory(gamma))
        results = [None] * R

        def worker(r):
            try:
                fun = lambda zz, r=r: svc.evaluate(r, zz)
                results[r] = minimize(fun, z[r], jac=True, method="L-BFGS-B", bounds=[(0.0, 1.0)] * p,
                                      options={"maxiter": lbfgs_maxiter})
            except RuntimeError as err:                                   # :657-667: record zeros, go on
                results[r] = err
            finally:
                svc.finish(r)

        threads = [threading.Thread(target=worker, args=(r,)) for r in range(R)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        launches += svc.launches
        for r in range(R):
            res = results[r]
            if isinstance(res, BaseException) or res is None:
                params_optims[r, stage] = z[r] * (hi - lo) + lo
                continue
            z[r] = np.clip(res.x, 0.0, 1.0)
            params_optims[r, stage] = z[r] * (hi - lo) + lo
            nll_optims[r, stage] = res.fun
            iters[r, stage] = res.nit
            nfev[r, stage] = res.nfev
        if verbose:
            print(f"stage {stage + 1}/{num_tempering_stages} gamma={gamma:g} "
                  f"best nll={nll_optims[:, stage].min():.6g} launches={svc.launches}")
    names = [k for k in opt_keys for _ in range(sizes[k])]
    return {"params_inits": z0 * (hi - lo) + lo, "params_optims": params_optims,
            "params_default": default_sorted[opt_idx_sorted], "params_name": np.array(names),
            "nll_optims": nll_optims, "num_lbfgs_iters": iters, "num_nll_evals": nfev,
            "num_nll_jac_evals": nfev.copy(), "gammas": np.array(gammas), "kernel_launches": launches}


def _optimize_device(plan, R, p, z0, lo, hi, default_sorted, opt_idx_sorted, perm, grad_idx_builder, x0_all, num_steps,
                     kw, Q_sqrt, parameter_sensitivity, schedule, num_stages, final_gamma_zero, maxiter, check_every,
                     opt_keys, sizes, verbose) -> Dict[str, np.ndarray]:
    """Tempered estimation with the device-resident lock-step L-BFGS (csrc/lbfgs.cu)."""
    import ctypes as C

    from . import _native as N
    dev = x0_all.device
    f64 = dict(dtype=torch.float64, device=dev)
    lib = N.lib()
    lo_d, hi_d = torch.as_tensor(lo, **f64), torch.as_tensor(hi, **f64)
    scale_d = (hi_d - lo_d).contiguous()
    flat0 = torch.as_tensor(default_sorted, **f64).repeat(R, 1)
    oi = torch.as_tensor(opt_idx_sorted, device=dev)
    pm = torch.as_tensor(perm, device=dev)
    zt = torch.as_tensor(z0, **f64).contiguous()
    nws = int(lib.odeu_lbfgs_workspace_doubles(R, p))
    ptr = lambda t_: C.c_void_p(t_.data_ptr())
    evals = 0

    def evaluate(gamma):
        nonlocal evals
        flat = flat0.clone()
        flat[:, oi] = zt * scale_d + lo_d                                  # inv_normalize (:735-742), on the device
        theta = flat[:, pm].contiguous()
        qd = qd_tan = None
        if parameter_sensitivity:
            qd, qd_tan = param_sensitivity(plan, x0_all, grad_idx_builder, t0=kw["t0"], theta=theta)
        nll, g = ekf_grad_run(plan, x0_all, num_steps, grad_idx_builder, theta=theta, Q_sqrt=Q_sqrt, gamma_sqrt=gamma ** 0.5,
                              Q_sqrt_diag=qd, Q_sqrt_diag_tangent=qd_tan, **kw)
        evals += 1
        return nll, g.contiguous()

    params_optims = np.zeros((R, num_stages, p))
    nll_optims = np.zeros((R, num_stages))
    iters = np.zeros((R, num_stages), dtype=np.int64)
    nfev = np.zeros((R, num_stages), dtype=np.int64)
    status = np.zeros((R, num_stages), dtype=np.int64)
    gammas, host_reads = [], 0
    max_evals = int(1.5 * maxiter) + 8
    for stage in range(num_stages):
        gamma = float(schedule.step(stage))
        if final_gamma_zero and stage + 1 == num_stages:
            gamma = 0.0
        gammas.append(gamma)
        ws = torch.zeros(nws, **f64)
        meta = ws[nws - 3 * R:].view(torch.int32).view(R, 6)      # the workspace ends with the int32 bookkeeping
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for k in range(max_evals):
            nll, g = evaluate(gamma)
            with torch.cuda.device(dev):
                N.check(lib.odeu_lbfgs_step(R, p, int(maxiter), int(k == 0), 1e-5, 1e7 * np.finfo(float).eps, ptr(ws), ptr(zt),
                                            ptr(nll), ptr(g), ptr(scale_d), st), "odeu_lbfgs_step")
            if check_every > 0 and (k + 1) % check_every == 0:
                host_reads += 1
                if not bool((meta[:, 4] == 0).any()):                       # ONE flag: is any restart still running?
                    break
        z_fin = ws[: R * p].view(R, p)
        f_fin = ws[3 * R * p: 3 * R * p + R]
        zt = z_fin.clone().contiguous()                                     # warm start of the next stage (:626-651)
        m_h = meta.cpu().numpy()                                            # the stage's results: one read
        params_optims[:, stage] = z_fin.cpu().numpy() * (hi - lo) + lo
        nll_optims[:, stage] = f_fin.cpu().numpy()
        iters[:, stage], nfev[:, stage], status[:, stage] = m_h[:, 2], m_h[:, 3], m_h[:, 4]
        if verbose:
            print(f"stage {stage + 1}/{num_stages} gamma={gamma:g} best nll={nll_optims[:, stage].min():.6g}")
    names = [k for k in opt_keys for _ in range(sizes[k])]
    return {"params_inits": z0 * (hi - lo) + lo, "params_optims": params_optims,
            "params_default": default_sorted[opt_idx_sorted], "params_name": np.array(names),
            "nll_optims": nll_optims, "num_lbfgs_iters": iters, "num_nll_evals": nfev, "num_nll_jac_evals": nfev.copy(),
            "gammas": np.array(gammas), "kernel_launches": evals, "lbfgs_status": status,
            "host_reads_inside_stages": host_reads}


def evaluate(filter_builder, solver_builder, ode_builder, *, x0, ts_y, ys_x, measurement_matrix,
             params_range: Dict[str, Tuple[float, float]], gamma_noise_weights, num_param_evals: Dict[str, int],
             params_optimized: Optional[Dict[str, bool]] = None, P0=None, t0: float = 0.0, tN: float = 80.0,
             num_tempering_stages: int = 10, final_gamma_zero: bool = True, obs_noise_var: float = 0.1,
             gamma_noise_schedule: NoiseSchedule = ExponentialDecaySchedule(), initial_state_parametrized: bool = False,
             parameter_sensitivity: bool = False, device="cuda", **_unused) -> Dict[str, np.ndarray]:
    """`evaluate()` of scripts/run_parameter_estimation.py:311-537: the NLL on the tensor grid
    `linspace(min, max, num_param_evals[k])` over the optimised parameters (sorted-key order, :442-459),
    for every tempering stage.  The reference evaluates the grid points one after the other and times
    each call (`perf_counter_ns`, :496-522 - its only timing instrumentation); here the whole grid is the
    batch axis of ONE launch per stage.  Returns the reference's datasets `param_evals`
    [N_grid, p_opt], `nll_evals` [stages, N_grid], `gammas`, and `timings` (ns per grid point = launch
    time / N_grid, first evaluation of the first stage dropped like :516-517)."""
    import time

    from .engine import ekf_run
    if measurement_matrix is None:
        raise ValueError("Measurement matrix is required!")              # :404-405
    if gamma_noise_weights is None:
        raise ValueError("Gamma noise weight vector is required!")       # :406-407
    if params_range is None:
        raise ValueError("Parameter ranges are required!")               # :408-409
    if num_param_evals is None:
        raise ValueError("Parameter evaluation counts are required!")    # :412-413
    dev = torch.device(device)
    plan = _plan_for(filter_builder, solver_builder, ode_builder)
    n = plan.n
    x0_raw = _arr(x0)
    x0_built = ode_builder.build_initial_value(x0_raw, ode_builder.params).reshape(-1)
    P0_sqrt = np.eye(n) * 1e-12 if P0 is None else np.linalg.cholesky(_arr(P0))
    num_steps, flags, ymap = observation_schedule(t0, tN, solver_builder.h, ts_y)
    H = _arr(measurement_matrix)
    L = H.shape[0]
    assert H.shape[1] == n, "Invalid measurement matrix!"
    ys = np.einsum("ij,tj->ti", H, _arr(ys_x).reshape(-1, n))
    w = _arr(gamma_noise_weights)
    assert w.shape[0] == n, "Invalid gamma noise weight vector!"
    keys_s, sizes, perm = param_layout(ode_builder)
    opt = {k: True for k in keys_s} if params_optimized is None else params_optimized
    default_sorted = np.concatenate([np.asarray(ode_builder.params[k], dtype=np.float64).reshape(-1) for k in keys_s])
    axes = [np.linspace(params_range[k][0], params_range[k][1], int(num_param_evals[k])) for k in keys_s for _ in range(sizes[k])]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(axes))          # :455-457
    off = np.cumsum([0] + [sizes[k] for k in keys_s])
    opt_idx = np.concatenate([np.arange(off[i], off[i + 1]) for i, k in enumerate(keys_s) if opt[k]]).astype(np.int64)
    flat = np.repeat(default_sorted[None, :], grid.shape[0], axis=0)
    flat[:, opt_idx] = grid[:, opt_idx]                                  # non-optimised entries keep their defaults (:735-742)
    theta = torch.as_tensor(flat[:, perm]).to(dev)
    G = grid.shape[0]
    if initial_state_parametrized:
        xb, _ = initial_value_and_tangent(ode_builder, x0_raw, flat, opt_idx[:0])
        x0_b = torch.as_tensor(xb).to(dev)
    else:
        x0_b = torch.as_tensor(np.repeat(x0_built[None, :], G, axis=0)).to(dev)
    kw = dict(t0=t0, P0_sqrt=P0_sqrt, theta=theta, H=H, R_sqrt=np.eye(L) * obs_noise_var ** 0.5,
              ys=torch.as_tensor(ys).to(dev), correct_flags=torch.as_tensor(flags.astype(np.uint8)).to(dev),
              xy_index_map=torch.as_tensor(ymap).to(dev))
    gidx = np.argsort(perm)[opt_idx]
    nll_evals, gammas, timings = [], [], []
    for stage in range(num_tempering_stages):
        gamma = float(gamma_noise_schedule.step(stage))
        if final_gamma_zero and stage + 1 == num_tempering_stages:
            gamma = 0.0
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter_ns()
        if parameter_sensitivity:                                          # Q_sqrt = diag(w(theta)) inside the loss (:750-769)
            qd, _ = param_sensitivity(plan, x0_b, gidx, t0=t0, theta=theta, want_tangent=False)
            nll, _ = ekf_grad_run(plan, x0_b, num_steps, gidx[:1], gamma_sqrt=gamma ** 0.5, Q_sqrt_diag=qd, **kw)
        else:
            nll = ekf_run(plan, x0_b, num_steps, Q_sqrt=np.diag(w), gamma_sqrt=gamma ** 0.5, want_final=False, minimal=True, **kw).nll
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter_ns()
        nll_evals.append(nll.cpu().numpy())
        gammas.append(gamma)
        timings += [(t2 - t1) / G] * (G - 1 if stage == 0 else G)
    return {"param_evals": grid[:, opt_idx], "nll_evals": np.stack(nll_evals), "gammas": np.array(gammas),
            "timings": np.array(timings)}


def optimize_baseline(solver_builder, ode_builder, *, x0, ts_y, ys_x, measurement_matrix,
                      params_range: Dict[str, Tuple[float, float]],
                      params_optimized: Optional[Dict[str, bool]] = None, t0: float = 0.0, tN: float = 80.0,
                      obs_noise_var: float = 0.1, lbfgs_maxiter: int = 200, num_random_runs: int = 0, seed: int = 7,
                      initial_state_parametrized: bool = False, device="cuda",
                      verbose: bool = False) -> Dict[str, np.ndarray]:
    """scripts/run_parameter_estimation_baseline.py:40-262: the plain-RK least-squares baseline,
    loss `nll` :552-632 = sum over observation steps of negative_log_gaussian_sqrt(y, H x_RK(theta), R_sqrt).

    Served by the SAME gradient kernels with a degenerate filter: P0 = 0, no process noise
    (disable_cov_update, Q = 0) keep P = 0, hence S = R, the gain is exactly zero, the state is the
    plain RK solution and every NLL term is the baseline's ("same kernel, different
    reduction/branch").  One stage, no tempering; arrays come back without the stage axis like the
    reference's datasets (:236-259)."""
    from .filters import SQRT_EKF

    n = ode_builder.build_initial_value(_arr(x0), ode_builder.params).size
    res = optimize(SQRT_EKF(disable_cov_update=True), solver_builder, ode_builder, x0=x0, ts_y=ts_y, ys_x=ys_x,
                   measurement_matrix=measurement_matrix, params_range=params_range,
                   gamma_noise_weights=np.zeros(n), params_optimized=params_optimized, t0=t0, tN=tN,
                   num_tempering_stages=1, final_gamma_zero=True, obs_noise_var=obs_noise_var,
                   lbfgs_maxiter=lbfgs_maxiter, num_random_runs=num_random_runs, seed=seed,
                   initial_state_parametrized=initial_state_parametrized, device=device, verbose=verbose,
                   _P0_sqrt=np.zeros((n, n)))
    for k in ("params_optims", "nll_optims", "num_lbfgs_iters", "num_nll_evals", "num_nll_jac_evals"):
        res[k] = res[k][:, 0]
    res.pop("gammas")
    return res
