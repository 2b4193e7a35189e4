"""Marshalling between the REFERENCE'S OWN plugin objects / state dictionaries and the C ABI.

This is the body of the reference-side binding `integration/b200.py::B200_SQRT_EKF` (a subclass of
the reference's `src.filters.sqrt_ekf.SQRT_EKF`).  It is kept here, free of any import from the
reference tree, so that it can be exercised on a GPU box that does not carry `/root/reference`:
everything is duck-typed on the reference's class NAMES and attributes

    ode_builder      type name (src/ode/__init__.py), `.params` dict in constructor order
                     (src/ode/ode.py:18-23), `.model` / `.num_compartments` (hodgkin_huxley.py)
    solver_builder   type name (src/solvers/__init__.py), `.h` (src/solvers/solver.py:18-26)
    filter_builder   `.cov_update_fn_builder`, `.static_cov_update_fn_builder` (type name, `.scale`),
                     `.disable_cov_update` (src/filters/sqrt_ekf.py:36-43)
    initial_state    the dict of SQRT_EKF.init_state (sqrt_ekf.py:45-84): t [1], x [1,N,D], P_sqrt
                     [1,n,n], Q_sqrt [n,n], gamma_sqrt [], R_sqrt [L,L]

and returns plain numpy arrays in the reference's `traj_states` layout (scripts/run_filter.py:
219-222): leading axis = saved steps, then the singleton batch axis of `init_state`.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np

from . import _native as N

ODE_IDS = {"Lorenz": N.ODE_LORENZ, "VanDerPol": N.ODE_VAN_DER_POL, "LotkaVolterra": N.ODE_LOTKA_VOLTERRA,
           "Pendulum": N.ODE_PENDULUM, "LCAO": N.ODE_LCAO, "HodgkinHuxley": N.ODE_HODGKIN_HUXLEY,
           "MultiCompartmentHodgkinHuxley": N.ODE_MULTI_HH}
SOLVER_IDS = {"RKF45": N.SOLVER_RKF45, "Dopri65": N.SOLVER_DOPRI65, "BS32": N.SOLVER_BS32,
              "HeunEuler": N.SOLVER_HEUN_EULER, "Kvaerno3": N.SOLVER_KVAERNO3, "ImplicitEuler": N.SOLVER_IMPLICIT_EULER}
COV_IDS = {"DiagonalCovarianceUpdate": N.COV_DIAGONAL, "OuterCovarianceUpdate": N.COV_OUTER,
           "StaticDiagonalCovarianceUpdate": N.COV_STATIC_DIAGONAL}
HH_VARIANT = {"full": 0, "reduced-1": 1, "reduced-4": 4}


def _np(a, dtype=np.float64) -> np.ndarray:
    """jax.Array / torch.Tensor / nested lists -> contiguous numpy."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.asarray(a, dtype=dtype)
    return a if a.flags.c_contiguous else np.ascontiguousarray(a)     # (ascontiguousarray would promote 0-d to 1-d)


def flat_theta(params: Dict[str, object]) -> np.ndarray:
    """Flat parameter vector in `ODEBuilder.params` (constructor keyword) order = the `theta` layout
    of include/odeu.h."""
    return np.concatenate([_np(v).reshape(-1) for v in params.values()])


def plan_kwargs(filter_builder, solver_builder, ode_builder, use_static_cov_fn: bool = False,
                state_shape=None) -> dict:
    """`odeu_plan_desc` for the reference's plugin objects (what jsonargparse instantiated from
    class_path / init_args, scripts/run_filter.py:31-47)."""
    oname, sname = type(ode_builder).__name__, type(solver_builder).__name__
    if sname == "DiffraxSolverBuilder":       # the diffrax solver object's class names the method (diffrax_solver.py:29-34)
        sname = getattr(solver_builder, "name", None) or type(solver_builder.solver).__name__
    if oname not in ODE_IDS:
        raise ValueError(f"Unsupported ODE builder: {oname}")
    if sname not in SOLVER_IDS:
        raise ValueError(f"Unsupported solver builder for the B200 path: {sname}")
    variant, nc = 0, 0
    if oname in ("HodgkinHuxley", "MultiCompartmentHodgkinHuxley"):
        model = ode_builder.model if oname == "HodgkinHuxley" else ode_builder.single_compartment_model.model
        variant = HH_VARIANT[model]
        nc = int(getattr(ode_builder, "num_compartments", 0)) if oname == "MultiCompartmentHodgkinHuxley" else 0
    elif oname == "LCAO":
        variant = int(state_shape[-1]) if state_shape is not None else 2     # D of the [2, D] state (lcao.py:51-61)
    cov = filter_builder.static_cov_update_fn_builder if use_static_cov_fn else filter_builder.cov_update_fn_builder
    cname = type(cov).__name__
    if cname not in COV_IDS:
        raise ValueError(f"Unsupported covariance update function: {cname}")
    return dict(ode_id=ODE_IDS[oname], solver_id=SOLVER_IDS[sname], step_size=float(solver_builder.h),
                ode_variant=variant, num_compartments=nc, cov_fn_id=COV_IDS[cname], cov_scale=float(cov.scale),
                disable_cov_update=bool(getattr(filter_builder, "disable_cov_update", False)))


def _default_runner():
    """The product path: ode_uncertainty_b200.engine.ekf_run on cuda:0 (no CPU fallback).  Returns a
    callable (plan_kwargs, x0 [B,n], T, **arrays) -> dict of numpy arrays in tests/util.run_ekf's layout."""
    import torch

    from .engine import Plan, ekf_run

    def run(pk, x0, T, *, save_interval, guard, **kw):
        if not torch.cuda.is_available():
            raise RuntimeError("B200_SQRT_EKF needs a CUDA device: the EKF-RK path has no CPU fallback")
        dev = torch.device("cuda")
        plan = Plan(**pk)
        t = lambda a, dt=torch.float64: None if a is None else torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
        r = ekf_run(plan, t(x0), int(T), t0=kw["t0"], P0_sqrt=kw["P0_sqrt"], theta_shared=kw.get("theta_shared"),
                    Q_sqrt=kw.get("Q_sqrt"), gamma_sqrt=kw.get("gamma_sqrt", 0.0), H=kw.get("H"), R_sqrt=kw.get("R_sqrt"),
                    ys=t(kw.get("ys")), correct_flags=t(kw.get("correct_flags"), torch.uint8),
                    xy_index_map=t(kw.get("xy_index_map"), torch.int64), save_interval=int(save_interval), guard=guard)
        c = lambda v: None if v is None else v.cpu().numpy()
        out = dict(nll=c(r.nll), xT=c(r.xT), PT=c(r.PT))
        out["traj"] = None if r.traj is None else {k: c(v) for k, v in r.traj.items()}
        return out

    return run


_runner: Optional[Callable] = None


def set_runner(fn: Optional[Callable]) -> None:
    """Test hook: replace the engine call (e.g. by the host-compiled kernel source) - never used by the
    product."""
    global _runner
    _runner = fn


def _run(*a, **k):
    return (_runner or _default_runner())(*a, **k)


def resolve_guard(guard: str, n: int) -> str:
    """"reference" (the verbatim sign-sensitive predicate, factor-form kernels) is served for n <= 4;
    larger systems run the intended predicate on the full-covariance kernels."""
    if guard == "auto":
        return "reference" if n <= 4 else "intended"
    return guard


def unroll(filter_builder, solver_builder, ode_builder, use_static_cov_fn, measurement_matrix, initial_state,
           ys, correct_flags, xy_index_map, num_steps: int, save_interval: int, guard: str = "auto",
           params: Optional[Dict[str, object]] = None) -> Dict[str, np.ndarray]:
    """`unroll()` of scripts/run_filter.py:166-224 in ONE launch.  Returns the reference's traj_states
    (numpy): t [Ts,1], x/eps [Ts,1,N,D], P_sqrt [Ts,1,n,n], y_hat [Ts,1,L], S_sqrt [Ts,1,L,L], and the
    constant entries Q_sqrt, gamma_sqrt, R_sqrt, y broadcast along the saved-step axis like the scan's
    stacked output.  In guard mode "reference" P_sqrt is the factor itself with the reference's
    Householder signs; otherwise (and always for S_sqrt) the Cholesky factor of the full matrix."""
    x0 = _np(initial_state["x"])                       # [1, N, D]
    _, Nn, D = x0.shape
    n = Nn * D
    pk = plan_kwargs(filter_builder, solver_builder, ode_builder, use_static_cov_fn, state_shape=(Nn, D))
    H = _np(measurement_matrix)
    R_sqrt = _np(initial_state["R_sqrt"])
    L = R_sqrt.shape[-1]
    g = resolve_guard(guard, n)
    kw = dict(t0=float(_np(initial_state["t"]).reshape(-1)[0]), P0_sqrt=_np(initial_state["P_sqrt"]).reshape(n, n),
              theta_shared=flat_theta(params if params is not None else ode_builder.params),
              Q_sqrt=_np(initial_state["Q_sqrt"]).reshape(n, n), gamma_sqrt=float(_np(initial_state["gamma_sqrt"]).reshape(-1)[0]))
    if L > 0:
        kw.update(H=H.reshape(L, n), R_sqrt=R_sqrt.reshape(L, L), ys=_np(ys).reshape(-1, L),
                  correct_flags=_np(correct_flags, np.uint8), xy_index_map=_np(xy_index_map, np.int64))
    r = _run(pk, x0.reshape(1, n), num_steps, save_interval=1 if save_interval <= 0 else save_interval, guard=g, **kw)
    tr = r["traj"]
    Ts = tr["t"].shape[0]
    P = tr["P"].reshape(Ts, 1, n, n)
    out = {"t": tr["t"].reshape(Ts, 1), "x": tr["x"].reshape(Ts, 1, Nn, D), "eps": tr["eps"].reshape(Ts, 1, Nn, D),
           "P_sqrt": tr["P_sqrt"].reshape(Ts, 1, n, n) if "P_sqrt" in tr else _factor(P)}
    if L > 0:
        out["y_hat"] = tr["y_hat"].reshape(Ts, 1, L)
        out["S_sqrt"] = _factor(tr["S"].reshape(Ts, 1, L, L))
    else:
        out["y_hat"] = np.zeros((Ts, 1, 0))
        out["S_sqrt"] = np.zeros((Ts, 1, 0, 0))
    bc = lambda a: np.broadcast_to(_np(a)[None], (Ts,) + _np(a).shape).copy()
    out["Q_sqrt"], out["gamma_sqrt"], out["R_sqrt"] = bc(initial_state["Q_sqrt"]), bc(initial_state["gamma_sqrt"]), bc(R_sqrt)
    # state["y"]: the observation looked up at each step (run_filter.py:206); slot 0 = the initial zeros
    y = np.zeros((num_steps + 1, L))
    if L > 0:
        y[1:] = _np(ys).reshape(-1, L)[_np(xy_index_map, np.int64)[:num_steps]]
    out["y"] = y[::max(save_interval, 1)]
    return out


def nll(filter_builder, solver_builder, ode_builder, measurement_matrix, initial_state, ys, correct_flags,
        xy_index_map, num_steps: int, params: Dict[str, object], guard: str = "auto") -> float:
    """The scan + `nlls.sum()` of `nll()` (scripts/run_parameter_estimation.py:771-794) for ONE parameter
    set given as the reference's params dict (already de-normalised and scattered, :735-742)."""
    x0 = _np(initial_state["x"])
    _, Nn, D = x0.shape
    n = Nn * D
    pk = plan_kwargs(filter_builder, solver_builder, ode_builder, False, state_shape=(Nn, D))
    R_sqrt = _np(initial_state["R_sqrt"])
    L = R_sqrt.shape[-1]
    kw = dict(t0=float(_np(initial_state["t"]).reshape(-1)[0]), P0_sqrt=_np(initial_state["P_sqrt"]).reshape(n, n),
              theta_shared=flat_theta(params), Q_sqrt=_np(initial_state["Q_sqrt"]).reshape(n, n),
              gamma_sqrt=float(_np(initial_state["gamma_sqrt"]).reshape(-1)[0]), H=_np(measurement_matrix).reshape(L, n),
              R_sqrt=R_sqrt.reshape(L, L), ys=_np(ys).reshape(-1, L), correct_flags=_np(correct_flags, np.uint8),
              xy_index_map=_np(xy_index_map, np.int64))
    r = _run(pk, x0.reshape(1, n), num_steps, save_interval=0, guard=resolve_guard(guard, n), **kw)
    return float(r["nll"][0])


def _factor(P: np.ndarray) -> np.ndarray:
    """Lower-triangular factor of a PSD batch; exact zeros (slot 0 of S, P0 = 0) stay zeros."""
    out = np.zeros_like(P)
    flat, oflat = P.reshape(-1, P.shape[-2], P.shape[-1]), out.reshape(-1, P.shape[-2], P.shape[-1])
    for i, M in enumerate(flat):
        if M.size == 0 or not np.any(M):
            continue
        try:
            oflat[i] = np.linalg.cholesky(M)
        except np.linalg.LinAlgError:
            w, V = np.linalg.eigh(M)
            oflat[i] = V * np.sqrt(np.clip(w, 0.0, None))
    return out


# ------------------------------------------------------------------------------------------------
# nll_p of scripts/run_parameter_estimation.py:228-241 / :462-475 (the partial of nll() the optimiser
# and evaluate() call) as one fused launch, value and forward-mode gradient.
def _default_grad_runner():
    import torch

    from .engine import Plan, ekf_grad_run, param_sensitivity

    def run(pk, x0, T, grad_idx, *, x0_tangent=None, sensitivity=False, **kw):
        if not torch.cuda.is_available():
            raise RuntimeError("B200_SQRT_EKF needs a CUDA device: the EKF-RK path has no CPU fallback")
        dev = torch.device("cuda")
        plan = Plan(**pk)
        t = lambda a, dt=torch.float64: None if a is None else torch.as_tensor(np.asarray(a), dtype=dt).to(dev)
        x0_d, th_d, x0t_d = t(x0), t(kw["theta"]), t(x0_tangent)
        qd = qdt = None
        Q_sqrt = kw.get("Q_sqrt")
        if sensitivity:                       # Q_sqrt = diag(w(theta)), run_parameter_estimation.py:750-769
            qd, qdt = param_sensitivity(plan, x0_d, grad_idx, t0=kw["t0"], theta=th_d, x0_tangent=x0t_d)
            Q_sqrt = None
        nll, g = ekf_grad_run(plan, x0_d, int(T), grad_idx, t0=kw["t0"], P0_sqrt=kw["P0_sqrt"], theta=th_d,
                              Q_sqrt=Q_sqrt, gamma_sqrt=kw["gamma_sqrt"], H=kw["H"], R_sqrt=kw["R_sqrt"], ys=t(kw["ys"]),
                              correct_flags=t(kw["correct_flags"], torch.uint8), xy_index_map=t(kw["xy_index_map"], torch.int64),
                              x0_tangent=x0t_d, Q_sqrt_diag=qd, Q_sqrt_diag_tangent=qdt)
        return nll.cpu().numpy(), g.cpu().numpy()

    return run


_grad_runner: Optional[Callable] = None


def set_grad_runner(fn: Optional[Callable]) -> None:
    """Test hook, see set_runner."""
    global _grad_runner
    _grad_runner = fn


class NllP:
    """Callable with the signature of the reference's `nll_p` (what `optimize_run` hands to
    ScipyBoundedMinimize, :599, and what `evaluate` times, :496-522):

        nll_p(params_norm_reduced, initial_state, x0, H, ys, correct_flags, xy_index_map,
              params_min_reduced, params_max_reduced, params_optimized, params_optimized_indices,
              params_default) -> NLL

    `value_and_grad(...)` returns (NLL, d NLL / d params_norm_reduced) with the gradient as a dict
    shaped like `params_norm_reduced` - what jaxopt expects with `value_and_grad=True`.  The gradient
    is the kernel's forward-mode derivative w.r.t. the physical parameters times (max - min)
    (inv_normalize, src/utils.py:156-178)."""

    def __init__(self, filter_builder, solver_builder, ode_builder, num_steps: int,
                 initial_state_parametrized: bool, parameter_sensitivity: bool) -> None:
        self.fb, self.sb, self.ob = filter_builder, solver_builder, ode_builder
        self.T, self.isp, self.ps = int(num_steps), bool(initial_state_parametrized), bool(parameter_sensitivity)

    def _evaluate(self, params_norm, initial_state, x0, H, ys, correct_flags, xy_index_map, params_min, params_max,
                  params_optimized, params_optimized_indices, params_default):
        from .runners import initial_value_and_tangent, param_layout
        keys_s, sizes, perm = param_layout(self.ob)                       # JAX flattens dicts in sorted-key order
        red = sorted(params_norm)                                         # the optimised keys
        pn = np.concatenate([_np(params_norm[k]).reshape(-1) for k in red])
        lo = np.concatenate([np.broadcast_to(_np(params_min[k]).reshape(-1), (sizes[k],)) for k in red])
        hi = np.concatenate([np.broadcast_to(_np(params_max[k]).reshape(-1), (sizes[k],)) for k in red])
        opt_idx = _np(params_optimized_indices, np.int64).reshape(-1)
        flat = np.concatenate([_np(params_default[k]).reshape(-1) for k in keys_s])
        flat[opt_idx] = pn * (hi - lo) + lo                               # :735-742
        theta = flat[perm][None, :]
        gidx = np.argsort(perm)[opt_idx].astype(np.int32)                 # builder positions of the optimised entries
        xs = _np(initial_state["x"])
        _, Nn, D = xs.shape
        n = Nn * D
        x0_b, x0_tan = xs.reshape(1, n), None
        if self.isp:                                                      # :744-748
            x0_b, x0_tan = initial_value_and_tangent(self.ob, _np(x0), flat[None, :], opt_idx)
        R_sqrt = _np(initial_state["R_sqrt"])
        L = R_sqrt.shape[-1]
        pk = plan_kwargs(self.fb, self.sb, self.ob, False, state_shape=(Nn, D))
        kw = dict(t0=float(_np(initial_state["t"]).reshape(-1)[0]), P0_sqrt=_np(initial_state["P_sqrt"]).reshape(n, n),
                  theta=theta, Q_sqrt=_np(initial_state["Q_sqrt"]).reshape(n, n),
                  gamma_sqrt=float(_np(initial_state["gamma_sqrt"]).reshape(-1)[0]), H=_np(H).reshape(L, n), R_sqrt=R_sqrt.reshape(L, L),
                  ys=_np(ys).reshape(-1, L), correct_flags=_np(correct_flags, np.uint8),
                  xy_index_map=_np(xy_index_map, np.int64))
        nll, g = (_grad_runner or _default_grad_runner())(pk, x0_b, self.T, gidx, x0_tangent=x0_tan,
                                                          sensitivity=self.ps, **kw)
        gn = g[0] * (hi - lo)
        grad, o = {}, 0
        for k in red:
            grad[k] = gn[o:o + sizes[k]].reshape(_np(params_norm[k]).shape)
            o += sizes[k]
        return float(nll[0]), grad

    _NAMES = ("initial_state", "x0", "measurement_matrix", "ys", "correct_flags", "xy_index_map", "params_min",
              "params_max", "params_optimized", "params_optimized_indices", "params_default")

    def _args(self, args, kwargs):
        # jaxopt calls fun(params, **kwargs) with the keyword names of lbfgsb.run(...) (:626-640);
        # evaluate() passes everything positionally (:497-510)
        return tuple(args) + tuple(kwargs[k] for k in self._NAMES[len(args) - 1:])

    def __call__(self, *args, **kwargs) -> float:
        return self._evaluate(*self._args(args, kwargs))[0]

    def value_and_grad(self, *args, **kwargs):
        return self._evaluate(*self._args(args, kwargs))
