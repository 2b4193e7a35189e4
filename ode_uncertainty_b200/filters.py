"""Filter plugins: mirror of `src/filters/{filter,sqrt_ekf,particle_filter}.py`.

Same class names and `init_args` as the reference (`cov_update_fn_builder`,
`static_cov_update_fn_builder`, `disable_cov_update`, `num_particles`), the same state
dictionaries, and the same single-step callables for API compatibility:

    FilterPredict   (solver, cov_update_fn, state) -> state          src/filters/filter.py:22-24
    FilterCorrect   (H, state) -> state                              src/filters/filter.py:31-33

Both run ONE step of the fused CUDA kernel.  The production entry points are the
whole-trajectory hooks `build_unroll()` / `build_nll()` (the reference keeps its time loop in
the scripts), implemented in runners.py.

State convention: the reference stores a square-root factor `P_sqrt` that is defined only up to
column signs; here `state["P_sqrt"]` is accepted as ANY factor (`P = P_sqrt P_sqrt^T` is what
enters the kernel) and returned as the Cholesky factor of the updated covariance.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _native as N
from .covariance_update_functions import (CovarianceUpdateFunction, CovarianceUpdateFunctionBuilder,
                                          DiagonalCovarianceUpdate, StaticCovarianceUpdateFunctionBuilder,
                                          StaticDiagonalCovarianceUpdate)
from .engine import Plan, ekf_run, pf_run
from .solvers import RKSolverBuilder


def _f64(a, device=None):
    t = torch.as_tensor(np.asarray(a.detach().cpu()) if isinstance(a, torch.Tensor) else np.asarray(a),
                        dtype=torch.float64)
    return t if device is None else t.to(device)


def _factor(P: torch.Tensor) -> torch.Tensor:
    """A lower-triangular factor of a PSD matrix batch; exact zeros stay zeros (P0 = 0 is legal)."""
    n = P.shape[-1]
    jitter = torch.zeros_like(P)
    L, info = torch.linalg.cholesky_ex(P + jitter)
    bad = info != 0
    if bool(bad.any()):
        # semi-definite (e.g. exact zero covariance): eigen-factor, still P = F F^T
        w, V = torch.linalg.eigh(P)
        F = V * torch.sqrt(torch.clamp(w, min=0.0)).unsqueeze(-2)
        L = torch.where(bad.reshape(bad.shape + (1, 1)), F, L)
    return L


class FilterBuilder:
    """src/filters/filter.py:36-133."""

    def __init__(self,
                 cov_update_fn_builder: CovarianceUpdateFunctionBuilder = DiagonalCovarianceUpdate(),
                 static_cov_update_fn_builder: StaticCovarianceUpdateFunctionBuilder = StaticDiagonalCovarianceUpdate(),
                 ) -> None:
        self.cov_update_fn_builder = cov_update_fn_builder
        self.static_cov_update_fn_builder = static_cov_update_fn_builder

    def init_state(self, solver_state: Dict[str, torch.Tensor], *args) -> Dict[str, torch.Tensor]:
        return dict(solver_state)

    def build_cov_update_fn(self) -> CovarianceUpdateFunction:
        raise NotImplementedError

    def build_static_cov_update_fn(self) -> CovarianceUpdateFunction:
        raise NotImplementedError

    def build_predict(self):
        raise NotImplementedError

    def build_parametrized_predict(self):
        predict = self.build_predict()

        def parametrized_predict(solver, cov_update_fn, ode, params, state):
            return predict(_BoundSolver(solver, ode, params), cov_update_fn, state)

        return parametrized_predict

    def build_correct(self):
        raise NotImplementedError   # like the reference (filter.py:122-133)

    # -- shared helper: plan for (solver, cov fn)
    def _plan(self, solver, cov_update_fn: Optional[CovarianceUpdateFunction], disable: bool = False) -> Plan:
        sb, ode, _ = _unwrap_solver(solver)
        b = ode.builder
        cov_id = N.COV_DIAGONAL if cov_update_fn is None else cov_update_fn.cov_fn_id
        scale = 1.0 if cov_update_fn is None else cov_update_fn.scale
        return Plan(b.ode_id, sb.solver_id, sb.h, ode_variant=b.ode_variant,
                    num_compartments=b.num_compartments_abi, cov_fn_id=cov_id, cov_scale=scale,
                    disable_cov_update=disable)


class _BoundSolver:
    """`partial(solver, ode, params)` of filter.py:117-118, keeping the pieces inspectable."""

    def __init__(self, parametrized_solver, ode, params):
        self.parametrized_solver, self.ode, self.params = parametrized_solver, ode, params

    def __call__(self, state):
        return self.parametrized_solver(self.ode, self.params, state)


class _SolverHandle:
    """What `SolverBuilder.build()` hands to a filter here: the builder itself, callable."""

    def __init__(self, builder: RKSolverBuilder):
        self.builder = builder
        self._step = builder.build()

    def __call__(self, state):
        return self._step(state)


def solver_handle(builder: RKSolverBuilder) -> _SolverHandle:
    """Replacement of `jax.jit(jax.vmap(solver_builder.build()))` (scripts/run_filter.py:83)."""
    if not hasattr(builder, "ode") or not hasattr(builder, "params"):
        raise AttributeError("Setup solver before usage!")
    return _SolverHandle(builder)


def _unwrap_solver(solver):
    if isinstance(solver, _SolverHandle):
        return solver.builder, solver.builder.ode, solver.builder.params
    if isinstance(solver, _BoundSolver):
        ps = solver.parametrized_solver
        sb = getattr(ps, "builder", None)
        if sb is None:
            raise TypeError("parametrized solver must come from parametrized_solver_handle()")
        return sb, solver.ode, solver.params
    raise TypeError("solver must come from solver_handle() / parametrized_solver_handle()")


class _ParamSolverHandle:
    def __init__(self, builder: RKSolverBuilder):
        self.builder = builder
        self._step = builder.build_parametrized()

    def __call__(self, ode, params, state):
        return self._step(ode, params, state)


def parametrized_solver_handle(builder: RKSolverBuilder) -> _ParamSolverHandle:
    """Replacement of `jax.vmap(solver_builder.build_parametrized(), (None, None, 0))`
    (scripts/run_parameter_estimation.py:222-224)."""
    return _ParamSolverHandle(builder)


class SQRT_EKF(FilterBuilder):
    """Square-root EKF plugin (src/filters/sqrt_ekf.py:33-376), computed in full-covariance form."""

    def __init__(self,
                 cov_update_fn_builder: CovarianceUpdateFunctionBuilder = DiagonalCovarianceUpdate(),
                 static_cov_update_fn_builder: StaticCovarianceUpdateFunctionBuilder = StaticDiagonalCovarianceUpdate(),
                 disable_cov_update: bool = False) -> None:
        super().__init__(cov_update_fn_builder, static_cov_update_fn_builder)
        self.disable_cov_update = disable_cov_update

    def init_state(self, solver_state, P0_sqrt, Q_sqrt, gamma_sqrt, R_sqrt) -> Dict[str, torch.Tensor]:
        """sqrt_ekf.py:45-84: adds the leading singleton batch axis and the filter entries."""
        st = dict(solver_state)
        R_sqrt = _f64(R_sqrt)
        L = R_sqrt.shape[-1]
        st["t"] = _f64(st["t"]).reshape(1)
        st["x"] = _f64(st["x"])[None]
        st["eps"] = _f64(st["eps"])[None]
        st["diffrax_state"] = torch.zeros(1)
        st["P_sqrt"] = _f64(P0_sqrt)[None]
        st["Q_sqrt"] = _f64(Q_sqrt)
        st["gamma_sqrt"] = _f64(gamma_sqrt)
        st["y"] = torch.zeros(L, dtype=torch.float64)
        st["y_hat"] = torch.zeros(1, L, dtype=torch.float64)
        st["R_sqrt"] = R_sqrt
        st["S_sqrt"] = torch.zeros(1, L, L, dtype=torch.float64)
        return st

    def build_cov_update_fn(self) -> CovarianceUpdateFunction:
        return self.cov_update_fn_builder.build_sqrt()

    def build_static_cov_update_fn(self) -> CovarianceUpdateFunction:
        return self.static_cov_update_fn_builder.build_sqrt()

    def build_predict(self):
        def predict(solver, cov_update_fn_sqrt, state, device="cuda"):
            """One predict step (sqrt_ekf.py:92-197) on the GPU; state tensors may live anywhere."""
            plan = self._plan(solver, cov_update_fn_sqrt, self.disable_cov_update)
            _, ode, params = _unwrap_solver(solver)
            dev = torch.device(device)
            x = _f64(state["x"], dev)
            M = x.shape[0]
            Ps = _f64(state["P_sqrt"], dev)
            P = Ps @ Ps.transpose(-1, -2)
            r = ekf_run(plan, x.reshape(M, -1), 1, t0=float(_f64(state["t"]).reshape(-1)[0]), P0=P,
                        theta_shared=ode.builder.flat_params(params), Q_sqrt=_f64(state["Q_sqrt"]),
                        gamma_sqrt=float(_f64(state["gamma_sqrt"])))
            out = dict(state)
            out.update(t=r.tT.reshape(1), x=r.xT.reshape(x.shape), eps=r.epsT.reshape(x.shape),
                       P_sqrt=_factor(r.PT), P=r.PT)
            return out
        return predict

    def build_correct(self):
        def correct(H, state, device="cuda"):
            """One measurement update (sqrt_ekf.py:337-376).  The ODE/solver are irrelevant to it;
            a Lorenz-shaped plan of matching dimension is not needed: `state` carries the plan
            context recorded by predict(), else the dimension picks a neutral plugin."""
            dev = torch.device(device)
            x = _f64(state["x"], dev)
            M = x.shape[0]
            n = x[0].numel()
            plan = _neutral_plan(n)
            Ps = _f64(state["P_sqrt"], dev)
            P = Ps @ Ps.transpose(-1, -2)
            Hh = _f64(H)
            L = Hh.shape[0]
            y = _f64(state["y"], dev).reshape(1, L)
            r = ekf_run(plan, x.reshape(M, -1), 1, t0=float(_f64(state["t"]).reshape(-1)[0]), P0=P,
                        H=Hh, R_sqrt=_f64(state["R_sqrt"]), ys=y,
                        correct_flags=torch.ones(1, dtype=torch.uint8, device=dev),
                        xy_index_map=torch.zeros(1, dtype=torch.int64, device=dev), skip_predict=True)
            out = dict(state)
            out.update(x=r.xT.reshape(x.shape), P_sqrt=_factor(r.PT), P=r.PT, y_hat=r.yhatT,
                       S_sqrt=_factor(r.ST), S=r.ST, nlg=r.nll)
            return out
        return correct

    # ---- whole-trajectory hooks (what the patched scripts call instead of lax.scan)
    def build_unroll(self):
        from .runners import ekf_unroll
        return ekf_unroll

    def build_nll(self):
        from .runners import ekf_nll
        return ekf_nll


_NEUTRAL = {2: N.ODE_VAN_DER_POL, 3: N.ODE_LORENZ, 4: (N.ODE_LCAO, 2, 0), 7: (N.ODE_HODGKIN_HUXLEY, 1, 0),
            8: (N.ODE_HODGKIN_HUXLEY, 0, 0), 14: (N.ODE_MULTI_HH, 1, 2)}


def _neutral_plan(n: int) -> Plan:
    """A plan of state dimension n for the ODE-independent measurement update."""
    if n not in _NEUTRAL:
        raise ValueError(f"no plugin of state dimension {n} for a stand-alone correct step")
    e = _NEUTRAL[n]
    ode_id, variant, nc = e if isinstance(e, tuple) else (e, 0, 0)
    return Plan(ode_id, N.SOLVER_RKF45, 1.0, ode_variant=variant, num_compartments=nc)


class ParticleFilter(FilterBuilder):
    """Perturbed-solver particle ensemble (src/filters/particle_filter.py:24-118): predict only,
    no weights, no resampling, `build_correct` not implemented - exactly like the reference."""

    def __init__(self,
                 cov_update_fn_builder: CovarianceUpdateFunctionBuilder = DiagonalCovarianceUpdate(),
                 static_cov_update_fn_builder: StaticCovarianceUpdateFunctionBuilder = StaticDiagonalCovarianceUpdate(),
                 num_particles: int = 100) -> None:
        super().__init__(cov_update_fn_builder, static_cov_update_fn_builder)
        self.M = num_particles

    def init_state(self, solver_state, prng_key) -> Dict[str, torch.Tensor]:
        """particle_filter.py:36-65; `prng_key` is an integer seed here (Philox, not threefry)."""
        st = dict(solver_state)
        st["t"] = _f64(st["t"]).reshape(1).expand(self.M).clone()
        st["x"] = _f64(st["x"])[None].expand((self.M,) + tuple(st["x"].shape)).clone()
        st["eps"] = torch.zeros_like(st["x"])
        st["diffrax_state"] = torch.zeros(self.M)
        st["prng_key"] = int(prng_key)
        st["step"] = 0
        return st

    def build_cov_update_fn(self) -> CovarianceUpdateFunction:
        return self.cov_update_fn_builder.build()

    def build_static_cov_update_fn(self) -> CovarianceUpdateFunction:
        return self.static_cov_update_fn_builder.build()

    def build_predict(self):
        def predict(solver, cov_update_fn, state, device="cuda"):
            plan = self._plan(solver, cov_update_fn)
            _, ode, params = _unwrap_solver(solver)
            dev = torch.device(device)
            x = _f64(state["x"], dev)
            M = x.shape[0]
            r = pf_run(plan, M, 1, x0=x.reshape(M, -1), t0=float(_f64(state["t"]).reshape(-1)[0]),
                       theta_shared=ode.builder.flat_params(params), seed=state["prng_key"],
                       step_offset=state.get("step", 0))
            out = dict(state)
            out.update(t=r.tT.reshape(1).expand(M).clone(), x=r.xT.reshape(x.shape),
                       eps=r.epsT.reshape(x.shape), step=state.get("step", 0) + 1)
            return out
        return predict

    def build_unroll(self):
        from .runners import pf_unroll
        return pf_unroll
