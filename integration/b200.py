"""Reference-side binding: drop this file into the reference tree as `src/filters/b200.py`.

    filter_builder:
      class_path: src.filters.b200.B200_SQRT_EKF        # was: src.filters.SQRT_EKF
      init_args: {cov_update_fn_builder: ..., disable_cov_update: ...}   # unchanged

`B200_SQRT_EKF` subclasses the reference's own `SQRT_EKF` (src/filters/sqrt_ekf.py:16-376), so
jsonargparse's type check against `FilterBuilder` (scripts/run_filter.py:33,
scripts/run_parameter_estimation.py:51) accepts it and every single-step method (`init_state`,
`build_predict`, `build_correct`, `build_cov_update_fn`) keeps working on JAX.  It adds the two
whole-trajectory hooks the fused kernel needs - the reference keeps its time loop in the scripts
(`lax.scan` in `unroll`, scripts/run_filter.py:166-224, and in `nll`,
scripts/run_parameter_estimation.py:771-794), so a single kernel launch cannot be reached through
`build_predict()` alone:

    build_unroll(solver_builder, ode_builder, use_static_cov_fn) -> unroll_fn
        unroll_fn(H, initial_state, ys, correct_flags, xy_index_map, num_steps, save_interval)
        returns the same `traj_states` dict as `unroll()` (same keys, shapes and dtypes).
    build_nll(solver_builder, ode_builder) -> nll_fn(params, initial_state, H, ys, correct_flags,
        xy_index_map, num_steps) -> NLL for one (already scattered) parameter dict.

The scripts prefer the hooks when the filter builder has them: `integration/run_filter.patch`
(4 lines at scripts/run_filter.py:148-161) and `integration/run_parameter_estimation.patch`.

The marshalling (reference state dict <-> `odeu_ekf_io`, include/odeu.h) lives in
`ode_uncertainty_b200/reference_binding.py`; the launch goes ctypes -> C ABI -> CUDA kernels on the
current device.  There is no CPU fallback: without `libodeu.so` or a GPU the hooks raise.
"""
from typing import Callable, Dict

from jax import Array
from jax import numpy as jnp

from src.covariance_update_functions import DiagonalCovarianceUpdate, StaticDiagonalCovarianceUpdate
from src.covariance_update_functions.covariance_update_function import CovarianceUpdateFunctionBuilder
from src.covariance_update_functions.static_covariance_update_function import (
    StaticCovarianceUpdateFunctionBuilder,
)
from src.filters.sqrt_ekf import SQRT_EKF
from src.ode.ode import ODEBuilder
from src.solvers.solver import SolverBuilder

from ode_uncertainty_b200 import reference_binding as rb


class B200_SQRT_EKF(SQRT_EKF):
    """SQRT_EKF whose whole-trajectory loop runs as one fused CUDA kernel launch on a B200."""

    def __init__(
        self,
        cov_update_fn_builder: CovarianceUpdateFunctionBuilder = DiagonalCovarianceUpdate(),
        static_cov_update_fn_builder: StaticCovarianceUpdateFunctionBuilder = StaticDiagonalCovarianceUpdate(),
        disable_cov_update: bool = False,
        guard: str = "auto",
    ) -> None:
        """
        Args as SQRT_EKF (src/filters/sqrt_ekf.py:19-43), plus
            guard (str): zero-gain guard of the correct step (sqrt_ekf.py:350-353).
                "reference": `all(S_sqrt < 1e-16)` exactly as written, on a factor with LAPACK's
                    Householder signs (factor-form kernels, state dimension <= 4);
                "intended": `all(|S_sqrt| < 1e-16)` (full-covariance kernels, any dimension);
                "auto" (default): "reference" where it is served, else "intended".
        """
        super().__init__(cov_update_fn_builder, static_cov_update_fn_builder, disable_cov_update)
        self.guard = guard

    def build_unroll(
        self, solver_builder: SolverBuilder, ode_builder: ODEBuilder, use_static_cov_fn: bool = False
    ) -> Callable[..., Dict[str, Array]]:
        def unroll_fn(
            measurement_matrix: Array,
            initial_state: Dict[str, Array],
            ys: Array,
            correct_flags: Array,
            xy_index_map: Array,
            num_steps: int,
            save_interval: int,
        ) -> Dict[str, Array]:
            out = rb.unroll(
                self,
                solver_builder,
                ode_builder,
                use_static_cov_fn,
                measurement_matrix,
                initial_state,
                ys,
                correct_flags,
                xy_index_map,
                num_steps,
                save_interval,
                guard=self.guard,
            )
            # same keys as the scan output of unroll() (diffrax_state is deleted there, :219)
            return {k: jnp.asarray(out[k]) for k in initial_state if k != "diffrax_state"}

        return unroll_fn

    def build_nll(self, solver_builder: SolverBuilder, ode_builder: ODEBuilder) -> Callable[..., Array]:
        def nll_fn(
            params: Dict[str, Array],
            initial_state: Dict[str, Array],
            measurement_matrix: Array,
            ys: Array,
            correct_flags: Array,
            xy_index_map: Array,
            num_steps: int,
        ) -> Array:
            return jnp.asarray(
                rb.nll(
                    self,
                    solver_builder,
                    ode_builder,
                    measurement_matrix,
                    initial_state,
                    ys,
                    correct_flags,
                    xy_index_map,
                    num_steps,
                    params,
                    guard=self.guard,
                )
            )

        return nll_fn

    def build_nll_p(
        self,
        num_steps: int,
        initial_state_parametrized: bool,
        parameter_sensitivity: bool,
        solver_builder: SolverBuilder,
        ode_builder: ODEBuilder,
    ) -> "rb.NllP":
        """Replacement of the jitted partial `nll_p` (scripts/run_parameter_estimation.py:228-241,
        :462-475): same call signature; `.value_and_grad` serves ScipyBoundedMinimize(value_and_grad=True)."""
        return _JaxNllP(
            rb.NllP(self, solver_builder, ode_builder, num_steps, initial_state_parametrized, parameter_sensitivity)
        )


class _JaxNllP:
    """`rb.NllP` with jax.Array results (value: Array []; gradient: dict of Arrays)."""

    def __init__(self, impl: "rb.NllP") -> None:
        self.impl = impl

    def __call__(self, *args, **kwargs) -> Array:
        return jnp.asarray(self.impl(*args, **kwargs))

    def value_and_grad(self, *args, **kwargs):
        value, grad = self.impl.value_and_grad(*args, **kwargs)
        return jnp.asarray(value), {k: jnp.asarray(v) for k, v in grad.items()}
