"""Covariance-update plugins (eps -> Q): mirror of `src/covariance_update_functions/`.

The mappings are fused into the EKF kernel (csrc/ekf_core.cuh::add_process_noise), so the
builders only carry the plugin identity and `scale` into the plan; `build()` / `build_sqrt()`
return a descriptor the filter builders consume (calling it on arrays raises: there is no host
implementation)."""
from __future__ import annotations

from . import _native as N


class CovarianceUpdateFunction:
    """Descriptor standing in for the reference's `(cov, eps) -> cov` callable
    (src/covariance_update_functions/covariance_update_function.py:5)."""

    def __init__(self, cov_fn_id: int, scale: float, sqrt: bool, static: bool = False) -> None:
        self.cov_fn_id, self.scale, self.sqrt, self.static = cov_fn_id, float(scale), sqrt, static

    def __call__(self, *args, **kwargs):
        raise NotImplementedError(
            "covariance-update functions are fused into the CUDA EKF kernel; pass the builder to "
            "SQRT_EKF / ParticleFilter instead of calling it on host arrays")


class CovarianceUpdateFunctionBuilder:
    """covariance_update_function.py:8-35."""
    cov_fn_id = -1

    def __init__(self, scale: float = 1.0) -> None:
        self.scale = scale

    def build(self) -> CovarianceUpdateFunction:
        if self.cov_fn_id < 0:
            raise NotImplementedError
        return CovarianceUpdateFunction(self.cov_fn_id, self.scale, sqrt=False)

    def build_sqrt(self) -> CovarianceUpdateFunction:
        if self.cov_fn_id < 0:
            raise NotImplementedError
        return CovarianceUpdateFunction(self.cov_fn_id, self.scale, sqrt=True)


class DiagonalCovarianceUpdate(CovarianceUpdateFunctionBuilder):
    """P += diag((scale * eps)^2)  (diagonal.py:11-58)."""
    cov_fn_id = N.COV_DIAGONAL


class OuterCovarianceUpdate(CovarianceUpdateFunctionBuilder):
    """P += (scale * eps)(scale * eps)^T  (outer.py:11-62)."""
    cov_fn_id = N.COV_OUTER


class StaticCovarianceUpdateFunctionBuilder:
    """static_covariance_update_function.py:9-46."""
    cov_fn_id = -1

    def __init__(self, scale: float = 1.0) -> None:
        self.scale = scale

    def build(self) -> CovarianceUpdateFunction:
        if self.cov_fn_id < 0:
            raise NotImplementedError
        return CovarianceUpdateFunction(self.cov_fn_id, self.scale, sqrt=False, static=True)

    def build_sqrt(self) -> CovarianceUpdateFunction:
        if self.cov_fn_id < 0:
            raise NotImplementedError
        return CovarianceUpdateFunction(self.cov_fn_id, self.scale, sqrt=True, static=True)


class StaticDiagonalCovarianceUpdate(StaticCovarianceUpdateFunctionBuilder):
    """P += scale^2 I  (static_diagonal.py:11-48)."""
    cov_fn_id = N.COV_STATIC_DIAGONAL
