"""ctypes binding of the C ABI in include/odeu.h.

There is no CPU fallback: if the CUDA library is missing or a symbol is absent this module
raises at import of the first symbol, and every run call requires CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ODEU_LIB", os.path.join(_HERE, "libodeu.so"))   # ODEU_LIB: A/B builds (dev)

# enums of include/odeu.h
ODE_LORENZ, ODE_VAN_DER_POL, ODE_LOTKA_VOLTERRA, ODE_PENDULUM, ODE_LCAO, ODE_HODGKIN_HUXLEY, ODE_MULTI_HH = range(7)
SOLVER_RKF45, SOLVER_DOPRI65, SOLVER_BS32, SOLVER_HEUN_EULER, SOLVER_KVAERNO3, SOLVER_IMPLICIT_EULER = range(6)
COV_DIAGONAL, COV_OUTER, COV_STATIC_DIAGONAL = range(3)
GUARD_INTENDED, GUARD_REFERENCE, GUARD_INTENDED_FACTOR = range(3)

# every symbol include/odeu.h declares (tests assert the library exports all of them)
SYMBOLS = (
    "odeu_version", "odeu_plan_create", "odeu_plan_destroy", "odeu_plan_state_dim",
    "odeu_plan_num_params", "odeu_plan_default_params", "odeu_ekf_run", "odeu_pf_run",
    "odeu_launch_count", "odeu_last_error", "odeu_bench_dfma", "odeu_ode_rhs",
    "odeu_ekf_grad_run", "odeu_ekf_workspace_bytes", "odeu_pf_weight_update",
    "odeu_ekf_dense_run", "odeu_ekf_dense_workspace_bytes", "odeu_param_sensitivity",
    "odeu_pf_reduce_scratch_bytes", "odeu_pf_weight_reduce", "odeu_pf_normalize", "odeu_pf_resample",
    "odeu_pf_scan_bytes", "odeu_pf_scan_resample",
    "odeu_pf_publish_triple", "odeu_pf_normalize_w", "odeu_pf_scan_resample_peer",
    "odeu_lbfgs_workspace_doubles", "odeu_lbfgs_step",
)


class PlanDesc(C.Structure):
    _fields_ = [
        ("ode_id", C.c_int32), ("ode_variant", C.c_int32), ("num_compartments", C.c_int32),
        ("solver_id", C.c_int32), ("step_size", C.c_double), ("cov_fn_id", C.c_int32),
        ("cov_scale", C.c_double), ("disable_cov_update", C.c_int32),
    ]


_dp = C.c_void_p  # all data pointers are passed as raw addresses


class EkfIO(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("T", C.c_int64), ("t0", C.c_double), ("L", C.c_int32),
        ("x0", _dp), ("P0", _dp), ("P0_sqrt", _dp), ("theta", _dp), ("theta_shared", _dp),
        ("Q_sqrt", _dp), ("gamma_sqrt", C.c_double), ("H", _dp), ("R_sqrt", _dp), ("ys", _dp),
        ("ys_per_trajectory", C.c_int32), ("correct_flags", _dp), ("xy_index_map", _dp),
        ("save_interval", C.c_int64), ("skip_predict", C.c_int32), ("workspace", _dp),
        ("workspace_bytes", C.c_int64),
        ("xT", _dp), ("epsT", _dp), ("PT", _dp), ("yhatT", _dp), ("ST", _dp), ("nll", _dp),
        ("tT", _dp), ("out_t", _dp), ("out_x", _dp), ("out_eps", _dp), ("out_P", _dp),
        ("out_yhat", _dp), ("out_S", _dp), ("cov_scale_batch", _dp), ("nll_nan_to_num", C.c_int32),
        ("Q_sqrt_diag_batch", _dp),
        ("guard_mode", C.c_int32), ("P0_sqrt_batch", _dp), ("PT_sqrt", _dp), ("out_P_sqrt", _dp),
        ("guard_counts", _dp),
    ]


class GradIO(C.Structure):
    _fields_ = [("p_opt", C.c_int32), ("idx", _dp), ("x0_tangent", _dp), ("grad", _dp),
                ("Q_sqrt_diag_tangent", _dp)]


class SensIO(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("t0", C.c_double), ("x0", _dp), ("theta", _dp), ("theta_shared", _dp),
        ("p_opt", C.c_int32), ("idx", _dp), ("x0_tangent", _dp), ("w", _dp), ("w_tangent", _dp),
    ]


class DenseIO(C.Structure):
    _fields_ = [
        ("B", C.c_int64), ("T", C.c_int64), ("t0", C.c_double), ("L", C.c_int32),
        ("x0", _dp), ("x", _dp), ("P", _dp), ("P0_sqrt", _dp), ("theta_shared", _dp),
        ("Q_sqrt_diag", _dp), ("gamma_sqrt", C.c_double), ("obs_index", _dp), ("R_sqrt", _dp),
        ("ys", _dp), ("ys_per_trajectory", C.c_int32), ("correct_flags", _dp), ("xy_index_map", _dp),
        ("eps", _dp), ("nll", _dp), ("tT", _dp), ("workspace", _dp), ("workspace_bytes", C.c_int64),
    ]


class PfIO(C.Structure):
    _fields_ = [
        ("M", C.c_int64), ("T", C.c_int64), ("t0", C.c_double),
        ("x0", _dp), ("x0_shared", _dp), ("theta_shared", _dp), ("seed", C.c_uint64), ("particle_offset", C.c_int64),
        ("step_offset", C.c_int64), ("save_interval", C.c_int64), ("noise_free", C.c_int32),
        ("xT", _dp), ("epsT", _dp), ("tT", _dp), ("out_t", _dp), ("out_x", _dp), ("out_eps", _dp),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make` or `python -c 'import __graft_entry__ as g; "
            "g.build()'`. There is no CPU fallback for the EKF-RK path.")
    L = C.CDLL(LIB_PATH)
    L.odeu_version.restype = C.c_int
    L.odeu_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    L.odeu_plan_create.restype = C.c_int
    L.odeu_plan_destroy.argtypes = [C.c_void_p]
    L.odeu_plan_destroy.restype = None
    L.odeu_plan_state_dim.argtypes = [C.c_void_p]
    L.odeu_plan_state_dim.restype = C.c_int
    L.odeu_plan_num_params.argtypes = [C.c_void_p]
    L.odeu_plan_num_params.restype = C.c_int
    L.odeu_plan_default_params.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.odeu_plan_default_params.restype = C.c_int
    L.odeu_ekf_run.argtypes = [C.c_void_p, C.POINTER(EkfIO), C.c_void_p]
    L.odeu_ekf_run.restype = C.c_int
    L.odeu_pf_run.argtypes = [C.c_void_p, C.POINTER(PfIO), C.c_void_p]
    L.odeu_pf_run.restype = C.c_int
    L.odeu_bench_dfma.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                  C.POINTER(C.c_double), C.c_void_p]
    L.odeu_bench_dfma.restype = C.c_int
    L.odeu_pf_weight_update.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.odeu_pf_weight_update.restype = C.c_int
    L.odeu_pf_reduce_scratch_bytes.argtypes = [C.c_int64]
    L.odeu_pf_reduce_scratch_bytes.restype = C.c_int64
    L.odeu_pf_weight_reduce.argtypes = [C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 8
    L.odeu_pf_weight_reduce.restype = C.c_int
    L.odeu_pf_normalize.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_double, C.c_void_p]
    L.odeu_pf_normalize.restype = C.c_int
    L.odeu_pf_resample.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_double] + [C.c_void_p] * 6
    L.odeu_pf_resample.restype = C.c_int
    L.odeu_pf_scan_bytes.argtypes = [C.c_int64]
    L.odeu_pf_scan_bytes.restype = C.c_int64
    L.odeu_pf_scan_resample.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_double] + [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    L.odeu_pf_scan_resample.restype = C.c_int
    L.odeu_pf_publish_triple.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.odeu_pf_publish_triple.restype = C.c_int
    L.odeu_pf_normalize_w.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 8 + [C.c_double, C.c_void_p]
    L.odeu_pf_normalize_w.restype = C.c_int
    L.odeu_pf_scan_resample_peer.argtypes = ([C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double] + [C.c_void_p] * 7
                                             + [C.c_int64, C.c_void_p])
    L.odeu_pf_scan_resample_peer.restype = C.c_int
    L.odeu_lbfgs_workspace_doubles.argtypes = [C.c_int32, C.c_int32]
    L.odeu_lbfgs_workspace_doubles.restype = C.c_int64
    L.odeu_lbfgs_step.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double] + [C.c_void_p] * 6
    L.odeu_lbfgs_step.restype = C.c_int
    L.odeu_ekf_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
    L.odeu_ekf_workspace_bytes.restype = C.c_int64
    L.odeu_ekf_dense_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    L.odeu_ekf_dense_workspace_bytes.restype = C.c_int64
    L.odeu_ekf_dense_run.argtypes = [C.c_void_p, C.POINTER(DenseIO), C.c_void_p]
    L.odeu_ekf_dense_run.restype = C.c_int
    L.odeu_ekf_grad_run.argtypes = [C.c_void_p, C.POINTER(EkfIO), C.POINTER(GradIO), C.c_void_p]
    L.odeu_ekf_grad_run.restype = C.c_int
    L.odeu_param_sensitivity.argtypes = [C.c_void_p, C.POINTER(SensIO), C.c_void_p]
    L.odeu_param_sensitivity.restype = C.c_int
    L.odeu_ode_rhs.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p]
    L.odeu_ode_rhs.restype = C.c_int
    L.odeu_launch_count.restype = C.c_int64
    L.odeu_last_error.argtypes = [C.c_char_p, C.c_size_t]
    L.odeu_last_error.restype = C.c_size_t
    _lib = L
    return L


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().odeu_last_error(buf, 1024)
    return buf.value.decode()


def check(rc: int, what: str) -> None:
    """Map the C status convention to the reference's Python exceptions."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (status {rc})"
    if rc < 0:
        raise ValueError(msg)
    raise RuntimeError(msg)
