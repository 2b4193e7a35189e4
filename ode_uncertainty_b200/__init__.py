"""B200-native batched EKF-over-embedded-Runge-Kutta path of f-lair/ode-uncertainty.

Hand-written sm_100a CUDA behind a C ABI (include/odeu.h, libodeu.so); Python only marshals
buffers.  No CPU fallback: the CUDA library must be built (`make`) and inputs must be CUDA
tensors.
"""
from . import _native
from .engine import DenseResult, EkfResult, PfResult, Plan, ekf_dense_run, ekf_grad_run, ekf_run, launch_count, pf_run

__all__ = ["Plan", "ekf_run", "ekf_grad_run", "pf_run", "ekf_dense_run", "DenseResult", "EkfResult", "PfResult", "launch_count", "_native"]
