"""ORACLE (test infrastructure): ctypes front-end of Oracle-B (oracle/ekf_ref.cpp), the
C++/std::thread restatement of the reference's square-root EKF.  Used by tests as a fast second
oracle (long horizons, full batch sizes) and by bench.py as the timed CPU baseline / reference
arm.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle_b.so")
_lib = None

ODE = {"Lorenz": (0, 0, 0, 3, 3), "VanDerPol": (1, 0, 0, 2, 1), "LotkaVolterra": (2, 0, 0, 2, 4),
       "Pendulum": (3, 0, 0, 2, 1), "LCAO": (4, 2, 0, 4, 3),
       "HodgkinHuxley/full": (5, 0, 0, 8, 15), "HodgkinHuxley/reduced-1": (5, 1, 0, 7, 15),
       "HodgkinHuxley/reduced-4": (5, 4, 0, 4, 15),
       "MultiHH/reduced-1/2": (6, 1, 2, 14, 30), "MultiHH/reduced-4/2": (6, 4, 2, 8, 30)}
SOLVER = {"RKF45": 0, "Dopri65": 1, "BS32": 2, "HeunEuler": 3}
COV = {"diagonal": 0, "outer": 1, "static_diagonal": 2}


def ode_key(name: str):
    """Registry lookup; "LCAO/<D>" = the oscillator chain with D oscillators (n = 2 D)."""
    if name.startswith("LCAO/"):
        D = int(name.split("/")[1])
        return (4, D, 0, 2 * D, 3)
    return ODE[name]


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "ekf_ref.cpp")
        if not os.path.exists(_SO) or (os.path.exists(src) and os.path.getmtime(_SO) < os.path.getmtime(src)):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = C.CDLL(_SO)
        _lib.oracle_ekf_run.restype = C.c_longlong
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _a(x, dt=np.float64):
    return None if x is None else np.ascontiguousarray(np.asarray(x, dtype=dt))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def ekf_run(ode: str, solver: str, h: float, x0, T: int, *, t0=0.0, P0_sqrt=None, theta=None,
            Q_sqrt=None, gamma_sqrt=0.0, H=None, R_sqrt=None, ys=None, ys_per_trajectory=False,
            correct_flags=None, xy_index_map=None, cov="diagonal", scale=1.0, disable=False,
            save_interval=0, guard="reference", nthreads=0, theta_default=None, count_fragile=True):
    """x0 [B, n]; theta None (-> theta_default), [p] shared or [B, p].  Returns dict like
    tests/util.run_ekf (xT, PT, nll, traj, guard_mismatch_steps, guard_fired_steps)."""
    oid, variant, nc, n, p = ode_key(ode)
    x0 = _a(x0).reshape(-1, n)
    B = x0.shape[0]
    P0s = _a(np.eye(n) * 1e-12 if P0_sqrt is None else P0_sqrt).reshape(n, n)
    th = _a(theta_default if theta is None else theta)
    per = int(th.ndim == 2)
    assert th.shape[-1] == p
    Q = _a(Q_sqrt)
    L = 0
    Hn = Rn = ysn = fl = mp = None
    if H is not None and ys is not None:
        Hn = _a(H)
        L = Hn.shape[0]
        Rn = _a(R_sqrt).reshape(L, L)
        ysn = _a(ys)
        fl = _a(correct_flags, np.uint8)
        mp = _a(xy_index_map, np.int64)
    xT, PT, nll = np.zeros((B, n)), np.zeros((B, n, n)), np.zeros(B)
    tr = {}
    if save_interval > 0:
        Ts = T // save_interval + 1
        tr = dict(t=np.zeros(Ts), x=np.zeros((Ts, B, n)), eps=np.zeros((Ts, B, n)),
                  P=np.zeros((Ts, B, n, n)), y_hat=np.zeros((Ts, B, L)), S=np.zeros((Ts, B, L, L)))
    fired, fragile = C.c_longlong(0), C.c_longlong(0)
    mism = lib().oracle_ekf_run(
        C.c_int(oid), C.c_int(variant), C.c_int(nc), C.c_int(n), C.c_int(p), C.c_int(SOLVER[solver]),
        C.c_double(h), C.c_int(COV[cov]), C.c_double(scale), C.c_int(int(disable)),
        C.c_longlong(B), C.c_longlong(T), C.c_double(t0), _p(x0), _p(P0s), _p(th), C.c_int(per),
        _p(Q), C.c_double(gamma_sqrt), C.c_int(L), _p(Hn), _p(Rn), _p(ysn),
        C.c_int(int(ys_per_trajectory)), _p(fl), _p(mp), C.c_longlong(save_interval),
        C.c_int(int(guard == "intended")), C.c_int(nthreads), _p(xT), _p(PT), _p(nll),
        _p(tr.get("t")), _p(tr.get("x")), _p(tr.get("eps")), _p(tr.get("P")),
        _p(tr.get("y_hat")) if L else None, _p(tr.get("S")) if L else None, C.byref(fired),
        C.byref(fragile) if count_fragile else None)
    # fragile_qr_columns: Householder columns whose sign the reference's own LAPACK decides by rounding
    # (sub-column at cancellation-noise level, see oracle/ekf_ref.cpp householder_R)
    return dict(xT=xT, PT=PT, nll=nll, traj=tr or None, guard_mismatch_steps=int(mism),
                guard_fired_steps=int(fired.value), fragile_qr_columns=int(fragile.value))


def rk_run(ode: str, solver: str, h: float, x0, T: int, *, t0=0.0, theta=None):
    oid, variant, nc, n, p = ode_key(ode)
    x0 = _a(x0).reshape(n)
    th = _a(theta)
    xs, es = np.zeros((T + 1, n)), np.zeros((T + 1, n))
    lib().oracle_rk_run(C.c_int(oid), C.c_int(variant), C.c_int(nc), C.c_int(n), C.c_int(p),
                        C.c_int(SOLVER[solver]), C.c_double(h), C.c_longlong(T), C.c_double(t0),
                        _p(x0), _p(th), _p(xs), _p(es))
    return xs, es
