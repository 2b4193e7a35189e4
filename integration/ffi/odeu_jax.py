"""JAX side of the gated XLA FFI shim (integration/ffi/odeu_ffi.cc): registers the handler and wraps
`jax.ffi.ffi_call`.  Needs jax + jaxlib with CUDA and the built libodeu_ffi.so; not importable in the
image this repository is developed in (no jax), hence not exercised by tests/ - the tested binding
is the ctypes one (integration/b200.py -> ode_uncertainty_b200/reference_binding.py)."""
import ctypes
import os

import jax
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_HERE, "libodeu_ffi.so"))
jax.ffi.register_ffi_target("odeu_ekf_run", jax.ffi.pycapsule(_lib.OdeuEkfRun), platform="CUDA")


def ekf_run(plan_handle: int, x0, T: int, *, t0=0.0, P0_sqrt, theta_shared, Q_sqrt=None, gamma_sqrt=0.0, H=None,
            R_sqrt=None, ys=None, correct_flags=None, xy_index_map=None, save_interval=0, guard_mode=1):
    """x0 [B, n] jax.Array on the GPU -> (xT [B, n], PT [B, n, n], nll [B], traj dict).  Jittable."""
    B, n = x0.shape
    L = 0 if H is None else int(np.asarray(H).shape[0])
    Ts = T // save_interval + 1 if save_interval > 0 else 0
    f64 = jnp.float64
    shp = lambda *s: jax.ShapeDtypeStruct(s, f64)
    empty = shp(0)
    outs = (shp(n, B), shp(n * n, B), shp(B),
            shp(Ts, n, B) if Ts else empty, shp(Ts, n, B) if Ts else empty, shp(Ts, n * n, B) if Ts else empty,
            shp(Ts, n * n, B) if Ts and guard_mode else empty, shp(Ts, L, B) if Ts and L else empty,
            shp(Ts, L * L, B) if Ts and L else empty, shp(Ts) if Ts else empty)
    z = jnp.zeros((0,), f64)
    call = jax.ffi.ffi_call("odeu_ekf_run", outs)
    flat = lambda a: np.zeros(0) if a is None else np.asarray(a, dtype=np.float64).reshape(-1)
    xT, PT, nll, ox, oe, oP, oPs, oy, oS, ot = call(
        x0.T, z, z if ys is None else jnp.asarray(ys, f64).reshape(-1, L),
        jnp.zeros((0,), jnp.uint8) if correct_flags is None else jnp.asarray(correct_flags, jnp.uint8),
        jnp.zeros((0,), jnp.int64) if xy_index_map is None else jnp.asarray(xy_index_map, jnp.int64),
        plan=np.int64(plan_handle), T=np.int64(T), L=np.int64(L), save_interval=np.int64(save_interval),
        guard_mode=np.int64(guard_mode), ys_per_trajectory=np.int64(0), t0=np.float64(t0),
        gamma_sqrt=np.float64(gamma_sqrt), P0_sqrt=flat(P0_sqrt), Q_sqrt=flat(Q_sqrt), H=flat(H), R_sqrt=flat(R_sqrt),
        theta_shared=flat(theta_shared))
    traj = dict(t=ot, x=jnp.transpose(ox, (0, 2, 1)), eps=jnp.transpose(oe, (0, 2, 1)),
                P=jnp.transpose(oP, (0, 2, 1)).reshape(Ts, B, n, n)) if Ts else None
    return xT.T, PT.T.reshape(B, n, n), nll, traj
