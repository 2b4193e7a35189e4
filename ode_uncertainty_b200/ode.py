"""ODE plugins: host-side mirror of the reference's `src/ode/` builders.

Same class names, constructor keywords, `params` dictionaries (insertion order = the C ABI's flat
parameter layout) and `build_initial_value` semantics as the reference; the right-hand side
itself is device code (csrc/odes.cuh), so `build()` returns a handle that evaluates
`f(t, x, params)` for a batch on the GPU through `odeu_ode_rhs` (there is no CPU evaluation).
"""
from __future__ import annotations

import ctypes as C
from ast import literal_eval
from typing import Dict, Optional

import numpy as np
import torch

from . import _native as N


class ODE:
    """Callable `(t, x[..., N, D], params) -> dx/dt` on CUDA tensors (src/ode/ode.py:6-7)."""

    def __init__(self, builder: "ODEBuilder") -> None:
        self.builder = builder

    def __call__(self, t, x: torch.Tensor, params: Optional[Dict[str, np.ndarray]] = None) -> torch.Tensor:
        from .engine import Plan
        if not x.is_cuda:
            raise RuntimeError("ODE right-hand sides run on the GPU only (no CPU fallback)")
        b = self.builder
        plan = Plan(b.ode_id, N.SOLVER_RKF45, 1.0, ode_variant=b.ode_variant,
                    num_compartments=b.num_compartments_abi)
        n = plan.n
        shape = x.shape
        xb = x.reshape(-1, n).to(torch.float64).t().contiguous()
        dx = torch.empty_like(xb)
        th = np.ascontiguousarray(b.flat_params(params if params is not None else b.params))
        st = torch.cuda.current_stream(x.device)
        with torch.cuda.device(x.device):
            N.check(N.lib().odeu_ode_rhs(plan.handle, xb.shape[1], float(t), C.c_void_p(xb.data_ptr()),
                                         None, th.ctypes.data_as(C.c_void_p), C.c_void_p(dx.data_ptr()),
                                         C.c_void_p(st.cuda_stream)), "odeu_ode_rhs")
        return dx.t().reshape(shape)


class ODEBuilder:
    """src/ode/ode.py:10-53."""
    ode_id = -1
    ode_variant = 0
    num_compartments_abi = 0
    shape = (1, 1)          # [N, D] of the state

    def __init__(self, **kwargs) -> None:
        self.params: Dict[str, np.ndarray] = {}
        for k, v in kwargs.items():
            if isinstance(v, (np.ndarray, torch.Tensor)):
                self.params[k] = np.asarray(v, dtype=np.float64)
            elif isinstance(v, float):
                self.params[k] = np.array(v, dtype=np.float64)

    def build(self) -> ODE:
        if self.ode_id < 0:
            raise NotImplementedError
        return ODE(self)

    def build_initial_value(self, initial_value, params) -> np.ndarray:
        return np.asarray(initial_value, dtype=np.float64)

    def flat_params(self, params: Dict[str, np.ndarray]) -> np.ndarray:
        """Flat vector in builder (insertion) order: the `theta` layout of include/odeu.h."""
        return np.concatenate([np.asarray(params[k], dtype=np.float64).reshape(-1) for k in self.params])

    @property
    def state_dim(self) -> int:
        return self.shape[0] * self.shape[1]


class Lorenz(ODEBuilder):
    """src/ode/lorenz.py:9-54."""
    ode_id, shape = N.ODE_LORENZ, (1, 3)

    def __init__(self, sigma: float = 10.0, beta: float = 8.0 / 3, rho: float = 28.0) -> None:
        super().__init__(sigma=sigma, beta=beta, rho=rho)


class VanDerPol(ODEBuilder):
    """src/ode/van_der_pol.py:9-46."""
    ode_id, shape = N.ODE_VAN_DER_POL, (2, 1)

    def __init__(self, damping: float = 5.0) -> None:
        super().__init__(damping=damping)


class LotkaVolterra(ODEBuilder):
    """src/ode/lotka_volterra.py:9-54."""
    ode_id, shape = N.ODE_LOTKA_VOLTERRA, (1, 2)

    def __init__(self, alpha: float = 1.5, beta: float = 1.0, gamma: float = 3.0, delta: float = 1.0) -> None:
        super().__init__(alpha=alpha, beta=beta, gamma=gamma, delta=delta)


class Pendulum(ODEBuilder):
    """src/ode/pendulum.py:9-46."""
    ode_id, shape = N.ODE_PENDULUM, (2, 1)

    def __init__(self, length: float = 3.0) -> None:
        super().__init__(length=length)


class LCAO(ODEBuilder):
    """src/ode/lcao.py:9-63 (D = 2 positions + 2 velocities).  The reference's right-hand side
    works for any number D of oscillators (the coupling is `flip(x)`, :57-59) and takes D from the
    shape of the state; here `num_oscillators` selects it (128 = BASELINE config 5, served by
    `ekf_dense_run`)."""
    ode_id, ode_variant, shape = N.ODE_LCAO, 2, (2, 2)

    def __init__(self, lin_coeff: float = 1.0, cubic_coeff: float = 2.0, coupling_coeff: float = 0.5,
                 num_oscillators: int = 2) -> None:
        super().__init__(lin_coeff=lin_coeff, cubic_coeff=cubic_coeff, coupling_coeff=coupling_coeff)
        self.ode_variant = int(num_oscillators)
        self.shape = (2, int(num_oscillators))


# ---- Hodgkin-Huxley steady states for build_initial_value (src/ode/hodgkin_huxley.py:12-36) ----
_e = np.exp
_a_m = lambda V, V_T: -0.32 * (V - V_T - 13.0) / (_e(-(V - V_T - 13.0) / 4.0) - 1.0)
_b_m = lambda V, V_T: 0.28 * (V - V_T - 40.0) / (_e((V - V_T - 40.0) / 5.0) - 1.0)
_a_n = lambda V, V_T: -0.032 * (V - V_T - 15.0) / (_e(-(V - V_T - 15.0) / 5.0) - 1.0)
_b_n = lambda V, V_T: 0.5 * _e(-(V - V_T - 10.0) / 40.0)
_a_h = lambda V, V_T: 0.128 * _e(-(V - V_T - 17.0) / 18.0)
_b_h = lambda V, V_T: 4.0 / (1.0 + _e(-(V - V_T - 40.0) / 5.0))
_a_q = lambda V: 0.055 * (-27.0 - V) / (_e((-27.0 - V) / 3.8) - 1.0)
_b_q = lambda V: 0.94 * _e((-75.0 - V) / 17.0)
_a_r = lambda V: 0.000457 * _e((-13.0 - V) / 50.0)
_b_r = lambda V: 0.0065 / (_e((-15.0 - V) / 28.0) + 1.0)
_HH_DIM = {"full": 8, "reduced-1": 7, "reduced-4": 4}
_HH_VARIANT = {"full": 0, "reduced-1": 1, "reduced-4": 4}


def _hh_steady_state(model: str, V0: float, p: Dict[str, float]) -> np.ndarray:
    V_T, V_x = float(p["V_T"]), float(p["V_x"])
    vals = [V0,
            1.0 / (1.0 + _b_m(V0, V_T) / _a_m(V0, V_T)),
            1.0 / (1.0 + _b_h(V0, V_T) / _a_h(V0, V_T)),
            1.0 / (1.0 + _b_n(V0, V_T) / _a_n(V0, V_T)),
            1.0 / (1.0 + _e(-(V0 + 35.0) / 10.0)),
            1.0 / (1.0 + _b_q(V0) / _a_q(V0)),
            1.0 / (1.0 + _b_r(V0) / _a_r(V0)),
            1.0 / (1.0 + _e((V0 + V_x + 81.0) / 4.0))]
    return np.array(vals[:_HH_DIM[model]], dtype=np.float64)


class HodgkinHuxley(ODEBuilder):
    """src/ode/hodgkin_huxley.py:61-281."""
    ode_id = N.ODE_HODGKIN_HUXLEY

    def __init__(self, model: str = "reduced-1", C: float = 1.0, A: float = 8.3e-5, g_Na: float = 25.0,
                 E_Na: float = 53.0, g_K: float = 7.0, E_K: float = -107.0, g_leak: float = 0.1,
                 E_leak: float = -70.0, V_T: float = -60.0, g_M: float = 0.01, tau_max: float = 4e3,
                 g_L: float = 0.01, E_Ca: float = 120.0, g_T: float = 0.01, V_x: float = 2.0) -> None:
        if model not in _HH_DIM:
            raise ValueError(f"Unknown model: {model}")   # hodgkin_huxley.py:249
        super().__init__(C=C, A=A, g_Na=g_Na, E_Na=E_Na, g_K=g_K, E_K=E_K, g_leak=g_leak, E_leak=E_leak,
                         V_T=V_T, g_M=g_M, tau_max=tau_max, g_L=g_L, E_Ca=E_Ca, g_T=g_T, V_x=V_x)
        self.model = model
        self.ode_variant = _HH_VARIANT[model]
        self.shape = (1, _HH_DIM[model])

    def build_initial_value(self, initial_value, params) -> np.ndarray:
        V0 = float(np.asarray(initial_value, dtype=np.float64)[0, 0])
        return _hh_steady_state(self.model, V0, params)[None, :]


class MultiCompartmentHodgkinHuxley(ODEBuilder):
    """src/ode/hodgkin_huxley.py:284-439 (parameters given as list strings, like the YAMLs)."""
    ode_id = N.ODE_MULTI_HH

    def __init__(self, model: str = "reduced-1", num_compartments: int = 2, coupling_coeffs: str = "[1.0]",
                 C: float = 1.0, A: str = "[4.15e-5, 4.15e-5]", g_Na: str = "[25.0, 20.0]",
                 E_Na: str = "[53.0, 53.0]", g_K: str = "[7.0, 10.0]", E_K: str = "[-107.0, -107.0]",
                 g_leak: str = "[0.09, 0.11]", E_leak: str = "[-70.0, -70.0]", V_T: str = "[-60.0, -60.0]",
                 g_M: str = "[0.01, 0.01]", tau_max: str = "[4e3, 4e3]", g_L: str = "[0.01, 0.01]",
                 E_Ca: str = "[120.0, 120.0]", g_T: str = "[0.01, 0.01]", V_x: str = "[2.0, 2.0]") -> None:
        if model not in _HH_DIM:
            raise ValueError(f"Unknown model: {model}")
        arr = lambda s: np.array(literal_eval(s), dtype=np.float64)
        super().__init__(coupling_coeffs=arr(coupling_coeffs)[None, :], C=np.array([C], dtype=np.float64),
                         A=arr(A), g_Na=arr(g_Na), E_Na=arr(E_Na), g_K=arr(g_K), E_K=arr(E_K),
                         g_leak=arr(g_leak), E_leak=arr(E_leak), V_T=arr(V_T), g_M=arr(g_M),
                         tau_max=arr(tau_max), g_L=arr(g_L), E_Ca=arr(E_Ca), g_T=arr(g_T), V_x=arr(V_x))
        self.model = model
        self.num_compartments = num_compartments
        self.num_compartments_abi = num_compartments
        self.ode_variant = _HH_VARIANT[model]
        self.D_dim = _HH_DIM[model]
        self.shape = (1, num_compartments * self.D_dim)

    def build_initial_value(self, initial_value, params) -> np.ndarray:
        iv = np.asarray(initial_value, dtype=np.float64)
        out = []
        for c in range(self.num_compartments):
            pc = {k: np.broadcast_to(v, (self.num_compartments,) + v.shape[1:])[c] for k, v in params.items()}
            out.append(_hh_steady_state(self.model, float(iv[0, c]), pc))
        return np.concatenate(out)[None, :]
