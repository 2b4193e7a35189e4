"""Solver plugins: mirror of `src/solvers/{solver,rksolver,rkf45,dopri65,bs32,heun_euler}.py`.

`build()` returns the single-step transition `state -> next_state` (API compatibility, tests);
it runs the plain RK kernel (`odeu_pf_run` with `noise_free`) for one step.  The production
path never steps through Python: filters hand the tableau id to the fused whole-trajectory
kernel."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from . import _native as N
from .ode import ODE


class SolverBuilder:
    """src/solvers/solver.py:15-75."""
    solver_id = -1

    def __init__(self, step_size: float = 0.1) -> None:
        self.h = step_size

    def setup(self, *args, **kwargs) -> None:
        pass

    def init_state(self, t0, x0) -> Dict[str, torch.Tensor]:
        x0 = torch.as_tensor(np.asarray(x0.cpu() if isinstance(x0, torch.Tensor) else x0), dtype=torch.float64)
        return {"t": torch.as_tensor(float(t0), dtype=torch.float64), "x": x0,
                "eps": torch.zeros_like(x0), "diffrax_state": torch.zeros(())}

    def build(self):
        raise NotImplementedError

    def build_parametrized(self):
        raise NotImplementedError


class RKSolverBuilder(SolverBuilder):
    """src/solvers/rksolver.py:11-157."""
    tableau = ""

    def __init__(self, step_size: float = 0.1) -> None:
        super().__init__(step_size=step_size)
        self.A = self.build_A()
        self.b = self.build_b()
        self.c = self.build_c()
        self.s = self.A.shape[0]

    def setup(self, ode: ODE, params: Dict[str, np.ndarray], *args, **kwargs) -> None:
        self.ode = ode
        self.params = params

    @classmethod
    def build_A(cls) -> np.ndarray:
        raise NotImplementedError

    @classmethod
    def build_b(cls) -> np.ndarray:
        raise NotImplementedError

    @classmethod
    def build_c(cls) -> np.ndarray:
        raise NotImplementedError

    def build(self):
        if not hasattr(self, "ode") or not hasattr(self, "params"):
            raise AttributeError("Setup solver before usage!")   # rksolver.py:99-100
        ps = self.build_parametrized()
        return lambda state: ps(self.ode, self.params, state)

    def build_parametrized(self):
        from .engine import Plan, pf_run

        def parametrized_solve(ode: ODE, params, state):
            """One step for a batch: state["t"] [] or [M], state["x"] [M, N, D] or [N, D] (CUDA)."""
            b = ode.builder
            x = state["x"]
            if not x.is_cuda:
                raise RuntimeError("solver steps run on the GPU only (no CPU fallback)")
            single = x.ndim == 2
            xb = x.reshape(1 if single else x.shape[0], -1)
            plan = Plan(b.ode_id, self.solver_id, self.h, ode_variant=b.ode_variant,
                        num_compartments=b.num_compartments_abi)
            t = state["t"]
            t0 = float(t.reshape(-1)[0]) if isinstance(t, torch.Tensor) else float(t)
            r = pf_run(plan, xb.shape[0], 1, x0=xb, t0=t0, theta_shared=b.flat_params(params),
                       noise_free=True)
            t_next = (t + self.h) if isinstance(t, torch.Tensor) else torch.as_tensor(t0 + self.h)
            return {"t": t_next, "x": r.xT.reshape(x.shape), "eps": r.epsT.reshape(x.shape),
                    "diffrax_state": torch.zeros(())}

        return parametrized_solve


def _arr(x):
    return np.array(x, dtype=np.float64)


class RKF45(RKSolverBuilder):
    """src/solvers/rkf45.py:7-34 (S=6)."""
    solver_id, tableau = N.SOLVER_RKF45, "RKF45"

    @classmethod
    def build_A(cls):
        return _arr([[0, 0, 0, 0, 0, 0], [1 / 4, 0, 0, 0, 0, 0], [3 / 32, 9 / 32, 0, 0, 0, 0],
                     [1932 / 2197, -7200 / 2197, 7296 / 2197, 0, 0, 0],
                     [439 / 216, -8.0, 3680 / 513, -845 / 4104, 0, 0],
                     [-8 / 27, 2.0, -3544 / 2565, 1859 / 4104, -11 / 40, 0]])

    @classmethod
    def build_b(cls):
        return _arr([[16 / 135, 0, 6656 / 12825, 28561 / 56430, -9 / 50, 2 / 55],
                     [25 / 216, 0, 1408 / 2565, 2197 / 4104, -1 / 5, 0]])

    @classmethod
    def build_c(cls):
        return _arr([0, 1 / 4, 3 / 8, 12 / 13, 1, 1 / 2])


class Dopri65(RKSolverBuilder):
    """src/solvers/dopri65.py:7-72 (S=8)."""
    solver_id, tableau = N.SOLVER_DOPRI65, "Dopri65"

    @classmethod
    def build_A(cls):
        return _arr([[0] * 8, [1 / 10] + [0] * 7, [-2 / 81, 20 / 81] + [0] * 6,
                     [615 / 1372, -270 / 343, 1053 / 1372] + [0] * 5,
                     [3243 / 5500, -54 / 55, 50949 / 71500, 4998 / 17875] + [0] * 4,
                     [-26492 / 37125, 72 / 55, 2808 / 23375, -24206 / 37125, 338 / 459] + [0] * 3,
                     [5561 / 2376, -35 / 11, -24117 / 31603, 899983 / 200772, -5225 / 1836, 3925 / 4056, 0, 0],
                     [465467 / 266112, -2945 / 1232, -5610201 / 14158144, 10513573 / 3212352,
                      -424325 / 205632, 376225 / 454272, 0, 0]])

    @classmethod
    def build_b(cls):
        return _arr([[821 / 10800, 0, 19683 / 71825, 175273 / 912600, 395 / 3672, 785 / 2704, 3 / 50, 0],
                     [61 / 864, 0, 98415 / 321776, 16807 / 146016, 1375 / 7344, 1375 / 5408, -37 / 1120, 1 / 10]])

    @classmethod
    def build_c(cls):
        return _arr([0, 1 / 10, 2 / 9, 3 / 7, 3 / 5, 4 / 5, 1.0, 1.0])


class BS32(RKSolverBuilder):
    """src/solvers/bs32.py:7-32 (S=4)."""
    solver_id, tableau = N.SOLVER_BS32, "BS32"

    @classmethod
    def build_A(cls):
        return _arr([[0, 0, 0, 0], [1 / 2, 0, 0, 0], [0, 3 / 4, 0, 0], [2 / 9, 1 / 3, 4 / 9, 0]])

    @classmethod
    def build_b(cls):
        return _arr([[7 / 24, 1 / 4, 1 / 3, 1 / 8], [2 / 9, 1 / 3, 4 / 9, 0]])

    @classmethod
    def build_c(cls):
        return _arr([0, 1 / 2, 3 / 4, 1.0])


class HeunEuler(RKSolverBuilder):
    """src/solvers/heun_euler.py:7-30 (S=2; b[1] = [0.5, 0] verbatim)."""
    solver_id, tableau = N.SOLVER_HEUN_EULER, "HeunEuler"

    @classmethod
    def build_A(cls):
        return _arr([[0, 0], [1.0, 0]])

    @classmethod
    def build_b(cls):
        return _arr([[0.5, 0.5], [0.5, 0.0]])

    @classmethod
    def build_c(cls):
        return _arr([0, 1.0])


class DiffraxSolverBuilder(RKSolverBuilder):
    """src/solvers/diffrax_solver.py:16-140: the implicit (stiff) solver plugin every shipped
    Hodgkin-Huxley configuration selects (`name: Kvaerno3`, configs/params/hodgkinhuxley*.yaml:10-14).

    The reference delegates to diffrax (`Kvaerno3` / `ImplicitEuler` with `Newton(rtol=atol=1e-8)`,
    one fixed step per call, `eps = 0`); diffrax is third-party and not part of this image, so the
    published methods are restated in CUDA (csrc/dirk.cuh: diagonally implicit RK, full Newton with a
    dense LU per stage, step Jacobian and parameter tangents by the implicit function theorem).
    Parity with diffrax is unpinned; the oracle is oracle/ref_torch.py::dirk_step."""
    _IDS = {"Kvaerno3": N.SOLVER_KVAERNO3, "ImplicitEuler": N.SOLVER_IMPLICIT_EULER}

    def __init__(self, name: str = "ImplicitEuler", step_size: float = 0.1) -> None:
        SolverBuilder.__init__(self, step_size=step_size)
        if name not in self._IDS:
            raise ValueError(f"DiffraxSolverBuilder(name={name!r}): the B200 path serves {sorted(self._IDS)} "
                             "(explicit methods: use the RKF45 / Dopri65 / BS32 / HeunEuler plugins)")
        self.name = self.tableau = name
        self.solver_id = self._IDS[name]
        if name == "Kvaerno3":
            g = 0.43586652150845899941601945
            self.A = _arr([[0, 0, 0, 0], [g, g, 0, 0],
                           [(-4 * g * g + 6 * g - 1) / (4 * g), (-2 * g + 1) / (4 * g), g, 0],
                           [(6 * g - 1) / (12 * g), -1 / ((24 * g - 12) * g), (-6 * g * g + 6 * g - 1) / (6 * g - 3), g]])
            self.c = _arr([0, 2 * g, 1, 1])
        else:
            self.A, self.c = _arr([[1.0]]), _arr([1.0])
        self.b = np.stack([self.A[-1], self.A[-1]])         # stiffly accurate; no embedded error is used (eps = 0)
        self.s = self.A.shape[0]
