"""Particle ensemble (src/filters/particle_filter.py:73-118): the random stream is keyed by global
particle and global step, with the Box-Muller pairs of an ODD number of normals per step shared by
two consecutive steps (pf_thread.cuh) - so results must not depend on where a run is cut (resume
on an even or an odd step) or how the particles are sharded, and every step's noise must be
standard normal and independent of its neighbour's."""
import numpy as np
import pytest

import util as U
from ode_uncertainty_b200 import _native as N


def _plan(ode_id=N.ODE_LORENZ, cov=N.COV_DIAGONAL):
    return U.make_plan(ode_id=ode_id, solver_id=N.SOLVER_RKF45, step_size=0.01, cov_fn_id=cov, cov_scale=1.0)


def _check(backend, M):
    for ode_id, x0, cov in ((N.ODE_LORENZ, [1.0, 1.0, 1.0], N.COV_DIAGONAL),      # 3 normals per step (odd)
                            (N.ODE_LORENZ, [1.0, 1.0, 1.0], N.COV_OUTER),         # 1 normal per step (odd)
                            (N.ODE_VAN_DER_POL, [2.0, 0.5], N.COV_DIAGONAL)):     # 2 normals per step (even)
        plan = _plan(ode_id, cov)
        full = U.run_pf(backend, plan, M, 12, x0_shared=x0, seed=7)
        for cut in (4, 5):                                     # resume on an even / odd global step
            a = U.run_pf(backend, plan, M, cut, x0_shared=x0, seed=7)
            b = U.run_pf(backend, plan, M, 12 - cut, x0_shared=x0, x0=a["xT"], t0=a["tT"], seed=7, step_offset=cut)
            assert np.array_equal(b["xT"], full["xT"]), (ode_id, cov, cut)
        part = U.run_pf(backend, plan, 7, 12, x0_shared=x0, seed=7, particle_offset=5)
        assert np.array_equal(part["xT"], full["xT"][5:12])


def test_resume_and_sharding_invariance_hostemu():
    _check("hostemu", 16)


@pytest.mark.gpu
def test_resume_and_sharding_invariance_gpu():
    _check("gpu", 4096)


def _one_step_z(backend, M, step_offset, seed):
    plan = _plan()
    r = U.run_pf(backend, plan, M, 1, x0_shared=[1.0, 1.0, 1.0], seed=seed, step_offset=step_offset)
    return (r["xT"][1:] - r["xT"][0]) / r["epsT"][1:]          # particle 0 is noise-free


@pytest.mark.parametrize("backend,M", [("hostemu", 20000), pytest.param("gpu", 400000, marks=pytest.mark.gpu)])
def test_every_step_draws_independent_standard_normals(backend, M):
    z0 = _one_step_z(backend, M, 0, 3)       # even step: normals 0..2 of the pair's blocks
    z1 = _one_step_z(backend, M, 1, 3)       # odd step: the kept normal + one more block
    tol = 5 / (M - 1) ** 0.5
    for z in (z0, z1):
        assert np.abs(z.mean(0)).max() < tol
        assert np.abs(z.var(0) - 1.0).max() < 2 * tol * 2 ** 0.5
        c = np.corrcoef(z.T)
        assert np.abs(c - np.eye(3)).max() < tol
        assert abs(np.mean(z ** 4) - 3.0) < 12 * tol             # kurtosis of a Gaussian
    cross = z0.T @ z1 / (M - 1)                                   # the two steps of a pair share blocks
    assert np.abs(cross).max() < tol
