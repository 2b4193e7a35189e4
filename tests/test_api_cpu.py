"""Host-side mirror of the reference's plugin API (no GPU): names, constructor arguments,
parameter dictionaries, tableaux, initial values, observation scheduling, error behaviour."""
import numpy as np
import pytest
import torch

import cases
from oracle import ref_torch as R
from ode_uncertainty_b200 import ode as O
from ode_uncertainty_b200 import runners, solvers as S
from ode_uncertainty_b200.covariance_update_functions import (DiagonalCovarianceUpdate, OuterCovarianceUpdate,
                                                              StaticDiagonalCovarianceUpdate)
from ode_uncertainty_b200.filters import SQRT_EKF, ParticleFilter, solver_handle


def test_tableaux_match_oracle_registry():
    for cls in (S.RKF45, S.Dopri65, S.BS32, S.HeunEuler):
        A, b, c = R.TABLEAUX[cls.tableau]
        sb = cls(step_size=0.01)
        np.testing.assert_array_equal(sb.A, A.numpy())
        np.testing.assert_array_equal(sb.b, b.numpy())
        np.testing.assert_array_equal(sb.c, c.numpy())
        assert sb.s == A.shape[0] and sb.h == 0.01


def test_ode_builders_params_order_and_defaults():
    from ode_uncertainty_b200 import Plan
    for name, (ode_id, variant, nc) in cases.ODE_IDS.items():
        if name.startswith("MultiHH/"):
            b = O.MultiCompartmentHodgkinHuxley(model=name.split("/")[1], num_compartments=2)
        elif name.startswith("HodgkinHuxley/"):
            b = O.HodgkinHuxley(model=name.split("/")[1])
        else:
            b = getattr(O, name)()
        plan = Plan(b.ode_id, 0, 0.01, ode_variant=b.ode_variant, num_compartments=b.num_compartments_abi)
        np.testing.assert_array_equal(b.flat_params(b.params), plan.default_params)
        assert b.state_dim == plan.n
        _, ref_params, shape = cases.ode_and_params(name)
        assert list(b.params) == list(ref_params)
        assert b.shape == tuple(shape)


def test_hh_initial_values_match_oracle():
    for model in ("full", "reduced-1", "reduced-4"):
        b = O.HodgkinHuxley(model=model)
        got = b.build_initial_value(np.array([[-70.0]]), b.params)
        ref = R.hh_initial_value(model, -70.0, R.ODES[f"HodgkinHuxley/{model}"][1])
        np.testing.assert_allclose(got, ref.numpy(), rtol=1e-14)
    mb = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)
    got = mb.build_initial_value(np.array([[-70.0, -65.0]]), mb.params)
    assert got.shape == (1, 14)
    np.testing.assert_allclose(got[0, :7], O.HodgkinHuxley("reduced-1").build_initial_value(
        np.array([[-70.0]]), {k: v for k, v in O.HodgkinHuxley("reduced-1").params.items()})[0], rtol=1e-14)
    with pytest.raises(ValueError, match="Unknown model"):
        O.HodgkinHuxley(model="reduced-2")


def test_error_behaviour_matches_reference():
    sb = S.RKF45(step_size=0.01)
    with pytest.raises(AttributeError, match="Setup solver before usage!"):   # rksolver.py:99-100
        sb.build()
    with pytest.raises(AttributeError, match="Setup solver before usage!"):
        solver_handle(sb)
    with pytest.raises(NotImplementedError):                                  # filter.py:122-133
        ParticleFilter().build_correct()
    with pytest.raises(NotImplementedError):
        DiagonalCovarianceUpdate().build()(np.eye(2), np.ones(2))
    with pytest.raises(ValueError, match="Unsupported filter builder"):       # run_filter.py:146
        runners.run_filter(filter_builder=object(), solver_builder=S.RKF45(0.1), ode_builder=O.Lorenz(),
                           x0="[[1.0, 1.0, 1.0]]", tN=0.2)


def test_cov_builders_carry_plugin_identity():
    assert DiagonalCovarianceUpdate(2.0).build_sqrt().scale == 2.0
    assert OuterCovarianceUpdate().build().cov_fn_id == 1
    assert StaticDiagonalCovarianceUpdate(0.5).build_sqrt().static
    f = SQRT_EKF(cov_update_fn_builder=OuterCovarianceUpdate(3.0), disable_cov_update=True)
    assert f.disable_cov_update and f.build_cov_update_fn().scale == 3.0


def test_init_state_shapes_like_reference():
    sb = S.RKF45(0.01)
    st = SQRT_EKF().init_state(sb.init_state(0.0, np.ones((1, 3))), np.eye(3) * 1e-12, np.zeros((3, 3)),
                               0.0, np.eye(2) * 0.1)
    assert st["t"].shape == (1,) and st["x"].shape == (1, 1, 3) and st["eps"].shape == (1, 1, 3)
    assert st["P_sqrt"].shape == (1, 3, 3) and st["y"].shape == (2,) and st["y_hat"].shape == (1, 2)
    assert st["S_sqrt"].shape == (1, 2, 2) and st["Q_sqrt"].shape == (3, 3)
    pst = ParticleFilter(num_particles=5).init_state(sb.init_state(0.0, np.ones((1, 3))), 7)
    assert pst["t"].shape == (5,) and pst["x"].shape == (5, 1, 3)


def test_sync_times_and_schedule():
    # observations every 3rd filter step, float arange like the reference
    t0, tN, h = 0.0, 1.0, 0.01
    ts_y = np.arange(0.0, 1.0 + 1e-9, 0.03)
    T, flags, ymap = runners.observation_schedule(t0, tN, h, ts_y)
    assert T == 100 and flags.shape == (100,)
    assert flags.sum() == 33 and flags[2] and not flags[0]          # t = 0.03 is step index 2
    np.testing.assert_array_equal(ymap[flags], np.arange(1, 34))    # the observation at t0 is never used
    xi, yi = runners.sync_times(np.array([0.1, 0.2, 0.3, 0.4]), np.array([0.2 + 5e-9, 0.3 - 5e-9, 0.5]))
    assert xi.tolist() == [1, 2] and yi.tolist() == [0, 1]
    # faithful port incl. the reference's edge case: with a single matching time the clamped index
    # makes isin_tolerance accept every later observation, and sync_times' own assert fires
    # (src/utils.py:181-215) - reproduced, not "fixed"
    with pytest.raises(AssertionError):
        runners.sync_times(np.array([0.1, 0.2, 0.3]), np.array([0.2, 0.5]))


def test_param_layout_sorted_vs_builder_order():
    b = O.HodgkinHuxley("reduced-1")
    keys_s, sizes, perm = runners.param_layout(b)
    assert keys_s[:6] == ["A", "C", "E_Ca", "E_K", "E_Na", "E_leak"]     # ASCII order, SURVEY 7.3-7
    flat_sorted = np.concatenate([np.asarray(b.params[k]).reshape(-1) for k in keys_s])
    np.testing.assert_array_equal(flat_sorted[perm], b.flat_params(b.params))
    mb = O.MultiCompartmentHodgkinHuxley()
    keys_s, sizes, perm = runners.param_layout(mb)
    flat_sorted = np.concatenate([np.asarray(mb.params[k]).reshape(-1) for k in keys_s])
    np.testing.assert_array_equal(flat_sorted[perm], mb.flat_params(mb.params))


def test_observation_schedule_equals_reference_sync_times():
    """Port of sync_times / the float-arange schedule against the reference's own function
    (fixture generated by running src/utils.py:sync_times, oracle/make_golden_ref.py)."""
    import os
    f = dict(np.load(os.path.join(cases.GOLDEN, "ref_sync_times.npz")))
    for tag in ("every1", "every3", "hh", "offset"):
        t0, tN, h, dt_y = f[f"{tag}_args"]
        ts_y = np.arange(t0, tN + 1e-9, dt_y)
        T, flags, ymap = runners.observation_schedule(t0, tN, h, ts_y)
        assert flags.shape[0] == int(f[f"{tag}_nx"])
        np.testing.assert_array_equal(np.nonzero(flags)[0], f[f"{tag}_x"])
        np.testing.assert_array_equal(ymap[flags], f[f"{tag}_y"])


def test_single_compartment_multi_hh_plan_is_the_single_compartment_plan():
    """odeu_plan_create maps MultiCompartmentHodgkinHuxley(num_compartments = 1) onto the single-compartment model
    (same equations and flat parameter order, hodgkin_huxley.py:284-439): same state dimension, parameter count and -
    through the host emulation of the kernel source - the same filter run bit for bit; other compartment counts
    than 1 and 2 are refused with an error, not served by something else."""
    import util as U
    from ode_uncertainty_b200 import Plan, _native as N
    from ode_uncertainty_b200 import ode as O
    single = O.HodgkinHuxley(model="reduced-1")
    multi = O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=1, coupling_coeffs="[]", C=1.0, A="[8.3e-5]",
                                            g_Na="[25.0]", E_Na="[53.0]", g_K="[7.0]", E_K="[-107.0]", g_leak="[0.1]",
                                            E_leak="[-70.0]", V_T="[-60.0]", g_M="[0.01]", tau_max="[4e3]", g_L="[0.01]",
                                            E_Ca="[120.0]", g_T="[0.01]", V_x="[2.0]")
    np.testing.assert_array_equal(multi.flat_params(multi.params), single.flat_params(single.params))
    p1 = Plan(N.ODE_HODGKIN_HUXLEY, N.SOLVER_RKF45, 0.01, ode_variant=1, disable_cov_update=True)
    pm = Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=1, disable_cov_update=True)
    assert (pm.n, pm.p) == (p1.n, p1.p) == (7, 15)
    x0 = single.build_initial_value(np.array([[-70.0]]), single.params).reshape(1, -1)
    np.testing.assert_array_equal(multi.build_initial_value(np.array([[-70.0]]), multi.params).reshape(1, -1), x0)
    T = 12
    H = np.zeros((1, 7)); H[0, 0] = 1.0
    kw = dict(t0=9.9, P0_sqrt=np.eye(7) * 1e-6, Q_sqrt=np.eye(7), gamma_sqrt=1e-2, H=H, R_sqrt=np.eye(1) * 0.3,
              ys=np.full((T, 1), -65.0), correct_flags=np.ones(T, np.uint8), xy_index_map=np.arange(T, dtype=np.int64),
              theta_shared=single.flat_params(single.params), minimal=True)
    a = U.run_ekf("hostemu", p1, x0, T, **kw)
    b = U.run_ekf("hostemu", pm, x0, T, **kw)
    for k in ("xT", "PT", "nll"):
        np.testing.assert_array_equal(a[k], b[k])
    with pytest.raises(ValueError, match="compartments=3"):
        Plan(N.ODE_MULTI_HH, N.SOLVER_RKF45, 0.01, ode_variant=1, num_compartments=3)
