"""Forward-mode NLL gradients of the kernel source (host-compiled) against
  (a) the reference's own reverse-mode gradient: jax.value_and_grad over the reference's nll()
      (scripts/run_parameter_estimation.py:685-796) executed over the jax look-alike
      (tests/golden/ref_*.npz: grad_norm, sorted-key order, w.r.t. normalised parameters), and
  (b) central finite differences of the kernel's own NLL.
Tolerance: 1e-6 relative on gradients."""
import os

import numpy as np
import pytest

import cases
import util as U
from ode_uncertainty_b200 import ode as O
from ode_uncertainty_b200 import runners

REF_GRAD = {"lv_rkf45_temper_q_only": O.LotkaVolterra, "lv_rkf45_temper_eps_plus_q": O.LotkaVolterra,
            "hh_r4_rkf45_temper": lambda: O.HodgkinHuxley(model="reduced-4"),
            "hh_r1_rkf45_temper": lambda: O.HodgkinHuxley(model="reduced-1"),
            "c3_mhh_r1_rkf45_temper": lambda: O.MultiCompartmentHodgkinHuxley(model="reduced-1", num_compartments=2)}


def _run(backend, name, theta_sorted=None, batch=1):
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    ob = REF_GRAD[name]()
    keys_s, sizes, perm = runners.param_layout(ob)
    inv = np.argsort(perm)                       # builder index of each sorted-order entry
    default_sorted = np.concatenate([ob.params[k].reshape(-1) for k in keys_s])
    ths = default_sorted if theta_sorted is None else theta_sorted
    # differentiate every parameter, requested in sorted-key order
    idx_builder = np.array([int(np.nonzero(perm == j)[0][0]) for j in range(perm.size)])
    nll, g = U.run_grad(backend, plan, np.repeat(m["x0"].reshape(1, -1).numpy(), batch, 0), m["T"], idx_builder, t0=m["t0"],
                        P0_sqrt=m["P0s"].numpy(), theta_shared=ths[perm], Q_sqrt=m["Q"].numpy(),
                        gamma_sqrt=m["gamma"] ** 0.5, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(),
                        ys=m["ys"].numpy(), correct_flags=m["flags"], xy_index_map=m["ymap"])
    return nll[batch - 1], g[batch - 1], default_sorted


def _check(backend, name, batch=1):
    ref = dict(np.load(os.path.join(cases.GOLDEN, f"ref_{name}.npz")))
    nll, g, _ = _run(backend, name, batch=batch)
    assert abs(nll - float(ref["nll_fn"])) <= 1e-9 * abs(float(ref["nll_fn"]))
    g_norm = g * (ref["hi"] - ref["lo"])          # d/d theta_norm = d/d theta * (max - min)
    scale = np.max(np.abs(ref["grad_norm"]))
    np.testing.assert_allclose(g_norm, ref["grad_norm"], rtol=1e-6, atol=1e-6 * scale)


@pytest.mark.parametrize("name", list(REF_GRAD))
def test_forward_mode_gradient_matches_reference_reverse_mode(name):
    _check("hostemu", name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REF_GRAD))
def test_cuda_gradient_matches_reference_reverse_mode(name):
    _check("gpu", name, batch=37)


def test_gradient_matches_finite_differences_lorenz_eps_branch():
    """Lorenz, eps enters P (NOISE_COVFN): d eps/d theta via sign(x0 - x1)."""
    spec = cases.CASES["lorenz_rkf45_obs_full"]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    kw = dict(t0=m["t0"], P0_sqrt=np.eye(3) * 1e-2, H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(),
              ys=m["ys"].numpy()[:40], correct_flags=m["flags"][:40], xy_index_map=m["ymap"][:40])
    x0 = m["x0"].reshape(1, -1).numpy()
    th = plan.default_params.copy()
    nll, g = U.run_grad("hostemu", plan, x0, 40, [0, 1, 2], theta_shared=th, **kw)
    for j in range(3):
        d = 1e-6 * abs(th[j])
        tp, tm = th.copy(), th.copy()
        tp[j] += d
        tm[j] -= d
        fp = U.run_ekf("hostemu", plan, x0, 40, theta_shared=tp, **kw)["nll"][0]
        fm = U.run_ekf("hostemu", plan, x0, 40, theta_shared=tm, **kw)["nll"][0]
        fd = (fp - fm) / (2 * d)
        assert abs(g[0, j] - fd) <= 2e-5 * max(1.0, abs(fd)), (j, g[0, j], fd)
    base = U.run_ekf("hostemu", plan, x0, 40, theta_shared=th, **kw)["nll"][0]
    assert abs(nll[0] - base) <= 1e-10 * abs(base)


# ---- initial_state_parametrized (scripts/run_parameter_estimation.py:744-748): x0 depends on theta
ISP = ["hh_r4_rkf45_temper", "hh_r1_rkf45_temper", "c3_mhh_r1_rkf45_temper"]


def _check_isp(backend, name, batch=1):
    """Against the reference's own nll() with initial_state_parametrized=True and its reverse-mode
    gradient (fixture keys *_isp), at a point off the defaults so that x0(theta) matters."""
    ref = dict(np.load(os.path.join(cases.GOLDEN, f"ref_{name}.npz")))
    if "grad_norm_isp" not in ref:
        pytest.skip("fixture without the initial_state_parametrized gradient")
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    ob = REF_GRAD[name]()
    keys_s, sizes, perm = runners.param_layout(ob)
    theta_sorted = ref["pn_isp"] * (ref["hi"] - ref["lo"]) + ref["lo"]
    x0_raw = np.full((1, 2 if name.startswith("c3") else 1), -70.0)
    idx_sorted = np.arange(theta_sorted.size)
    xb, tan = runners.initial_value_and_tangent(ob, x0_raw, np.repeat(theta_sorted[None], batch, 0), idx_sorted)
    idx_builder = np.array([int(np.nonzero(perm == j)[0][0]) for j in range(perm.size)])
    nll, g = U.run_grad(backend, plan, xb, m["T"], idx_builder, t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                        theta_shared=theta_sorted[perm], Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
                        H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                        xy_index_map=m["ymap"], x0_tangent=tan)
    assert abs(nll[batch - 1] - float(ref["nll_fn_isp"])) <= 1e-9 * abs(float(ref["nll_fn_isp"]))
    g_norm = g[batch - 1] * (ref["hi"] - ref["lo"])
    scale = np.max(np.abs(ref["grad_norm_isp"]))
    np.testing.assert_allclose(g_norm, ref["grad_norm_isp"], rtol=1e-6, atol=1e-6 * scale)
    # x0 really depends on the parameters here (steady-state gates move with V_T); its influence on
    # the gradient is small because the filter forgets its initial condition within a few observations
    assert np.abs(tan).max() > 1e-5 and np.abs(xb[0] - m["x0"].reshape(-1).numpy()).max() > 1e-6


@pytest.mark.parametrize("name", ISP)
def test_initial_state_parametrized_gradient_matches_reference(name):
    _check_isp("hostemu", name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ISP)
def test_cuda_initial_state_parametrized_gradient(name):
    _check_isp("gpu", name, batch=5)


# ---- parameter_sensitivity (scripts/run_parameter_estimation.py:750-769): Q_sqrt = diag(w(theta))
def _check_ps(backend, name, batch=1, isp=False):
    """Against the reference's own nll(..., parameter_sensitivity=True) and its reverse-mode gradient
    (fixture keys *_ps / *_isp_ps), at a point off the defaults."""
    ref = dict(np.load(os.path.join(cases.GOLDEN, f"ref_{name}.npz")))
    sfx = "_isp_ps" if isp else "_ps"
    if "grad_norm" + sfx not in ref:
        pytest.skip("fixture without the parameter_sensitivity gradient")
    spec = cases.CASES[name]
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    ob = REF_GRAD[name]()
    keys_s, sizes, perm = runners.param_layout(ob)
    pn = ref["pn_isp"] if isp else ref["pn_ps"]
    theta_sorted = pn * (ref["hi"] - ref["lo"]) + ref["lo"]
    idx_builder = np.array([int(np.nonzero(perm == j)[0][0]) for j in range(perm.size)])
    tan = None
    if isp:
        x0_raw = np.full((1, 2 if name.startswith("c3") else 1), -70.0)
        xb, tan = runners.initial_value_and_tangent(ob, x0_raw, np.repeat(theta_sorted[None], batch, 0),
                                                    np.arange(theta_sorted.size))
    else:
        xb = np.repeat(m["x0"].reshape(1, -1).numpy(), batch, 0)
    w, wt = U.run_sens(backend, plan, xb, idx_builder, t0=m["t0"], theta_shared=theta_sorted[perm], x0_tangent=tan)
    n = xb.shape[1]
    np.testing.assert_allclose(np.linalg.norm(w, axis=1), np.sqrt(n), rtol=1e-12)
    assert (w >= 0).all()
    nll, g = U.run_grad(backend, plan, xb, m["T"], idx_builder, t0=m["t0"], P0_sqrt=m["P0s"].numpy(),
                        theta_shared=theta_sorted[perm], Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5,
                        H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].numpy(), correct_flags=m["flags"],
                        xy_index_map=m["ymap"], x0_tangent=tan, Q_sqrt_diag=w, Q_sqrt_diag_tangent=wt)
    want = float(ref["nll_fn" + sfx])
    assert abs(nll[batch - 1] - want) <= 1e-9 * abs(want)
    g_norm = g[batch - 1] * (ref["hi"] - ref["lo"])
    scale = np.max(np.abs(ref["grad_norm" + sfx]))
    np.testing.assert_allclose(g_norm, ref["grad_norm" + sfx], rtol=1e-6, atol=1e-6 * scale)
    # the weights change the loss (they replace the configured Q_sqrt) and their tangent matters
    assert np.abs(wt).max() > 0


@pytest.mark.parametrize("name", list(REF_GRAD))
def test_parameter_sensitivity_gradient_matches_reference(name):
    _check_ps("hostemu", name)


@pytest.mark.parametrize("name", ISP)
def test_parameter_sensitivity_with_initial_state_parametrized(name):
    _check_ps("hostemu", name, isp=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(REF_GRAD))
def test_cuda_parameter_sensitivity_gradient(name):
    _check_ps("gpu", name, batch=5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ISP)
def test_cuda_parameter_sensitivity_with_initial_state_parametrized(name):
    _check_ps("gpu", name, batch=3, isp=True)
