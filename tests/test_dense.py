"""Large-state path (BASELINE config 5, odeu_ekf_dense_run: oscillator chain with n = 2 D states,
dense J P J^T on DMMA) against
  * fixtures produced by the REFERENCE'S OWN CODE (scripts/run_filter.py::unroll over
    src/filters/sqrt_ekf.py, src/solvers/rksolver.py, src/ode/lcao.py with a [2, D] state) run
    through the jax look-alike: tests/golden/ref_<case>.npz (oracle/make_golden_ref.py dense), and
  * a live Oracle-A run (guard="intended") for the case on which the reference's sign-sensitive
    zero-gain guard fires.
Tolerances (float64), relative to the largest entry: x 1e-10, nll 1e-9, P 1e-9 when eps does not
enter P, else 5e-6 + the eps-cancellation floor (tests/cases.py::compare)."""
import os

import numpy as np
import pytest
import torch

import cases


def _ref(name):
    p = os.path.join(cases.GOLDEN, f"ref_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} missing (python oracle/make_golden_ref.py dense)")
    return dict(np.load(p))


def _gold(name):
    """Per-step x, eps, diag(P), final P and nll the CUDA path is compared with."""
    spec = cases.DENSE_CASES[name]
    if spec.get("guard") == "intended":
        g = cases.run_oracle(spec)
        return dict(x=g["x"], eps=g["eps"], P_last=g["P"][-1], P_diag=np.stack([np.diag(p) for p in g["P"]]),
                    nll=g["nll"])
    return _ref(name)


@pytest.mark.parametrize("name", list(cases.DENSE_CASES))
def test_dense_oracle_a_equals_reference_code(name):
    """Pins Oracle-A at n = 128 / 256 on the reference code's own output (verbatim guard)."""
    spec = dict(cases.DENSE_CASES[name])
    spec.pop("guard", None)
    ref = _ref(name)
    g = cases.run_oracle(spec)
    np.testing.assert_array_equal(g["t"], ref["t"])
    np.testing.assert_allclose(g["x"], ref["x"], rtol=1e-13, atol=1e-300)
    scale = np.abs(ref["P_last"]).max()
    assert np.abs(g["P"][-1] - ref["P_last"]).max() <= 1e-9 * scale
    assert abs(float(g["nll"]) - float(ref["nll"])) <= 1e-11 * max(1.0, abs(float(ref["nll"])))


def _tols(spec, gold):
    eps_in_P = not spec.get("disable", False)
    xmax = np.abs(gold["x"]).max(axis=1)
    ulp = np.spacing(xmax)
    emax = np.abs(gold["eps"]).max(axis=1)
    s = spec.get("scale", 1.0) if spec.get("cov", "diagonal") != "static_diagonal" else 0.0
    floor = np.cumsum(s * s * (2 * emax * 16 * ulp + (16 * ulp) ** 2)) if eps_in_P else np.zeros_like(emax)
    return (5e-6 if eps_in_P else 1e-9), floor


def _run_dense(spec, T, *, x, P=None, step0=0):
    from ode_uncertainty_b200 import ekf_dense_run
    m = cases.materialize(spec)
    plan = cases.make_plan_for(spec)
    dev = torch.device("cuda:0")
    kw = dict(t0=m["t0"] + step0 * m["h"], Q_sqrt=m["Q"].numpy(), gamma_sqrt=m["gamma"] ** 0.5)
    if P is None:
        kw["P0_sqrt"] = m["P0s"].numpy()
    else:
        kw.update(P=P, inplace=True)
    if m["L"] > 0:
        kw.update(H=m["H"].numpy(), R_sqrt=m["Rs"].numpy(), ys=m["ys"].to(dev),
                  correct_flags=torch.as_tensor(m["flags"][step0:]).to(dev),
                  xy_index_map=torch.as_tensor(m["ymap"][step0:]).to(dev))
    return ekf_dense_run(plan, x, T, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases.DENSE_CASES))
def test_dense_cuda_whole_run(name):
    spec = cases.DENSE_CASES[name]
    gold = _gold(name)
    m = cases.materialize(spec)
    B = 3
    x0 = torch.as_tensor(np.repeat(m["x0"].reshape(1, -1).numpy(), B, 0)).cuda()
    out = _run_dense(spec, m["T"], x=x0)
    rtolP, floor = _tols(spec, gold)
    for b in (0, B - 1):
        x = out.xT[b].cpu().numpy()
        assert np.abs(x - gold["x"][-1]).max() <= 1e-10 * np.abs(gold["x"][-1]).max()
        P = out.PT[b].cpu().numpy()
        assert np.abs(P - gold["P_last"]).max() <= rtolP * np.abs(gold["P_last"]).max() + floor[-1]
        assert np.abs(P - P.T).max() <= 1e-12 * np.abs(P).max()          # stays symmetric to rounding
        nll = float(out.nll[b])
        assert abs(nll - float(gold["nll"])) <= 1e-9 * max(1.0, abs(float(gold["nll"])))
        e = out.epsT[b].cpu().numpy()
        assert np.abs(e - gold["eps"][-1]).max() <= 16 * np.spacing(np.abs(gold["x"][-1]).max()) + 1e-9 * gold["eps"][-1].max()
    assert abs(out.tT - (m["t0"] + m["T"] * m["h"])) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c5_lcao64_rkf45_obs", "lcao64_dopri65_temper"])
def test_dense_cuda_stepwise_resume(name):
    """One step per call, state carried in place: every step's mean and diag(P) against the
    fixture, and the summed NLL."""
    spec = cases.DENSE_CASES[name]
    gold = _gold(name)
    m = cases.materialize(spec)
    B = 2
    x = torch.as_tensor(np.repeat(m["x0"].reshape(1, -1).numpy(), B, 0)).cuda()
    rtolP, floor = _tols(spec, gold)
    P, nll = None, 0.0
    for k in range(m["T"]):
        out = _run_dense(spec, 1, x=x, P=P, step0=k)
        x, P = out.xT, out.PT
        nll += float(out.nll[1])
        xs = x[1].cpu().numpy()
        assert np.abs(xs - gold["x"][k + 1]).max() <= 1e-10 * np.abs(gold["x"][k + 1]).max(), k
        d = torch.diagonal(P[1]).cpu().numpy()
        assert np.abs(d - gold["P_diag"][k + 1]).max() <= rtolP * np.abs(gold["P_diag"][k + 1]).max() + floor[k + 1], k
    assert abs(nll - float(gold["nll"])) <= 1e-9 * max(1.0, abs(float(gold["nll"])))


@pytest.mark.gpu
def test_dense_rejects_unsupported_inputs():
    from ode_uncertainty_b200 import ekf_dense_run
    spec = cases.DENSE_CASES["c5_lcao64_rkf45_obs"]
    plan = cases.make_plan_for(spec)
    x = torch.zeros(2, 128, dtype=torch.float64, device="cuda")
    Hbad = np.zeros((1, 128)); Hbad[0, :2] = 0.5
    with pytest.raises(ValueError):
        ekf_dense_run(plan, x, 1, H=Hbad, R_sqrt=np.eye(1), ys=torch.zeros(1, 1, device="cuda", dtype=torch.float64),
                      correct_flags=torch.ones(1), xy_index_map=torch.zeros(1, dtype=torch.int64))
    with pytest.raises(ValueError):
        ekf_dense_run(plan, x, 1, Q_sqrt=np.ones((128, 128)))
    with pytest.raises(RuntimeError):
        ekf_dense_run(plan, x.cpu(), 1)
