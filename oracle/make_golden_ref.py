"""ORACLE tooling (test infrastructure): golden vectors produced by the REFERENCE'S OWN CODE.

Executes, unmodified, the reference sources under /root/reference
    src/ode/*.py, src/solvers/{rksolver,rkf45,dopri65,bs32,heun_euler}.py, src/utils.py,
    src/covariance_update_functions/*.py, src/filters/{filter,sqrt_ekf}.py,
    scripts/run_filter.py::unroll, scripts/run_parameter_estimation.py::nll
over the torch-backed `jax` look-alike in oracle/jax_shim (JAX itself is not installable here),
on the parity cases of tests/cases.py, and stores the results as tests/golden/ref_<case>.npz.
These fixtures pin Oracle-A/B (tests/test_oracle_ref.py) and are compared directly with the
CUDA path.  Needs /root/reference; the committed fixtures travel instead of it.

    python oracle/make_golden_ref.py [case ...]
"""
import copy
import importlib.util
import os
import sys
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ODEU_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "jax_shim"), os.path.join(ROOT, "oracle", "jax_shim", "stubs"), REF]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import jax  # noqa: E402  (the shim)
import torch  # noqa: E402
from jax import numpy as jnp  # noqa: E402
from jax.flatten_util import ravel_pytree  # noqa: E402

import src.covariance_update_functions as ref_cov  # noqa: E402
import src.filters as ref_filters  # noqa: E402
import src.ode as ref_ode  # noqa: E402
import src.solvers as ref_solvers  # noqa: E402
from src.utils import negative_log_gaussian_sqrt  # noqa: E402


def _load_script(name):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, "scripts", f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


run_filter = _load_script("run_filter")
run_pe = _load_script("run_parameter_estimation")

import cases  # noqa: E402


def builders(spec):
    name = spec["ode"]
    if name.startswith("MultiHH/"):
        _, model, nc = name.split("/")
        ob = ref_ode.MultiCompartmentHodgkinHuxley(model=model, num_compartments=int(nc))
    elif name.startswith("HodgkinHuxley/"):
        ob = ref_ode.HodgkinHuxley(model=name.split("/")[1])
    elif name.startswith("LCAO/"):
        ob = ref_ode.LCAO()          # the state shape [2, D] carries D (src/ode/lcao.py:51-61)
    else:
        ob = getattr(ref_ode, name)()
    sb = getattr(ref_solvers, spec["solver"])(step_size=spec.get("h", 0.01))
    cov = spec.get("cov", "diagonal")
    scale = spec.get("scale", 1.0)
    kw = {}
    if cov == "diagonal":
        kw["cov_update_fn_builder"] = ref_cov.DiagonalCovarianceUpdate(scale=scale)
    elif cov == "outer":
        kw["cov_update_fn_builder"] = ref_cov.OuterCovarianceUpdate(scale=scale)
    else:
        kw["static_cov_update_fn_builder"] = ref_cov.StaticDiagonalCovarianceUpdate(scale=scale)
    fb = ref_filters.SQRT_EKF(disable_cov_update=spec.get("disable", False), **kw)
    return ob, sb, fb, cov == "static_diagonal"


def run_case(spec):
    """Mirrors scripts/run_filter.py:main (lines 71-161) around the reference's own unroll()."""
    m = cases.materialize(spec)
    ob, sb, fb, use_static = builders(spec)
    ode = ob.build()
    sb.setup(ode, ob.params)
    solver = jax.jit(jax.vmap(sb.build()))
    filter_predict = jax.jit(fb.build_predict(), static_argnums=(0, 1))
    if use_static:
        cov_update_fn = partial(fb.build_static_cov_update_fn(), fb.static_cov_update_fn_builder.scale)
    else:
        cov_update_fn = fb.build_cov_update_fn()
    L = m["L"]
    if L > 0:
        filter_correct = fb.build_correct()
        H, ys = m["H"], m["ys"]
        flags = torch.as_tensor(m["flags"].astype(bool))
        ymap = torch.as_tensor(m["ymap"])
    else:
        filter_correct = lambda _, x: x
        H, ys = jnp.eye(m["n"]), jnp.zeros((1, 0))
        flags = jnp.zeros(m["T"], dtype=bool)
        ymap = jnp.zeros(m["T"], dtype=int)
    solver_state = sb.init_state(jnp.array(m["t0"]), m["x0"])
    st = fb.init_state(solver_state, m["P0s"], m["Q"], jnp.array(m["gamma"] ** 0.5), m["Rs"])
    traj = run_filter.unroll(filter_predict, filter_correct, solver, cov_update_fn, H, st, ys, flags,
                             ymap, m["T"], 1, True)
    Ts = m["T"] + 1
    out = dict(t=traj["t"].reshape(Ts).numpy(), x=traj["x"].reshape(Ts, -1).numpy(),
               eps=traj["eps"].reshape(Ts, -1).numpy())
    Ps = traj["P_sqrt"].reshape(Ts, m["n"], m["n"])
    out["P"] = (Ps @ Ps.transpose(1, 2)).numpy()
    out["y_hat"] = traj["y_hat"].reshape(Ts, L).numpy()
    Ss = traj["S_sqrt"].reshape(Ts, L, L)
    out["S"] = (Ss @ Ss.transpose(1, 2)).numpy()
    # NLL exactly as scripts/run_parameter_estimation.py:771-794 adds it: nlg on observation steps
    nll = 0.0
    if L > 0:
        for idx in range(m["T"]):
            if bool(flags[idx]):
                y = ys[int(ymap[idx])]
                nll = nll + float(negative_log_gaussian_sqrt(y, traj["y_hat"][idx + 1][0], traj["S_sqrt"][idx + 1][0]))
    out["nll"] = np.array(nll)
    return out


def run_nll_and_grad(spec):
    """The reference's own nll() (scripts/run_parameter_estimation.py:685-796) and its gradient
    w.r.t. the normalised parameters (the quantity jaxopt.ScipyBoundedMinimize differentiates,
    :599), for the tempering cases.  All parameters optimised, range = default * [0.5, 2]."""
    m = cases.materialize(spec)
    ob, sb, fb, _ = builders(spec)
    ode = ob.build()
    sb.setup(ode, ob.params)
    solver = jax.vmap(sb.build_parametrized(), (None, None, 0))
    filter_predict = fb.build_parametrized_predict()
    filter_correct = fb.build_correct()
    cov_update_fn = fb.build_cov_update_fn()
    solver_state = sb.init_state(jnp.array(m["t0"]), m["x0"])
    st = fb.init_state(solver_state, m["P0s"], m["Q"], jnp.array(m["gamma"] ** 0.5), m["Rs"])
    keys = list(ob.params)
    shp = {k: ob.params[k].shape[-1:] for k in keys}
    lo = {k: torch.minimum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(shp[k]) for k in keys}
    hi = {k: torch.maximum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(shp[k]) for k in keys}
    opt = {k: jnp.full(shp[k], True) for k in keys}
    pn = {k: (ob.params[k].reshape(shp[k]) - lo[k]) / (hi[k] - lo[k]) for k in keys}
    idx = jnp.flatnonzero(ravel_pytree(opt)[0])
    flags = torch.as_tensor(m["flags"].astype(bool))
    ymap = torch.as_tensor(m["ymap"])

    def f(params_norm):
        return run_pe.nll(m["T"], False, False, filter_predict, filter_correct, solver, ode,
                          ob.build_initial_value, cov_update_fn, params_norm, copy.copy(st), m["x0"],
                          m["H"], m["ys"], flags, ymap, lo, hi, opt, idx, ob.params)

    val, g = jax.value_and_grad(f)(pn)
    gflat, _ = ravel_pytree(g)          # sorted-key order, like the optimiser sees it
    out = dict(nll_fn=np.array(float(val)), grad_norm=gflat.numpy(), grad_names=np.array(sorted(keys)),
               lo=ravel_pytree(lo)[0].numpy(), hi=ravel_pytree(hi)[0].numpy())
    # parameter_sensitivity=True (:750-769): Q_sqrt = diag(w(theta)), w from |d x_1 / d theta| of one
    # solver step, inside the differentiated loss.  Evaluated OFF the defaults.
    pn_ps = {k: pn[k] + 0.05 for k in keys}

    def f_ps(params_norm):
        return run_pe.nll(m["T"], False, True, filter_predict, filter_correct, solver, ode,
                          ob.build_initial_value, cov_update_fn, params_norm, copy.copy(st), m["x0"],
                          m["H"], m["ys"], flags, ymap, lo, hi, opt, idx, ob.params)

    val3, g3 = jax.value_and_grad(f_ps)(pn_ps)
    out.update(nll_fn_ps=np.array(float(val3)), grad_norm_ps=ravel_pytree(g3)[0].numpy(),
               pn_ps=ravel_pytree(pn_ps)[0].numpy())
    # initial_state_parametrized=True (:744-748): x0 = build_initial_value(V0, theta) is a function of
    # the parameters (Hodgkin-Huxley steady-state gates depend on V_T, V_x); evaluated at a point OFF
    # the defaults so that x0(theta) differs from the fixed x0 of the case
    name = spec["ode"]
    if name.startswith("HodgkinHuxley/") or name.startswith("MultiHH/"):
        nc = int(name.split("/")[2]) if name.startswith("MultiHH/") else 1
        x0_raw = jnp.full((1, nc), -70.0)
        pn_isp = {k: pn[k] + (0.07 if k in ("V_T", "V_x") else 0.0) for k in keys}

        def f_isp(params_norm):
            return run_pe.nll(m["T"], True, False, filter_predict, filter_correct, solver, ode,
                              ob.build_initial_value, cov_update_fn, params_norm, copy.copy(st), x0_raw,
                              m["H"], m["ys"], flags, ymap, lo, hi, opt, idx, ob.params)

        val2, g2 = jax.value_and_grad(f_isp)(pn_isp)
        out.update(nll_fn_isp=np.array(float(val2)), grad_norm_isp=ravel_pytree(g2)[0].numpy(),
                   pn_isp=ravel_pytree(pn_isp)[0].numpy())

        def f_isp_ps(params_norm):      # both options together
            return run_pe.nll(m["T"], True, True, filter_predict, filter_correct, solver, ode,
                              ob.build_initial_value, cov_update_fn, params_norm, copy.copy(st), x0_raw,
                              m["H"], m["ys"], flags, ymap, lo, hi, opt, idx, ob.params)

        val4, g4 = jax.value_and_grad(f_isp_ps)(pn_isp)
        out.update(nll_fn_isp_ps=np.array(float(val4)), grad_norm_isp_ps=ravel_pytree(g4)[0].numpy())
    return out


def run_baseline_nll_and_grad(spec):
    """The reference's plain-RK least-squares baseline loss (scripts/run_parameter_estimation_baseline.py
    ::nll :552-632) and its gradient w.r.t. the normalised parameters, off the defaults."""
    run_base = _load_script("run_parameter_estimation_baseline")
    m = cases.materialize(spec)
    ob, sb, _, _ = builders(spec)
    ode = ob.build()
    sb.setup(ode, ob.params)
    solver = sb.build_parametrized()
    st = sb.init_state(jnp.array(m["t0"]), m["x0"])
    keys = list(ob.params)
    shp = {k: ob.params[k].shape[-1:] for k in keys}
    lo = {k: torch.minimum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(shp[k]) for k in keys}
    hi = {k: torch.maximum(0.5 * ob.params[k], 2.0 * ob.params[k]).reshape(shp[k]) for k in keys}
    opt = {k: jnp.full(shp[k], True) for k in keys}
    pn = {k: (ob.params[k].reshape(shp[k]) - lo[k]) / (hi[k] - lo[k]) + 0.05 for k in keys}
    idx = jnp.flatnonzero(ravel_pytree(opt)[0])
    flags = torch.as_tensor(m["flags"].astype(bool))
    ymap = torch.as_tensor(m["ymap"])

    def f(params_norm):
        return run_base.nll(m["T"], False, solver, ode, ob.build_initial_value, params_norm, copy.copy(st),
                            m["x0"], m["H"], m["ys"], m["Rs"], flags, ymap, lo, hi, idx, ob.params)

    val, g = jax.value_and_grad(f)(pn)
    return dict(nll_base=np.array(float(val)), grad_norm_base=ravel_pytree(g)[0].numpy(),
                pn_base=ravel_pytree(pn)[0].numpy(), lo=ravel_pytree(lo)[0].numpy(), hi=ravel_pytree(hi)[0].numpy())


def sync_times_fixture():
    """The reference's own sync_times (src/utils.py:181-215) on the float aranges its scripts
    build (scripts/run_filter.py:99-105) for a few observation cadences."""
    from src.utils import sync_times
    out = {}
    for tag, (t0, tN, h, dt_y) in {"every1": (0.0, 2.0, 0.01, 0.01), "every3": (0.0, 2.0, 0.01, 0.03),
                                   "hh": (0.0, 100.0, 0.01, 0.01), "offset": (10.0, 12.0, 0.01, 0.05)}.items():
        ts_y = jnp.arange(t0, tN + 1e-9, dt_y)
        ts_x = jnp.arange(t0 + h, tN + h, h)
        xi, yi = sync_times(ts_x, ts_y)
        out[f"{tag}_args"] = np.array([t0, tN, h, dt_y])
        out[f"{tag}_x"] = xi.numpy()
        out[f"{tag}_y"] = yi.numpy()
        out[f"{tag}_nx"] = np.array(ts_x.shape[0])
    np.savez_compressed(os.path.join(cases.GOLDEN, "ref_sync_times.npz"), **out)
    print("sync_times fixture:", {k: v.shape for k, v in out.items() if k.endswith("_x")})


GRAD_CASES = ["lv_rkf45_temper_q_only", "lv_rkf45_temper_eps_plus_q", "hh_r4_rkf45_temper",
              "hh_r1_rkf45_temper", "c3_mhh_r1_rkf45_temper"]

def dense_fixture(name):
    """Large-state cases: the reference code's trajectory, storing the covariance of the last
    step only (a [T+1, n, n] stack would be megabytes)."""
    out = run_case(cases.DENSE_CASES[name])
    keep = dict(t=out["t"], x=out["x"], eps=out["eps"], y_hat=out["y_hat"], nll=out["nll"],
                P_last=out["P"][-1], P_diag=np.stack([np.diag(p) for p in out["P"]]))
    np.savez_compressed(os.path.join(cases.GOLDEN, f"ref_{name}.npz"), **keep)
    print(f"{name}: nll={float(out['nll']):.12g}")


if __name__ == "__main__":
    if sys.argv[1:] and sys.argv[1] == "dense":
        for name in (sys.argv[2:] or list(cases.DENSE_CASES)):
            dense_fixture(name)
        sys.exit(0)
    if sys.argv[1:] and sys.argv[1] == "baseline":
        for name in (sys.argv[2:] or GRAD_CASES):
            out = run_baseline_nll_and_grad(cases.CASES[name])
            np.savez_compressed(os.path.join(cases.GOLDEN, f"ref_baseline_{name}.npz"), **out)
            print(f"baseline {name}: nll={float(out['nll_base']):.12g}")
        sys.exit(0)
    names = sys.argv[1:] or list(cases.CASES)
    if not sys.argv[1:] or "sync_times" in names:
        sync_times_fixture()
        names = [n for n in names if n != "sync_times"]
    for name in names:
        spec = cases.CASES[name]
        out = run_case(spec)
        if name in GRAD_CASES:
            try:
                out.update(run_nll_and_grad(spec))
            except Exception as e:  # keep the trajectory fixture even if the gradient path fails
                print(f"  [grad skipped for {name}: {type(e).__name__}: {e}]")
        np.savez_compressed(os.path.join(cases.GOLDEN, f"ref_{name}.npz"), **out)
        extra = f" nll_fn={float(out['nll_fn']):.12g}" if "nll_fn" in out else ""
        print(f"{name}: nll={float(out['nll']):.12g}{extra}")
