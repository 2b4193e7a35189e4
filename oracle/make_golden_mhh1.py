"""ORACLE tooling (test infrastructure): the reference's OWN MultiCompartmentHodgkinHuxley with num_compartments = 1
(src/ode/hodgkin_huxley.py:284-439, empty coupling_coeffs) evaluated over the torch-backed `jax` look-alike:
right-hand side at a few states + its initial value, stored as tests/golden/ref_mhh1_rhs.npz.  Pins the mapping of
single-compartment multi-compartment plans onto the single-compartment kernels (csrc/api.cu, odeu_plan_create).
Needs /root/reference; the committed fixture travels instead of it.

    python oracle/make_golden_mhh1.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ODEU_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(ROOT, "oracle", "jax_shim"), os.path.join(ROOT, "oracle", "jax_shim", "stubs"), REF]

import jax  # noqa: E402,F401  (the shim)
from jax import numpy as jnp  # noqa: E402

import src.ode as ref_ode  # noqa: E402

ARGS = dict(coupling_coeffs="[]", C=1.1, A="[7.9e-5]", g_Na="[24.0]", E_Na="[52.0]", g_K="[7.5]", E_K="[-105.0]", g_leak="[0.095]",
            E_leak="[-69.0]", V_T="[-59.0]", g_M="[0.012]", tau_max="[3.9e3]", g_L="[0.011]", E_Ca="[118.0]", g_T="[0.009]",
            V_x="[2.5]")
out = {"args": np.array(repr(ARGS))}
rng = np.random.default_rng(5)
for model in ("reduced-1", "reduced-4", "full"):
    ob = ref_ode.MultiCompartmentHodgkinHuxley(model=model, num_compartments=1, **ARGS)
    ode = ob.build()
    x0 = ob.build_initial_value(jnp.asarray([[-68.0]]), ob.params)
    n = int(np.asarray(x0).size)
    xs, fs, ts = [], [], []
    for k in range(6):
        x = np.asarray(x0, dtype=np.float64).reshape(1, n).copy()
        x[0, 0] += rng.uniform(-15, 40)
        x[0, 1:] = np.clip(x[0, 1:] + rng.uniform(-0.1, 0.1, n - 1), 0.01, 0.99)
        t = 9.5 + 0.4 * k                                   # straddles the stimulus onset
        f = ode(jnp.asarray(t), jnp.asarray(x), ob.params)
        xs.append(x.reshape(-1)); fs.append(np.asarray(f, dtype=np.float64).reshape(-1)); ts.append(t)
    out[f"x_{model}"], out[f"f_{model}"], out[f"t_{model}"] = np.array(xs), np.array(fs), np.array(ts)
    out[f"x0_{model}"] = np.asarray(x0, dtype=np.float64).reshape(-1)
    print(model, n, np.abs(out[f"f_{model}"]).max())
np.savez(os.path.join(ROOT, "tests", "golden", "ref_mhh1_rhs.npz"), **out)
