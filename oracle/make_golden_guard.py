"""ORACLE tooling (test infrastructure): LAPACK-backed fixture for the sign-sensitive zero-gain guard
on the HEADLINE workload (BASELINE config 2, Van der Pol half: H = I, observation at every step).

Oracle-A (oracle/ref_torch.py: torch.linalg.qr = LAPACK geqrf, the same sign convention as the
reference's jsp.linalg.qr on CPU) runs trajectory 0 of bench.workload_inputs("VanDerPol") with the
guard applied verbatim (`all(S_sqrt < 1e-16)`, src/filters/sqrt_ekf.py:351).  On this workload the
guard FIRES for a healthy all-negative factor (first at step 1007), i.e. the reference drops that
observation; the fixture records where, so tests can pin the factor-form kernels (guard_mode
reference) and Oracle-B's Householder to LAPACK's actual decisions.

    python oracle/make_golden_guard.py          # ~25 s, writes tests/golden/oracleA_c2_vdp_guardref.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import ref_torch as R  # noqa: E402

T = 1100

if __name__ == "__main__":
    torch.set_default_dtype(torch.float64)
    w = bench.workload_inputs("VanDerPol", 8, T, 0)
    ys = bench.observations("VanDerPol", T, w)
    ode, params, shape = R.ODES["VanDerPol"]
    st = R.init_state(w["t0"], torch.tensor(w["x0"][0]).reshape(shape), torch.tensor(w["P0_sqrt"]),
                      torch.zeros(2, 2), 0.0, torch.tensor(w["R_sqrt"]))
    solver = lambda t, x: R.rk_step(ode, params, "RKF45", 0.01, t, x)
    cov = R.cov_update_sqrt("diagonal", 1.0)
    H = torch.tensor(w["H"])
    xs, Ps, fired, nll = [], [], [], torch.zeros(())
    for k in range(T):
        st["y"] = torch.tensor(ys[k])
        st = R.ekf_predict(solver, cov, False, st)
        st = R.ekf_correct(H, st, "reference")
        nll = nll + R.negative_log_gaussian_sqrt(st["y"], st["y_hat"][0], st["S_sqrt"][0])
        fired.append(bool(st["guard_ref"]))
        xs.append(st["x"][0].flatten().numpy().copy())
        Ps.append(st["P_sqrt"][0].numpy().copy())
    fired = np.array(fired)
    out = dict(T=T, x=np.array(xs), P_sqrt=np.array(Ps), fired=fired, nll=float(nll))
    p = os.path.join(ROOT, "tests", "golden", "oracleA_c2_vdp_guardref.npz")
    np.savez_compressed(p, **out)
    print(f"wrote {p}: guard fired at steps {np.nonzero(fired)[0].tolist()}, nll = {float(nll):.12g}")
