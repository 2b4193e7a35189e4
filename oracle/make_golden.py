"""ORACLE tooling (test infrastructure): regenerate tests/golden/oracleA_*.npz from Oracle-A.

    python oracle/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

if __name__ == "__main__":
    names = sys.argv[1:] or list(cases.CASES)
    os.makedirs(cases.GOLDEN, exist_ok=True)
    for name in names:
        out = cases.run_oracle(cases.CASES[name])
        np.savez_compressed(cases.golden_path(name), **out)
        print(f"{name}: T={out['t'].shape[0] - 1} nll={float(out['nll']):.12g} "
              f"guard_mismatch={int(out['guard_mismatch_steps'])} guard_fired={int(out['guard_fired_steps'])}")
