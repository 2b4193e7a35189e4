"""The C-ABI shared library loads and exports every symbol include/odeu.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from ode_uncertainty_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "odeu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(odeu_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(N.LIB_PATH)
    declared = _declared_symbols()
    assert set(declared) == set(N.SYMBOLS), (declared, N.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), f"libodeu.so does not export {s}"
    assert N.lib().odeu_version() == 1


def test_struct_layouts_match_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "odeu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)

    def fields(struct_name):
        body = re.search(r"typedef struct \{([^}]*)\} " + struct_name + ";", hdr).group(1)
        return [re.findall(r"(\w+)\s*;", line)[0] for line in body.split("\n") if ";" in line]

    assert fields("odeu_plan_desc") == [f[0] for f in N.PlanDesc._fields_]
    assert fields("odeu_ekf_io") == [f[0] for f in N.EkfIO._fields_]
    assert fields("odeu_pf_io") == [f[0] for f in N.PfIO._fields_]


def test_plan_metadata_for_every_plugin():
    from ode_uncertainty_b200 import Plan
    import cases
    expect = {"Lorenz": (3, 3), "VanDerPol": (2, 1), "LotkaVolterra": (2, 4), "Pendulum": (2, 1),
              "LCAO": (4, 3), "HodgkinHuxley/full": (8, 15), "HodgkinHuxley/reduced-1": (7, 15),
              "HodgkinHuxley/reduced-4": (4, 15), "MultiHH/reduced-1/2": (14, 30),
              "MultiHH/reduced-4/2": (8, 30)}
    for name, (ode_id, variant, nc) in cases.ODE_IDS.items():
        for solver in cases.SOLVERS.values():
            p = Plan(ode_id=ode_id, solver_id=solver, step_size=0.01, ode_variant=variant,
                     num_compartments=nc)
            assert (p.n, p.p) == expect[name]


def test_product_path_refuses_cpu_tensors():
    import pytest
    import torch
    from ode_uncertainty_b200 import Plan, ekf_run
    p = Plan(ode_id=N.ODE_LORENZ, solver_id=N.SOLVER_RKF45, step_size=0.01)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ekf_run(p, torch.ones(2, 3, dtype=torch.float64), 3)
