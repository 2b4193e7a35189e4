"""Config-file front end: the reference's YAML experiment definitions select the B200 plugins.

The reference builds its plugin objects with jsonargparse from `class_path` / `init_args` trees
(`CLI(main, as_positional=False)`, scripts/run_filter.py:227-228; e.g.
configs/ekf_trajectory_conrad_baseline/rkf45/lorenz.yaml:2-18).  This module reads the SAME trees
with PyYAML and instantiates the classes of this package that carry the same names and
`init_args` (`src.filters.SQRT_EKF` -> `ode_uncertainty_b200.filters.SQRT_EKF`, ...), then calls
the runner that mirrors the script's `main()`:

    python -m ode_uncertainty_b200.cli run_filter            --config cfg.yaml [--key value ...]
    python -m ode_uncertainty_b200.cli run_parameter_estimation optimize --config cfg.yaml
    python -m ode_uncertainty_b200.cli run_parameter_estimation evaluate --config cfg.yaml
    python -m ode_uncertainty_b200.cli run_parameter_estimation_baseline optimize --config cfg.yaml
    python -m ode_uncertainty_b200.cli run_calibration       --config cfg.yaml

Differences to the reference CLI: results are written as `.npz` with the reference's dataset names
(h5py is not part of this image; an `.h5` output path is rewritten to `.npz`), and observation
files (`y_path`) are `.npz` files with datasets `t`, `x` (what scripts/run_ode_solver.py stores).
Keys the B200 path has no use for (`disable_pbar`, `num_processes`) are ignored.
"""
from __future__ import annotations

import argparse
import importlib
import os
import sys
from typing import Any, Dict

import numpy as np
import yaml

# reference module -> module of this package with the same class names
_MODULES = {
    "src.filters": "ode_uncertainty_b200.filters",
    "src.filters.sqrt_ekf": "ode_uncertainty_b200.filters",
    "src.filters.particle_filter": "ode_uncertainty_b200.filters",
    "src.solvers": "ode_uncertainty_b200.solvers",
    "src.ode": "ode_uncertainty_b200.ode",
    "src.covariance_update_functions": "ode_uncertainty_b200.covariance_update_functions",
    "src.noise_schedules": "ode_uncertainty_b200.noise_schedules",
}
_IGNORED = {"disable_pbar", "num_processes"}


def instantiate(node: Any) -> Any:
    """Recursively build `{class_path, init_args}` trees (jsonargparse's convention)."""
    if isinstance(node, dict) and "class_path" in node:
        mod, _, cls = node["class_path"].rpartition(".")
        if mod.startswith("ode_uncertainty_b200"):
            target = mod
        elif mod in _MODULES:
            target = _MODULES[mod]
        else:
            raise ValueError(f"class_path {node['class_path']!r}: no B200 plugin for module {mod!r}")
        klass = getattr(importlib.import_module(target), cls, None)
        if klass is None:
            raise ValueError(f"class_path {node['class_path']!r}: {target} has no class {cls!r}")
        kwargs = {k: instantiate(v) for k, v in (node.get("init_args") or {}).items()}
        return klass(**kwargs)
    if isinstance(node, dict):
        return {k: instantiate(v) for k, v in node.items()}
    return node


def load_config(path: str, overrides: Dict[str, str] | None = None) -> Dict[str, Any]:
    with open(path) as fh:
        cfg = yaml.safe_load(fh) or {}
    for k, v in (overrides or {}).items():
        cfg[k] = yaml.safe_load(v)
    cfg = {k: instantiate(v) for k, v in cfg.items() if k not in _IGNORED}
    out = cfg.get("output")
    if isinstance(out, str) and out.endswith(".h5"):
        cfg["output"] = out[:-3] + ".npz"
    return cfg


def _load_observations(cfg: Dict[str, Any]) -> None:
    """`y_path` -> (ts_y, ys_x), like scripts/run_parameter_estimation.py:132-147 reads the H5 file."""
    yp = cfg.pop("y_path", None)
    if yp is None:
        raise ValueError("y_path is required")
    if yp.endswith(".h5"):
        yp = yp[:-3] + ".npz"
    dat = np.load(yp)
    cfg["ts_y"], cfg["ys_x"] = dat["t"], dat["x"]


def main(argv=None) -> Dict[str, np.ndarray]:
    ap = argparse.ArgumentParser(prog="ode_uncertainty_b200.cli", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("script", choices=["run_filter", "run_parameter_estimation", "run_parameter_estimation_baseline",
                                       "run_calibration"])
    ap.add_argument("subcommand", nargs="?", default=None, help="optimize (run_parameter_estimation)")
    ap.add_argument("--config", required=True)
    args, rest = ap.parse_known_args(argv)
    if len(rest) % 2 or any(not k.startswith("--") for k in rest[0::2]):
        ap.error("overrides must be given as --key value pairs")
    overrides = {k[2:]: v for k, v in zip(rest[0::2], rest[1::2])}
    cfg = load_config(args.config, overrides)
    if not args.script.startswith("run_parameter_estimation"):
        cfg.pop("initial_state_parametrized", None)
        cfg.pop("parameter_sensitivity", None)
    if isinstance(cfg.get("output"), str):
        os.makedirs(os.path.dirname(os.path.abspath(cfg["output"])), exist_ok=True)
    from . import estimation, runners
    if args.script == "run_filter":
        return runners.run_filter(**cfg)
    if args.script == "run_calibration":
        return runners.calibration(**cfg)
    if args.subcommand not in (None, "optimize", "evaluate"):
        ap.error("run_parameter_estimation serves the subcommands `optimize` and `evaluate`")
    output = cfg.pop("output", None)
    _load_observations(cfg)
    if args.subcommand == "evaluate":                        # scripts/run_parameter_estimation.py:311-537
        if args.script != "run_parameter_estimation":
            ap.error("`evaluate` belongs to run_parameter_estimation")
        res = estimation.evaluate(cfg.pop("filter_builder"), cfg.pop("solver_builder"), cfg.pop("ode_builder"), **cfg)
        if output is not None:
            os.makedirs(os.path.dirname(os.path.abspath(output)), exist_ok=True)
            np.savez(output, **res)
        return res
    cfg.pop("num_param_evals", None)
    if args.script == "run_parameter_estimation_baseline":   # scripts/run_parameter_estimation_baseline.py:40-262
        res = estimation.optimize_baseline(cfg.pop("solver_builder"), cfg.pop("ode_builder"), **cfg)
        if output is not None:
            os.makedirs(os.path.dirname(os.path.abspath(output)), exist_ok=True)
            np.savez(output, **res)
        return res
    res = estimation.optimize(cfg.pop("filter_builder"), cfg.pop("solver_builder"), cfg.pop("ode_builder"), **cfg)
    if output is not None:
        os.makedirs(os.path.dirname(os.path.abspath(output)), exist_ok=True)
        np.savez(output, **res)
    return res


if __name__ == "__main__":
    main(sys.argv[1:])
