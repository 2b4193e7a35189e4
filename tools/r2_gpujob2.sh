#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_guard_reference.py tests/test_per_trajectory_obs.py tests/test_full_size.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest_gpu.log
tail -15 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --guard reference --no-cpu-baseline > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err; echo "bench ref exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --guard intended --no-cpu-baseline > gpurun_out/r2b_bench_intended.json 2> gpurun_out/r2b_bench_intended.err; echo "bench int exit $?"
python - <<'PY'
import json
for m in ("reference", "intended"):
    try:
        d = json.load(open(f"gpurun_out/r2b_bench_{m}.json"))
        print(m, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["guard"], d["extras"].get("lorenz_traj_steps_per_s"), d["extras"].get("vdp_traj_steps_per_s"))
    except Exception as e:
        print(m, "failed", e)
PY
