#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_pf_ext.py tests/test_pf_stream.py tests/test_parity_gpu.py tests/test_integration_reference.py tests/test_guard_reference.py -m gpu -x -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest_gpu.log
tail -15 gpurun_out/r2c_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print(d["value"], d["e2e"]["value"], d["roofline"]["frac"])
for k, v in d["extras"].items():
    print(k, v)
PY
