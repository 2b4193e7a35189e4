#!/bin/bash
cd "$GRAFT_REPO_ROOT"
echo "== loss"; timeout 300 python tools/bench_c3.py 4096 1000 2>&1 | tail -1
echo "== grad"; timeout 300 python tools/bench_c3.py 4096 200 --grad 2>&1 | tail -2
echo "== single"; timeout 300 python tools/bench_c3.py 4096 1000 --single 2>&1 | tail -1
echo "== single grad"; timeout 300 python tools/bench_c3.py 4096 200 --single --grad 2>&1 | tail -2
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_grad.py tests/test_per_trajectory_obs.py tests/test_estimation.py tests/test_baseline_loss.py tests/test_guard_reference.py -m gpu -x -q 2>&1 | tail -3
