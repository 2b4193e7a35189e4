#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/bench_c4.py 1000000 5000 2>&1 | grep "C4 particle"
timeout 300 python tools/bench_c4.py 1000000 1000 --bootstrap --force-resample 2>&1 | grep "C4"
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_pf_stream.py tests/test_pf_ext.py tests/test_implicit.py -m gpu -x -q 2>&1 | tail -2
