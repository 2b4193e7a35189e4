#!/bin/bash
cd "$GRAFT_REPO_ROOT"
echo "== new"; timeout 300 python tools/bench_c3.py 4096 1000 2>&1 | tail -1
echo "== base"; ODEU_LIB=build/ab/libodeu_base.so timeout 300 python tools/bench_c3.py 4096 1000 2>&1 | tail -1
echo "== new grad"; timeout 300 python tools/bench_c3.py 4096 200 --grad 2>&1 | tail -2
echo "== base grad"; ODEU_LIB=build/ab/libodeu_base.so timeout 300 python tools/bench_c3.py 4096 200 --grad 2>&1 | tail -2
TAG=rows7 bash tools/r2_gpujob6.sh > /dev/null 2>&1
