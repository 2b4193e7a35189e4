#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ncu --set full --clock-control none --import-source on -k regex:ekf_thread_sched_kernel -s 1 -c 1 -f -o gpurun_out/lor_factor python tools/profile_c2.py 2000 Lorenz > gpurun_out/lor_factor.log 2>&1
ncu -i gpurun_out/lor_factor.ncu-rep --page source --csv > gpurun_out/lor_factor_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/lor_factor_source.csv
