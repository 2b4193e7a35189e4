#!/bin/bash
cd "$GRAFT_REPO_ROOT"
TAG=${TAG:-rows7}
ncu --set full --clock-control none --import-source on -k regex:ekf_rows_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_grad python tools/bench_c3.py 4096 50 --grad > gpurun_out/${TAG}_grad.log 2>&1
ncu -i gpurun_out/${TAG}_grad.ncu-rep --page raw --csv > gpurun_out/${TAG}_grad_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_grad.ncu-rep --page source --csv > gpurun_out/${TAG}_grad_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
tail -3 gpurun_out/${TAG}_grad.log
