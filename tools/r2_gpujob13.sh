#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ncu --set full --clock-control none --import-source on -k regex:pf_thread_kernel -s 1 -c 1 -f -o gpurun_out/pf_pred python tools/bench_c4.py 1000000 1000 > gpurun_out/pf_pred.log 2>&1
ncu -i gpurun_out/pf_pred.ncu-rep --page raw --csv > gpurun_out/pf_pred_raw.csv 2>/dev/null
ncu -i gpurun_out/pf_pred.ncu-rep --page source --csv > gpurun_out/pf_pred_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/pf_pred_raw.csv gpurun_out/r2_c4_pf_thread_ncu_full.csv
rm -f gpurun_out/*.ncu-rep
tail -2 gpurun_out/pf_pred.log
