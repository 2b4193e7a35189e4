#!/bin/bash
# ncu full + source capture of the C3 row kernels (loss, and loss + gradient)
cd "$GRAFT_REPO_ROOT"
TAG=${TAG:-rows5}
ncu --set full --clock-control none --import-source on -k regex:ekf_rows_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_nll python tools/bench_c3.py 4096 200 > gpurun_out/${TAG}_nll.log 2>&1
ncu -i gpurun_out/${TAG}_nll.ncu-rep --page raw --csv > gpurun_out/${TAG}_nll_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_nll.ncu-rep --page source --csv > gpurun_out/${TAG}_nll_source.csv 2>/dev/null
if [ -n "$GRAD" ]; then
ncu --set full --clock-control none --import-source on -k regex:ekf_rows_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_grad python tools/bench_c3.py 4096 50 --grad > gpurun_out/${TAG}_grad.log 2>&1
ncu -i gpurun_out/${TAG}_grad.ncu-rep --page raw --csv > gpurun_out/${TAG}_grad_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_grad.ncu-rep --page source --csv > gpurun_out/${TAG}_grad_source.csv 2>/dev/null
fi
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -n 6
