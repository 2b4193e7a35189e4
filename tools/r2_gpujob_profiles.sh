#!/bin/bash
# round 2: full GPU suite, bench (both arms), ncu launch list and full captures of the dominant kernels
set -x
cd "$GRAFT_REPO_ROOT"
if [ -z "$SKIP_PYTEST" ]; then
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
fi
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench.err || exit 1
python bench.py --guard intended --no-cpu-baseline --steps 10 > gpurun_out/r2_bench_n1_intended.json 2>> gpurun_out/r2_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
for sys_ in Lorenz VanDerPol; do
  ncu --set full --clock-control none --import-source on -k regex:ekf_thread_sched_kernel -s 1 -c 1 -f -o gpurun_out/r2_${sys_}_sched python tools/profile_c2.py 10000 $sys_ > gpurun_out/r2_ncu_${sys_}.log 2>&1
  ncu -i gpurun_out/r2_${sys_}_sched.ncu-rep --page raw --csv > gpurun_out/r2_${sys_}_sched_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r2_${sys_}_sched_raw.csv gpurun_out/r2_${sys_}_sched_ncu_full.csv
done
ncu --set full --clock-control none --import-source on -k regex:ekf_thread_sched_kernel -s 1 -c 1 -f -o gpurun_out/r2_Lorenz_sched_intended python tools/profile_c2.py 10000 Lorenz --guard=intended > gpurun_out/r2_ncu_Lorenz_intended.log 2>&1
ncu -i gpurun_out/r2_Lorenz_sched_intended.ncu-rep --page raw --csv > gpurun_out/r2_Lorenz_sched_intended_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_Lorenz_sched_intended_raw.csv gpurun_out/r2_Lorenz_sched_intended_ncu_full.csv
# C3 row kernels (loss, loss + gradient): full capture + source-level shared-memory wavefront table
TAG=r2_c3_rows GRAD=1 bash tools/r2_gpujob6.sh > /dev/null 2>&1
for k in nll grad; do
  python tools/ncu_summary.py gpurun_out/r2_c3_rows_${k}_raw.csv gpurun_out/r2_c3_rows_${k}_ncu_full.csv
  python tools/ncu_lds.py gpurun_out/r2_c3_rows_${k}_source.csv > gpurun_out/r2_c3_rows_${k}_lds.txt
done
rm -f gpurun_out/r2_c3_rows_*_source.csv
rm -f gpurun_out/*_raw.csv gpurun_out/*.ncu-rep      # (the reports are 34 MB each: gpurun returns at most 64 MiB)
ls -la gpurun_out | tail -n 14
