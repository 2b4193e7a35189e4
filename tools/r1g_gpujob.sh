set -x
python tools/write_bw.py > gpurun_out/r1g_write_bw.log 2>&1
python bench.py > gpurun_out/r1g_bench_n1.json 2> gpurun_out/r1g_bench.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1g_bench_reference_arm.json 2>> gpurun_out/r1g_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1g_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ekf_thread_sched_kernel -s 1 -c 1 -f -o gpurun_out/r1g_lorenz_sched python tools/profile_c2.py 10000 Lorenz > gpurun_out/r1g_ncu_lorenz.log 2>&1
ncu -i gpurun_out/r1g_lorenz_sched.ncu-rep --page raw --csv > gpurun_out/r1g_lorenz_sched_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:ekf_thread_sched_kernel -s 1 -c 1 -f -o gpurun_out/r1g_vdp_sched python tools/profile_c2.py 10000 VanDerPol > gpurun_out/r1g_ncu_vdp.log 2>&1
ncu -i gpurun_out/r1g_vdp_sched.ncu-rep --page raw --csv > gpurun_out/r1g_vdp_sched_raw.csv 2>/dev/null
ls -la gpurun_out | tail -n 12
