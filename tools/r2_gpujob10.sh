#!/bin/bash
cd "$GRAFT_REPO_ROOT"
N=${NG:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm_n$N.json 2>> gpurun_out/r2_bench_n$N.err; echo "reference arm exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n$N.json"))
print(d["value"], d["e2e"]["value"], d["n_gpus"], d["roofline"]["frac"])
for k in ("c3_scale", "c4_scale"):
    print(k, {a: b for a, b in d["extras"].get(k, {}).items() if a != "sample"})
print(open("gpurun_out/r2_bench_reference_arm_n$N.json").read()[:300])
PY
tail -3 gpurun_out/r2_bench_n$N.err
