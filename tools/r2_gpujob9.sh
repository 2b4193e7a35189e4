#!/bin/bash
cd "$GRAFT_REPO_ROOT"
N=${NG:-2}
echo "== 1 GPU"; timeout 300 python tools/bench_c4.py 1000000 1000 --bootstrap --force-resample 2>&1 | grep -E "C4|loglik"
echo "== $N GPUs"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_c4.py 1000000 1000 --bootstrap --force-resample 2>&1 | grep -E "C4|loglik"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n$N.json"))
print(d["value"], d["e2e"]["value"], d["n_gpus"])
for k in ("c3_scale", "c4_scale"):
    print(k, {a: b for a, b in d["extras"].get(k, {}).items() if a != "sample"})
PY
tail -3 gpurun_out/r2_bench_n$N.err
