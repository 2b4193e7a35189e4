#!/bin/bash
set -x
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_per_trajectory_obs.py -m gpu -x -q > gpurun_out/r2d_pytest_obs.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest_obs.log
tail -15 gpurun_out/r2d_pytest_obs.log
for sys in Lorenz VanDerPol; do for g in reference intended; do
  timeout 300 python tools/ab_obs_stage.py --system $sys --guard $g >> gpurun_out/r2d_ab_obs.jsonl 2>> gpurun_out/r2d_ab_obs.err
  ODEU_NO_OBS_STAGE=1 timeout 300 python tools/ab_obs_stage.py --system $sys --guard $g >> gpurun_out/r2d_ab_obs.jsonl 2>> gpurun_out/r2d_ab_obs.err
done; done
cat gpurun_out/r2d_ab_obs.jsonl
tail -5 gpurun_out/r2d_ab_obs.err
