"""Bootstrap filter over NVLink peer memory (particle_filter_ext._bootstrap_fused): on a box with at
least two GPUs, two ranks run the peer-memory formulation (symmetric buffers, remote stores / loads, no NCCL on the data
path) and the NCCL all-gather formulation of the same job and must agree bit for bit on every rank - particles, weights,
ESS history, resampling events, log-likelihood - with forced resampling and with the ESS rule (tools/pf_peer_check.py
asserts it).  EXTENSION: no reference oracle exists for the weight / resampling steps."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_memory_path_is_bit_identical_to_the_all_gather_path():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "tools", "pf_peer_check.py"), "200000"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("ess_frac=")]
    assert len(lines) == 2 and all("bit-identical on all ranks: True" in l for l in lines), r.stdout[-2000:]
