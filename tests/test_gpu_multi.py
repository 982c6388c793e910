"""GPU, two ranks over NCCL (skips on a one-GPU box): the row-sharded sweep with the delta all-reduce inside the C ABI
(msb_state_allreduce_deltas) leaves bit-identical replicas, equals the torch.distributed-carried collective, and
equals ONE GPU sweeping all the rows.  The checks themselves live in scripts/check_replicas.py (also run by hand
under `gpurun --gpus 2`; its output is kept under profiles/)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_two_rank_sweep_over_nccl_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "check_replicas.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=850)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "all ok" in out.stdout


@pytest.mark.gpu
def test_nccl_is_reachable_through_the_abi():
    from common_b200 import dist as cbd
    assert cbd.nccl_version() >= 22000
