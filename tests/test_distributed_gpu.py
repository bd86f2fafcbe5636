"""Column-slab partition on >= 2 GPUs (peer-memory and NCCL halos) against the one-GPU
operator; runs tests/dist_peer_check.py under torchrun.  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_slab_partition_matches_one_gpu(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "dist_peer_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "PEER CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
