"""NCCL run of the global-batch layer on 2 GPUs against the reference-generated golden of the concatenated
batch (tests/dist_parity.py).  Skipped on boxes with fewer than 2 GPUs; the gloo test covers the host logic."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.timeout(600)
def test_two_rank_nccl_matches_reference_on_concatenated_batch():
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "dist_parity.py")],
                         capture_output=True, text=True, timeout=550)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("[dist_parity] rank") == 2
