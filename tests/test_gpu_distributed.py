"""NCCL runs of the global-batch layer on 2 / 4 / 8 GPUs against the reference semantics on the concatenated batch
(tests/dist_parity.py: the reference-generated golden, and the float64 oracle on the shape and kernel path bench.py times).
Skipped on boxes with fewer GPUs; the gloo test covers the host logic."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.timeout(900)
def test_nccl_ranks_match_reference_on_concatenated_batch(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29577 + world), os.path.join(ROOT, "tests", "dist_parity.py")],
                         capture_output=True, text=True, timeout=850)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("[dist_parity] rank") == world
