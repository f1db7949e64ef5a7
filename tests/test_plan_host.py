"""Host-side exhaustive check of the flat-sweep work decomposition (csrc/plan.h + the piece/batch
iterators of csrc/kernels_nchw.cuh): tests/host/plan_check.cu is compiled for the host and run."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_partition_covers_every_vector_exactly_once(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "plan_check")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "plan_check.cu")],
                   check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "0 failures" in res.stdout


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_forward_queue_order_and_single_kernel_plans(tmp_path):
    """tests/host/queue_check.cu: the ordered statistics/apply queue of the one-kernel forwards visits every item once, every
    apply item after all statistics items of its channel (the no-deadlock precondition), `window` channels later; window /
    ring / resident plans tile their planes within the shared-memory, mbarrier and workspace limits."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "queue_check")
    res = subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                          os.path.join(ROOT, "tests", "host", "queue_check.cu")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:]
    assert "0 failures" in res.stdout and not res.stdout.startswith("0 cases")
