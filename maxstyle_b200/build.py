"""Build recipe for libmaxstyle_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m maxstyle_b200.build` compiles the library in-tree with nvcc; there is no JIT and no
torch extension involved -- the C ABI (include/maxstyle_b200.h) is bound with ctypes, so the
library does not link against torch at all.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libmaxstyle_b200.so")
SOURCES = ["capi.cu"]
HEADERS = ["common.cuh", "plan.h", "kernels_nchw.cuh", "kernels_nhwc.cuh", "fused_fwd.cuh", "resident_fwd.cuh", "ring_fwd.cuh", "cluster_fwd.cuh", "pair_fwd.cuh", "ce2d.cuh", "tables.cuh", os.path.join("..", "..", "include", "maxstyle_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",     # B200 only; no PTX for other targets, no fallback arch
    "-lineinfo", "-O3", "-std=c++17",
    "--compiler-options", "-fPIC", "-shared",
]


def find_nvcc() -> str:
    cand = [os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc"), shutil.which("nvcc") or ""]
    for c in cand:
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found (looked in $CUDA_HOME/bin, /usr/local/cuda/bin and PATH)")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the shared library if it is missing or older than its sources; return its path."""
    if not force and not is_stale():
        return LIB_PATH
    cmd = [find_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH + ".tmp"] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
