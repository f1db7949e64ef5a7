#!/bin/bash
# Quick GPU visit: parity tests (all failures shown), per-kernel timings, host overhead, bench line.
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke under memcheck"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 | tee $OUT/memcheck.txt
echo "== kernel bench"; timeout 600 python tools/kernel_bench.py --host ${KB_ARGS:-} 2>&1 | tee $OUT/kernel_bench.txt | head -60
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -2 | tee $OUT/bench.json
