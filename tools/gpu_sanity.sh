#!/bin/bash
# Last check of the tree: GPU tests, smoke, bench.
set -u
OUT=gpurun_out/${1:-sanity}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-260
