#!/bin/bash
set -u
TAG=${1:-v3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench (window default)"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench (ring)"; MAXSTYLE_RING=1 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_ring.json
echo "== config 2 loop"; timeout 600 python tests/loop_config2.py --width 64 2>&1 | tail -1 | tee $OUT/loop_config2.txt
timeout 600 python tests/loop_config2.py --width 16 2>&1 | tail -1 | tee -a $OUT/loop_config2.txt
