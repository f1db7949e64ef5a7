#!/bin/bash
set -u
TAG=${1:-v9}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_graphed.py -m gpu -q -x --tb=short 2>&1 | tail -25 | tee $OUT/pytest_graphed.txt
echo "== bench graphed"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench eager"; BENCH_EAGER=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_eager.json
echo "== bench graphed ring"; MAXSTYLE_RING=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_ring.json
