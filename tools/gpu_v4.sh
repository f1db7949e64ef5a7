#!/bin/bash
set -u
TAG=${1:-v4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== forward paths (resident v2 = 130)"
for spec in "20,64,224,224 f32" "20,64,224,224 bf16" "64,64,112,112 f32" "32,16,192,192 f32" "64,256,112,112 bf16" "20,16,96,96 f32"; do
  set -- $spec
  for sw in "130,3,4" "546,3,4" "354,3,4"; do
    timeout 120 python tools/kernel_bench.py --fwd-only --shape $1 --dtype $2 --sweeps "$sw" --iters 50 2>&1 | tail -1 | sed "s/^{/{\"stats_sweep\": \"$sw\", /" | tee -a $OUT/fwd_paths.txt
  done
done
echo "== bench default"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench resident forced"; MAXSTYLE_SWEEP="130,3,4" timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_resident.json
echo "== bench ring"; MAXSTYLE_RING=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_ring.json
