#!/bin/bash
# GPU visit for the resident forward + host pipeline: tests first (bounded), then forward-only timings of the three paths.
set -u
TAG=${1:-res}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== resident/fused tests"; timeout 300 python -m pytest tests -m gpu -q -x -k "fused or resident or pipeline or nhwc_instance_stats" 2>&1 | tail -25 | tee $OUT/pytest_res.txt
for sw in "2,3,4" "98,3,4" "18,3,4"; do
  echo "== fwd-only sweeps=$sw"; timeout 120 python tools/kernel_bench.py --fwd-only --sweeps "$sw" --iters 50 2>&1 | tail -1 | tee -a $OUT/fwd_only.txt
done
for shp in "32,16,192,192" "64,64,112,112" "20,16,96,96"; do
  for sw in "2,3,4" "98,3,4" "18,3,4"; do
    echo "== fwd-only $shp sweeps=$sw"; timeout 120 python tools/kernel_bench.py --fwd-only --shape $shp --sweeps "$sw" --iters 50 2>&1 | tail -1 | tee -a $OUT/fwd_only.txt
  done
done
echo "== bf16"; for sw in "2,3,4" "98,3,4" "18,3,4"; do timeout 120 python tools/kernel_bench.py --fwd-only --dtype bf16 --sweeps "$sw" --iters 50 2>&1 | tail -1 | tee -a $OUT/fwd_only.txt; done
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== nhwc stats"; timeout 300 python tools/kernel_bench.py --layout nhwc --sweeps "2,2,4" 2>&1 | head -1 | tee $OUT/kernel_bench_nhwc.txt
