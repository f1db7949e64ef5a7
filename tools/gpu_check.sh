#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the three streaming kernels.
# Usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1
tail -2 $OUT/launches.log
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stats_nchw|apply_nchw|bwd_nchw|fwd_|tables_kernel' -s 12 -c 4 \
    -f -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
tail -2 $OUT/prof.log
ls -la $OUT
