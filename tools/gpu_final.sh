#!/bin/bash
# End-of-round visit: all GPU tests, smoke, bench (both arms), ncu launch list + full capture of the step's two kernels.
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -6 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.json | cut -c1-200
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1
tail -1 $OUT/launches.log | cut -c1-200
echo "== ncu full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fwd_|bwd_nchw' -s 8 -c 2 \
    -f -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
tail -1 $OUT/prof.log | cut -c1-200
ls $OUT
