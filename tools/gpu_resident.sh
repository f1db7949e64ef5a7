#!/bin/bash
set -u
OUT=gpurun_out/${1:-fused}
mkdir -p $OUT
echo "== fused tests"; timeout 600 python -m pytest tests -m gpu -q -k "fused" --maxfail=20 2>&1 | tail -15
echo "== kernel bench"; timeout 600 python tools/kernel_bench.py --sweeps "2,3,4" 2>&1 | tail -1 | tee -a $OUT/kernel_bench.txt
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== two-pass"; MAXSTYLE_SWEEP="18,3,4" timeout 600 python tools/kernel_bench.py --sweeps "18,3,4" 2>&1 | tail -1 | tee -a $OUT/kernel_bench.txt
echo "== racecheck smoke"; timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
