#!/bin/bash
# One GPU-box visit covering everything at HEAD: parity tests, smoke, bench (both arms), NHWC kernel timings, quick config-4 sweep.
set -u
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -30 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/bench_ref.json
echo "== kernel bench nchw"; timeout 300 python tools/kernel_bench.py --sweeps "2,3,4" 2>&1 | tail -3 | tee $OUT/kernel_bench_nchw.txt
echo "== kernel bench nhwc"; timeout 300 python tools/kernel_bench.py --layout nhwc --sweeps "2,2,4;0,0,0" 2>&1 | tee $OUT/kernel_bench_nhwc.txt | tail -6
echo "== kernel bench nhwc bf16"; timeout 300 python tools/kernel_bench.py --layout nhwc --dtype bf16 --sweeps "2,2,4" 2>&1 | tee $OUT/kernel_bench_nhwc_bf16.txt | tail -4
echo "== sweep"; timeout 600 python tools/sweep.py ${SWEEP_ARGS:---quick} --out $OUT/sweep.jsonl 2>&1 | tail -60
