#!/bin/bash
# GPU visit for the NHWC kernels: parity tests, per-kernel timings in both layouts, the config-4 sweep.
set -u
TAG=${1:-nhwc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== kernel bench nhwc"; timeout 300 python tools/kernel_bench.py --layout nhwc --sweeps "2,2,4;0,0,0" 2>&1 | tee $OUT/kernel_bench_nhwc.txt | head -20
echo "== kernel bench nhwc bf16"; timeout 300 python tools/kernel_bench.py --layout nhwc --dtype bf16 --sweeps "2,2,4" 2>&1 | tee $OUT/kernel_bench_nhwc_bf16.txt | head -20
echo "== sweep"; timeout 900 python tools/sweep.py ${SWEEP_ARGS:---quick} --out $OUT/sweep.jsonl 2>&1 | tail -80
