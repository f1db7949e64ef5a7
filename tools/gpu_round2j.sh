#!/bin/bash
out=gpurun_out/${1:-r2j}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -15
bash tools/gpu_ce.sh ${1:-r2j}_ce > $out/ce.log 2>&1; grep fwd_bwd $out/ce.log
