#!/bin/bash
set -u
TAG=${1:-v5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== loop/executor tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "loop or executor" 2>&1 | tail -30 | tee $OUT/pytest_loop.txt
echo "== config 2 loop"; timeout 600 python tests/loop_config2.py --width 64 2>&1 | tail -1 | tee $OUT/loop_config2.txt
timeout 600 python tests/loop_config2.py --width 16 2>&1 | tail -1 | tee -a $OUT/loop_config2.txt
echo "== all tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
