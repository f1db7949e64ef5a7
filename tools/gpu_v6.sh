#!/bin/bash
set -u
TAG=${1:-v6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_loop.py -m gpu -q -x --tb=long 2>&1 > $OUT/pytest_loop_full.txt; tail -5 $OUT/pytest_loop_full.txt
timeout 900 python -m pytest tests/test_gpu_loop.py -m gpu -q -x --tb=long -k "executor" 2>&1 > $OUT/pytest_exec_full.txt; tail -5 $OUT/pytest_exec_full.txt
timeout 600 python tests/loop_config2.py --width 16 > $OUT/loop16.txt 2>&1; tail -3 $OUT/loop16.txt
