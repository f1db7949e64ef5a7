#!/bin/bash
set -u
TAG=${1:-m4}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench --gpus $NG"
t0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $NG > $OUT/bench_n$NG.log 2>&1; echo "rc=$? wall=$(( $(date +%s) - t0 ))s"
grep '^{"metric"' $OUT/bench_n$NG.log | tail -1 | tee $OUT/bench_n$NG.json | cut -c1-1200
grep -n "Error\|error\|Traceback" $OUT/bench_n$NG.log | head -5
