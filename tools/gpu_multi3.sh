#!/bin/bash
set -u
TAG=${1:-m3}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== bench --gpus 1"; timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
echo "== bench --gpus $NG"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $NG > $OUT/bench_n$NG.log 2>&1; tail -1 $OUT/bench_n$NG.log | tee $OUT/bench_n$NG.json | cut -c1-1500
grep -n "Error\|error" $OUT/bench_n$NG.log | head -10
echo "== bench --gpus $NG --impl reference"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $NG --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "== distributed test"; timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -3
