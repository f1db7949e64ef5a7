#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): distributed parity test + the bench at 1..N GPUs.
set -u
TAG=${1:-multi}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
echo "== nhwc + fused tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 -k "nhwc or fused or golden" 2>&1 | tail -8 | tee $OUT/pytest_sel.txt
echo "== distributed parity (NCCL, $NG ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_parity.py 2>&1 | tail -12 | tee $OUT/dist_parity.txt
for n in 1 $NG; do
  echo "== bench --gpus $n"
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 \
        bench.py --gpus $n 2>&1 | tail -1 | tee $OUT/bench_n$n.json
  fi
done
