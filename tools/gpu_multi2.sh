#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity of the global-batch layer + bench at 1 and N GPUs on the same box.
set -u
TAG=${1:-multi}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus.txt
nvidia-smi topo -m 2>&1 | head -12 | tee $OUT/topo.txt
echo "== distributed parity (NCCL, $NG ranks)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_parity.py 2>&1 | tail -6 | tee $OUT/dist_parity.txt
echo "== bench --gpus 1"; timeout 600 python bench.py --gpus 1 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench_n1.json
echo "== bench --gpus $NG"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $NG 2>&1 | tail -1 | tee $OUT/bench_n$NG.json
echo "== fwd breakdown at $NG ranks"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29543 \
    tools/dist_breakdown.py 2>&1 | tail -3 | tee $OUT/dist_breakdown.txt
