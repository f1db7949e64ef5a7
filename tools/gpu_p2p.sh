#!/bin/bash
set -u
TAG=${1:-p2p}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== dist parity incl. graphed p2p / nccl"
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_parity.py > $OUT/dist_parity.log 2>&1; echo "rc=$?"; grep "dist_parity\|Error\|error\|assert" $OUT/dist_parity.log | head -12 | cut -c1-900
for ex in p2p nccl; do
  echo "== bench --gpus $NG exchange=$ex"
  BENCH_EXCHANGE=$ex timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
      bench.py --gpus $NG --no-e2e > $OUT/bench_$ex.log 2>&1; echo "rc=$?"
  grep '^{"metric"' $OUT/bench_$ex.log | tail -1 | tee $OUT/bench_$ex.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'n_gpus': d['n_gpus'], 'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'],4), 'fwd_ms': round(d['step_roofline']['fwd_ms'],4), 'bwd_ms': round(d['step_roofline']['bwd_ms'],4), 'exchange': d['config'].get('exchange')}))"
  grep -n "Error\|Traceback" $OUT/bench_$ex.log | head -3
done
echo "== bench --gpus 1"; timeout 200 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'n_gpus': d['n_gpus'], 'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'],4)}))"
