#!/bin/bash
set -u
TAG=${1:-p2p4}
NG=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_parity.py > $OUT/dist_parity.log 2>&1; echo "parity rc=$? ok_ranks=$(grep -o 'dist_parity\] rank [0-9]*/[0-9]* ok' $OUT/dist_parity.log | wc -l)"
grep -n "Error\|assert" $OUT/dist_parity.log | head -5 | cut -c1-300
BENCH_EXCHANGE=auto timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --gpus $NG --no-e2e > $OUT/bench.log 2>&1; echo "bench rc=$?"
grep '^{"metric"' $OUT/bench.log | tail -1 | tee $OUT/bench_n$NG.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'n_gpus': d['n_gpus'], 'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'],4), 'fwd_ms': round(d['step_roofline']['fwd_ms'],4), 'bwd_ms': round(d['step_roofline']['bwd_ms'],4), 'exchange': d['config'].get('exchange')}))"
grep -n "Error\|Traceback" $OUT/bench.log | head -3
