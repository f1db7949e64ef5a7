#!/bin/bash
# Rank start barrier experiment: one-kernel multi-GPU forward with / without the barrier, against the three-call p2p path.
set -u
TAG=${1:-bar}
NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 \
    tests/dist_parity.py > $OUT/dist_parity.log 2>&1; echo "parity rc=$? ok_ranks=$(grep -o 'dist_parity\] rank [0-9]*/[0-9]* ok' $OUT/dist_parity.log | wc -l)"
grep -n "Error\|assert" $OUT/dist_parity.log | head -4 | cut -c1-300
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  MAXSTYLE_ONE_KERNEL=$1 MAXSTYLE_START_BARRIER=$2 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29542 \
      bench.py --gpus $NG --no-e2e > $OUT/bench_$1$2.log 2>&1; rc=$?
  grep '^{"metric"' $OUT/bench_$1$2.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'one_kernel': $1, 'barrier': $2, 'rc': $rc, 'n_gpus': d['n_gpus'], 'value': round(d['value']), 'ms_per_step': round(d['ms_per_step'],4), 'fwd_ms': round(d['step_roofline']['fwd_ms'],4), 'bwd_ms': round(d['step_roofline']['bwd_ms'],4)}))" | tee -a $OUT/bar.txt
done
