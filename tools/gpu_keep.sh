#!/bin/bash
set -u
TAG=${1:-keep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for mb in 0 32 64 96; do
  echo "== keep $mb MB"; MAXSTYLE_FUSED_KEEP_MB=$mb timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'keep_mb': $mb, 'ms_per_step': round(d['ms_per_step'],4), 'value': round(d['value']), 'fwd_ms': round(d['step_roofline']['fwd_ms'],4), 'bwd_ms': round(d['step_roofline']['bwd_ms'],4), 'bwd_frac': round(d['roofline']['frac'],3)}))" | tee -a $OUT/keep.txt
done
echo "== keep 64 ring"; MAXSTYLE_RING=1 timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-160
echo "== tests"; timeout 600 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | tail -4
