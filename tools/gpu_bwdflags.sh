#!/bin/bash
set -u
OUT=gpurun_out/${1:-bwdflags}
mkdir -p $OUT
for sw in "2,3,4" "2,3,5" "2,3,12" "2,3,0" "2,3,13" "2,3,8" "0,3,4" "6,3,4"; do
  MAXSTYLE_FUSED_KEEP_MB=0 MAXSTYLE_SWEEP=$sw timeout 300 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'sweeps': '$sw', 'ms_per_step': round(d['ms_per_step'],4), 'fwd_ms': round(d['step_roofline']['fwd_ms'],4), 'bwd_ms': round(d['step_roofline']['bwd_ms'],4), 'bwd_frac': round(d['roofline']['frac'],3)}))" | tee -a $OUT/bwdflags.txt
done
