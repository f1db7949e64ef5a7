#!/bin/bash
set -u
OUT=gpurun_out/${1:-tune}
mkdir -p $OUT
for w in 32; do for p in 32 64 128; do for ch in 1 2 3 4; do
MAXSTYLE_FUSED_CHUNK=$ch MAXSTYLE_FUSED_WINDOW_MB=$w MAXSTYLE_FUSED_PIECE_KB=$p python tools/kernel_bench.py --fwd-only --sweeps "2,3,4" 2>&1 | tail -1 | tee -a $OUT/tune.txt
done; done; done
echo "== ncu fused"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fwd_fused' -s 6 -c 1 -f -o $OUT/prof_fused python tools/kernel_bench.py --fwd-only --iters 4 --sweeps "2,3,4" > $OUT/prof.log 2>&1
tail -2 $OUT/prof.log
echo "== fused tests"; timeout 600 python -m pytest tests -m gpu -q -k "fused" --maxfail=20 2>&1 | tail -5
