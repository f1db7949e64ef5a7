#!/bin/bash
# GPU visit for the streamed (TMA ring) forward: parity tests, then forward-only timings against the other paths and over the ring knobs.
set -u
TAG=${1:-ring}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== tests"; timeout 600 python -m pytest tests -m gpu -q -x -k "fused or resident or mixstyle or loop" 2>&1 | tail -25 | tee $OUT/pytest_ring.txt
RING=544   # NO_RESIDENT | FORCE_RING
for spec in "20,64,224,224 f32" "20,64,224,224 bf16" "64,64,112,112 f32" "32,16,192,192 f32"; do
  set -- $spec
  for sw in "$((RING+2)),3,4" "$((2+32+256+64)),3,4" "130,3,4"; do
    timeout 120 python tools/kernel_bench.py --fwd-only --shape $1 --dtype $2 --sweeps "$sw" --iters 50 2>&1 | tail -1 | sed "s/^{/{\"stats_sweep\": \"$sw\", /" | tee -a $OUT/fwd_paths.txt
  done
done
echo "== ring knobs (config 1 fp32)"
for st in 3 4 6; do for pc in 2 4 8 14; do
  MAXSTYLE_RING_STAGES=$st MAXSTYLE_RING_PIECE_CHUNKS=$pc timeout 120 python tools/kernel_bench.py --fwd-only --sweeps "$((RING+2)),3,4" --iters 50 2>&1 | tail -1 | sed "s/^{/{\"stages\": $st, \"piece_chunks\": $pc, /" | tee -a $OUT/ring_knobs.txt
done; done
for wm in 16 48 64; do
  MAXSTYLE_FUSED_WINDOW_MB=$wm timeout 120 python tools/kernel_bench.py --fwd-only --sweeps "$((RING+2)),3,4" --iters 50 2>&1 | tail -1 | sed "s/^{/{\"window_mb_set\": $wm, /" | tee -a $OUT/ring_knobs.txt
done
echo "== config 2 loop"; timeout 600 python tests/loop_config2.py --width 64 2>&1 | tail -1 | tee $OUT/loop_config2.txt
timeout 600 python tests/loop_config2.py --width 16 2>&1 | tail -1 | tee -a $OUT/loop_config2.txt
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q --maxfail=20 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/bench.json
