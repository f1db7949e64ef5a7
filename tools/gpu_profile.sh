#!/bin/bash
# Round profile visit: all GPU tests, smoke, bench (both arms), ncu launch list + full captures, named configs, forward-path comparison.
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.json
echo "== forward paths"
for spec in "20,64,224,224 f32" "20,64,224,224 bf16" "64,64,112,112 f32" "32,16,192,192 f32" "64,256,112,112 bf16"; do
  set -- $spec
  for sw in "2,3,4" "130,3,4" "546,3,4" "354,3,4" "18,3,4"; do   # default | resident | TMA ring | L2 window | two-pass
    timeout 120 python tools/kernel_bench.py --fwd-only --shape $1 --dtype $2 --sweeps "$sw" --iters 50 2>&1 | tail -1 | sed "s/^{/{\"stats_sweep\": \"$sw\", /" | tee -a $OUT/fwd_paths.txt
  done
done
echo "== named configs (3 per-GPU layers, 5 per GPU)"
timeout 600 python tools/sweep.py --points "32,16,96,96,f32,NCHW;32,16,192,192,f32,NCHW;32,1,192,192,f32,NCHW;256,32,512,512,f32,NCHW;20,64,224,224,f32,NCHW;20,64,224,224,f32,NHWC;20,64,224,224,bf16,NCHW" --out $OUT/configs.jsonl 2>&1 | tail -8
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1
tail -1 $OUT/launches.log | cut -c1-300
echo "== ncu full (bench: window forward + backward)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_|bwd_nchw' -s 8 -c 2 \
    -f -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/prof.log 2>&1
tail -1 $OUT/prof.log | cut -c1-200
echo "== ncu full (two-pass kernels + resident forward)"
MAXSTYLE_SWEEP="18,3,4" timeout 900 ncu --set full --clock-control none --import-source on -k regex:'stats_nchw|apply_nchw|tables_kernel' -s 9 -c 3 \
    -f -o $OUT/prof2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/prof2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fwd_resident' -s 8 -c 1 \
    -f -o $OUT/prof3 python tools/kernel_bench.py --fwd-only --dtype bf16 --sweeps "2,3,4" --iters 10 > $OUT/prof3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nhwc' -s 12 -c 4 \
    -f -o $OUT/prof4 python tools/kernel_bench.py --layout nhwc --sweeps "2,2,4" --iters 6 > $OUT/prof4.log 2>&1
echo "== config 2 loop"; timeout 600 python tests/loop_config2.py --width 64 2>&1 | tail -1 | tee $OUT/loop_config2.txt
timeout 600 python tests/loop_config2.py --width 16 2>&1 | tail -1 | tee -a $OUT/loop_config2.txt
echo "== config 4 sweep (full grid)"; timeout 1500 python tools/sweep.py --max-gb 8 --iters 7 --out $OUT/sweep.jsonl 2>&1 | tail -4
ls -la $OUT
