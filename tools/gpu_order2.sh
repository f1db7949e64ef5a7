#!/bin/bash
# 2-GPU check of the global-cycle-order forward: parity at 2 ranks, then config 5 / config 1 bench lines
out=gpurun_out/${1:-order2}
mkdir -p $out
N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_cluster_fwd.py -q -x -k "order or cycle or replays" 2>&1 | tail -2
timeout 900 $TR --master-port 29631 tests/dist_parity.py > $out/dist_parity.txt 2>&1
echo "dist_parity rc=$?" | tee -a $out/dist_parity.txt
grep -E "Error|assert|Traceback|global_cycle" $out/dist_parity.txt | cut -c1-300 | head -6
timeout 900 $TR --master-port 29632 bench.py --gpus $N --config 5 --steps 20 --warmup 3 > $out/bench_config5.json 2> $out/bench.err; cut -c1-250 $out/bench_config5.json
python - <<PY
import json
d = json.loads([l for l in open("$out/bench_config5.json") if l.startswith("{")][-1])
print("config5", d["value"], d["ms_per_step"], d["step_roofline"]["frac_of_peak"], d["parity"]["ok"], d["config"].get("one_kernel_forward"))
PY
timeout 600 $TR --master-port 29633 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > $out/bench.json 2>> $out/bench.err; cut -c1-200 $out/bench.json
tail -3 $out/bench.err
