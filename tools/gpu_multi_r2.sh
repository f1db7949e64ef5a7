#!/bin/bash
# N-GPU round-2 run (under gpurun --gpus N): parity on the benchmarked path, bench line, link ceiling, configs 3 and 5.
N=${1:-2}
out=gpurun_out/${2:-multi$N}
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 900 $TR --master-port 29611 tests/dist_parity.py > $out/dist_parity.txt 2>&1
echo "dist_parity rc=$?" | tee -a $out/dist_parity.txt
grep -c "\[dist_parity\] rank" $out/dist_parity.txt; grep -E "Error|assert|Traceback" $out/dist_parity.txt | head -5
timeout 900 $TR --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; cut -c1-1500 $out/bench.json; tail -3 $out/bench.err
[ "${3:-full}" = "full" ] && MAXSTYLE_ONE_KERNEL=0 timeout 900 $TR --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-parity > $out/bench_three_call.json 2>> $out/bench.err
[ "${3:-full}" = "full" ] && BENCH_EXCHANGE=nccl timeout 900 $TR --master-port 29614 bench.py --gpus $N --steps 100 --warmup 5 --no-e2e --no-parity > $out/bench_nccl.json 2>> $out/bench.err
python - <<PY
import json
for f in ("bench", "bench_three_call", "bench_nccl"):
    try:
        d = json.loads([l for l in open("$out/" + f + ".json") if l.startswith("{")][-1])
        print(f, round(d["value"]), "samples/s", round(d["ms_per_step"] * 1e3, 1), "us/step fwd", round(d["step_roofline"]["fwd_ms"] * 1e3, 1), "bwd", round(d["step_roofline"]["bwd_ms"] * 1e3, 1), d.get("parity", {}).get("ok"), d["config"].get("exchange", "")[:60])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 $TR --master-port 29615 tools/pcie_ceiling.py > $out/pcie_ceiling.json 2>> $out/bench.err; cat $out/pcie_ceiling.json
timeout 600 $TR --master-port 29616 bench.py --gpus $N --config 3 --steps 100 --warmup 5 > $out/bench_config3.json 2>> $out/bench.err; cut -c1-900 $out/bench_config3.json
timeout 900 $TR --master-port 29617 bench.py --gpus $N --config 5 --steps 20 --warmup 3 > $out/bench_config5.json 2>> $out/bench.err; cut -c1-900 $out/bench_config5.json
tail -5 $out/bench.err
