#!/bin/bash
out=gpurun_out/${1:-cfg5}
mkdir -p $out
N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29641 bench.py --gpus $N --config 5 --steps 20 --warmup 3 > $out/bench_config5.json 2> $out/bench.err; echo "rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$out/bench_config5.json") if l.startswith("{")][-1])
print("config5", d["value"], d["ms_per_step"], d["step_roofline"]["frac_of_peak"], d["parity"], d["config"].get("one_kernel_forward"))
PY
tail -5 $out/bench.err
