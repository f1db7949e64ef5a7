#!/bin/bash
out=gpurun_out/${1:-last}
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.txt
timeout 400 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cut -c1-260 $out/bench.json
timeout 300 python tools/sweep.py --points "32,16,96,96,f32,NCHW;32,16,192,192,f32,NCHW;32,1,192,192,f32,NCHW;20,64,224,224,f32,NCHW;20,64,224,224,f32,NHWC;20,64,224,224,bf16,NCHW;64,64,112,112,f32,NCHW" > $out/configs.jsonl 2> $out/cfg.err; cut -c1-330 $out/configs.jsonl
