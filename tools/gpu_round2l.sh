#!/bin/bash
out=gpurun_out/${1:-r2l}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -15
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cut -c1-300 $out/bench.json
timeout 600 python tools/cluster_bench.py --variants default,pair_p0,pair_p1,pair_p2,pair_p3,resident,window > $out/fwd_paths.jsonl 2> $out/fwd_paths.err
timeout 900 python tests/loop_config2_ref.py > $out/loop_config2_ref.txt 2> $out/loop.err
cut -c1-1500 $out/loop_config2_ref.txt
