#!/bin/bash
# full GPU suite + bench + forward path timings (run under gpurun)
out=gpurun_out/${1:-r2a}
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
tail -15 $out/pytest_gpu.txt
timeout 600 python tools/cluster_bench.py --variants default,two_pass,window,resident,pair_p0,pair_p1,pair_p2,pair_p3,pair_p4,pair_p6,pair_p8,pair_p12,pair_p16 > $out/fwd_paths.jsonl 2> $out/fwd_paths.err
echo "fwd paths rc=$?"; tail -3 $out/fwd_paths.err
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
echo "bench rc=$?"; cat $out/bench.json | cut -c1-3000; tail -5 $out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench.err
cat $out/bench_ref.json | cut -c1-1200
