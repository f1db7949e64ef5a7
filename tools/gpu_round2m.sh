#!/bin/bash
out=gpurun_out/${1:-r2m}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -8
timeout 300 python tools/kernel_bench.py --sweeps "2,3,4" --iters 30 --host > $out/kernel_bench.txt 2>&1; grep host_us $out/kernel_bench.txt
timeout 300 python tools/kernel_bench.py --shape 20,16,192,192 --sweeps "2,3,4" --iters 30 --host 2>&1 | grep host_us >> $out/kernel_bench.txt; tail -1 $out/kernel_bench.txt
BENCH_EAGER=1 timeout 600 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline > $out/bench_eager.json 2> $out/bench.err; cut -c1-200 $out/bench_eager.json
timeout 900 python tests/loop_config2_ref.py > $out/loop_config2_ref.txt 2> $out/loop.err
cut -c1-700 $out/loop_config2_ref.txt
timeout 600 python bench.py > $out/bench.json 2>> $out/bench.err; cut -c1-200 $out/bench.json
