#!/bin/bash
# cluster-resident forward: parity tests, then timings of every forward path (run under gpurun)
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_cluster_fwd.py -x -q > gpurun_out/c1/pytest_cluster.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/c1/pytest_cluster.txt
tail -5 gpurun_out/c1/pytest_cluster.txt
timeout 600 python tools/cluster_bench.py > gpurun_out/c1/cluster_bench.jsonl 2> gpurun_out/c1/cluster_bench.err
echo "bench rc=$?"
cat gpurun_out/c1/cluster_bench.jsonl | cut -c1-260
tail -3 gpurun_out/c1/cluster_bench.err
