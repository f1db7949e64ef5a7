#!/bin/bash
# full GPU suite, smoke, default bench (both arms), then the round-2 evidence capture
out=gpurun_out/${1:-r2k}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.txt
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench.err
cut -c1-600 $out/bench.json
bash tools/gpu_profile_r2.sh ${1:-r2k}_prof
