#!/bin/bash
# compute-sanitizer on the single-kernel forwards and the backward: memcheck, racecheck, synccheck.
# Small shapes only: the tools slow kernels down 10-100x.
out=gpurun_out/${1:-san}
mkdir -p $out
K="flag_variants or replays or (matches_two_pass and (shape4 or shape5 or shape6 or shape8))"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python -m pytest tests/test_gpu_cluster_fwd.py -q -x -k "$K" > $out/$tool.txt 2>&1
  echo "$tool rc=$?" | tee -a $out/$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out/$tool.txt | tail -3
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -q -x -k "golden_forward_backward or fused_forward" > $out/memcheck_parity.txt 2>&1
echo "memcheck parity rc=$?" | tee -a $out/memcheck_parity.txt
grep -E "ERROR SUMMARY|passed|failed" $out/memcheck_parity.txt | tail -3
