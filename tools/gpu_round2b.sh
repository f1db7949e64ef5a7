#!/bin/bash
out=gpurun_out/${1:-r2b}
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_reference_callers.py tests/test_gpu_cluster_fwd.py -q -s -k "reference or unet or fcn64 or generate or default_prob or cycle or flag or replays" > $out/pytest_callers.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_callers.txt
grep -E "passed|failed|seed|n_iter|Error|assert" $out/pytest_callers.txt | tail -60
V=default,window,pair_p1,pair_p2,pair_p3,pair_p4,pair_p8,pair_p16
S="20,64,224,224,f32;20,64,224,224,bf16;64,64,112,112,f32;32,16,192,192,f32;64,32,512,512,f32"
MAXSTYLE_PAIR_OOL=1 timeout 300 python tools/cluster_bench.py --variants $V --shapes "$S" > $out/fwd_ool.jsonl 2> $out/fwd.err
MAXSTYLE_PAIR_OOL=2 timeout 300 python tools/cluster_bench.py --variants $V --shapes "$S" > $out/fwd_inline.jsonl 2>> $out/fwd.err
tail -3 $out/fwd.err
