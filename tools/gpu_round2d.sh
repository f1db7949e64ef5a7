#!/bin/bash
out=gpurun_out/${1:-r2d}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cluster_fwd.py -x -q > $out/pytest_pair.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_pair.txt
tail -4 $out/pytest_pair.txt
V=window,pair_p1,pair_p2,pair_p3,pair_p4,pair_p5,pair_p6,pair_p8,pair_p12,pair_p16
S="20,64,224,224,f32;64,64,112,112,f32;64,32,512,512,f32;20,64,224,224,bf16;32,16,192,192,f32;32,16,96,96,f32"
MAXSTYLE_PAIR_MINB=3 timeout 300 python tools/cluster_bench.py --variants $V --shapes "$S" > $out/fwd_minb3.jsonl 2> $out/fwd.err
MAXSTYLE_PAIR_MINB=4 timeout 300 python tools/cluster_bench.py --variants $V --shapes "$S" > $out/fwd_minb4.jsonl 2>> $out/fwd.err
tail -3 $out/fwd.err
