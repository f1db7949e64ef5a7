#!/bin/bash
out=gpurun_out/${1:-r2c}
mkdir -p $out
V=window,pair_p1,pair_p2,pair_p3,pair_p4,pair_p6
S="20,64,224,224,f32;64,64,112,112,f32;64,32,512,512,f32;20,64,224,224,bf16"
MAXSTYLE_PAIR_MINB=4 timeout 300 python tools/cluster_bench.py --variants $V,pair_p16 --shapes "$S" > $out/fwd_minb4.jsonl 2> $out/fwd.err
MAXSTYLE_PAIR_MINB=3 timeout 300 python tools/cluster_bench.py --variants $V,pair_p16 --shapes "$S" > $out/fwd_minb3.jsonl 2>> $out/fwd.err
tail -3 $out/fwd.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_pair -s 4 -c 1 -o $out/pair_p2 python tools/cluster_bench.py --shapes "20,64,224,224,f32" --variants pair_p2 --iters 3 > $out/ncu1.log 2>&1
MAXSTYLE_PAIR_MINB=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_pair -s 4 -c 1 -o $out/pair_p2_minb3 python tools/cluster_bench.py --shapes "20,64,224,224,f32" --variants pair_p2 --iters 3 > $out/ncu2.log 2>&1
tail -2 $out/ncu1.log $out/ncu2.log
ls -la $out
