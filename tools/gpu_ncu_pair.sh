#!/bin/bash
out=gpurun_out/${1:-ncu_pair}
mkdir -p $out
for p in 2 4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwd_pair -s 4 -c 1 -o $out/pair_p$p python tools/cluster_bench.py --shapes "20,64,224,224,f32" --variants pair_p$p --iters 3 > $out/ncu_p$p.log 2>&1
done
ls -la $out
