#!/bin/bash
# ncu --set full of the paired forward on config-5 planes (64 x 32 x 512 x 512 fp32: 1 MiB planes, 16 pieces each)
out=gpurun_out/${1:-ncu_c5}
mkdir -p $out
timeout 400 ncu --set full --clock-control none -k regex:"fwd_pair" -s 3 -c 1 -f -o /tmp/c5 python tools/cluster_bench.py --shapes "64,32,512,512,f32" --variants default --iters 3 > $out/ncu.log 2>&1
python tools/ncu_summary.py /tmp/c5.ncu-rep --lines 6 > $out/ncu_config5_fwd.txt 2>&1
python tools/pm_series.py /tmp/c5.ncu-rep >> $out/ncu_config5_fwd.txt 2>&1
head -30 $out/ncu_config5_fwd.txt | cut -c1-200
