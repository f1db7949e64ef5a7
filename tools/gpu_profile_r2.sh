#!/bin/bash
# round-2 evidence for profiles/: ncu launch list of the bench command, ncu --set full of the forward and backward kernels,
# forward-path timings, config lines, sanitizers (run under gpurun, 1 GPU)
out=gpurun_out/${1:-prof_r2}
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file $out/launches.csv python bench.py --steps 10 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_pair|bwd_nchw" -s 12 -c 2 -o $out/bench_kernels python bench.py --steps 3 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $out/ncu_bench.log 2>&1
timeout 600 python tools/cluster_bench.py --variants default,two_pass,window,resident,pair_p0,pair_p1,pair_p2,pair_p3,pair_p4,pair_p8,pair_p16,cluster_cs1_p1,cluster_cs2_p1,cluster_cs2_p2 > $out/fwd_paths.jsonl 2> $out/fwd_paths.err
timeout 600 python bench.py --config 3 --steps 100 --warmup 5 > $out/bench_config3.json 2> $out/cfg.err
timeout 900 python bench.py --config 5 --steps 20 --warmup 3 > $out/bench_config5.json 2>> $out/cfg.err
timeout 900 python tools/sweep.py --points "32,16,96,96,f32,nchw;32,16,192,192,f32,nchw;32,1,192,192,f32,nchw;20,64,224,224,f32,nchw;20,64,224,224,f32,nhwc;20,64,224,224,bf16,nchw" > $out/configs.jsonl 2>> $out/cfg.err
timeout 900 python tests/loop_config2_ref.py > $out/loop_config2_ref.txt 2>> $out/cfg.err
timeout 900 python tests/loop_config2.py > $out/loop_config2.txt 2>> $out/cfg.err
bash tools/gpu_sanitize.sh ${1:-prof_r2}_san > $out/sanitize.log 2>&1
tail -12 $out/sanitize.log
ls -la $out
