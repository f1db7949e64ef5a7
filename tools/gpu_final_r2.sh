#!/bin/bash
# round-2 evidence visit (1 GPU): full GPU suite, smoke, both bench arms, ncu launch list + --set full of the step's two kernels,
# forward paths, configs 3 / 5, sweep points, the real-solver loop, cross entropy, sanitizers
out=gpurun_out/${1:-final_r2}
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu.txt
grep -E "passed|failed|FAILED|rc=" $out/pytest_gpu.txt | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.txt
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > $out/bench_ref.json 2>> $out/bench.err
cut -c1-400 $out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > $out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fwd_pair|bwd_nchw" -s 12 -c 2 -f -o $out/prof python bench.py --steps 3 --warmup 5 --no-e2e --no-cpu-baseline --no-parity > $out/ncu_bench.log 2>&1
python tools/pm_series.py $out/prof.ncu-rep > $out/pm_series_bench.txt 2>&1
timeout 600 python tools/cluster_bench.py --variants default,two_pass,window,resident,pair_p0,pair_p1,pair_p2,pair_p3,pair_p4,pair_p8,pair_p16 > $out/fwd_paths.jsonl 2> $out/fwd_paths.err
timeout 600 python bench.py --config 3 --steps 100 --warmup 5 > $out/bench_config3.json 2> $out/cfg.err
timeout 900 python bench.py --config 5 --steps 20 --warmup 3 > $out/bench_config5.json 2>> $out/cfg.err
timeout 900 python tools/sweep.py --points "32,16,96,96,f32,nchw;32,16,192,192,f32,nchw;32,1,192,192,f32,nchw;20,64,224,224,f32,nchw;20,64,224,224,f32,nhwc;20,64,224,224,bf16,nchw;64,64,112,112,f32,nchw" > $out/configs.jsonl 2>> $out/cfg.err
timeout 900 python tests/loop_config2_ref.py > $out/loop_config2_ref.txt 2>> $out/cfg.err
bash tools/gpu_ce.sh ${1:-final_r2}_ce > $out/ce.log 2>&1; cp gpurun_out/${1:-final_r2}_ce/ce_timing.txt $out/ce2d.txt
bash tools/gpu_sanitize.sh ${1:-final_r2}_san > $out/sanitize.log 2>&1
tail -12 $out/sanitize.log
ls -la $out
