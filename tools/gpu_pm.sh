#!/bin/bash
# DRAM throughput time series (ncu PM sampling) of the paired forward / the backward; backward slice oversubscription
out=gpurun_out/${1:-pm1}
mkdir -p $out
S="20,64,224,224,f32"
cap() { tag=$1; shift; env "$@" timeout 300 ncu --section PmSampling --clock-control none -k regex:"fwd_pair" -s 8 -c 1 -o /tmp/$tag -f python tools/cluster_bench.py --shapes "$S" --variants pair_p0 --iters 10 > $out/$tag.log 2>&1; python tools/pm_series.py /tmp/$tag.ncu-rep > $out/$tag.series.txt 2>&1; }
cap st3000 X=0
cap st0 MAXSTYLE_PAIR_STAGGER_NS=-1
cap st4500 MAXSTYLE_PAIR_STAGGER_NS=4500
for o in 1 2 3 4 6; do
  echo "== MAXSTYLE_BWD_OVERSUB=$o" >> $out/bwd.txt
  MAXSTYLE_BWD_OVERSUB=$o python tools/kernel_bench.py --sweeps "2,3,4" --iters 50 2>>$out/err.txt | head -1 | cut -c1-300 >> $out/bwd.txt
  MAXSTYLE_BWD_OVERSUB=$o python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>>$out/err.txt | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print(d['ms_per_step'], d['step_roofline']['fwd_ms'], d['step_roofline']['bwd_ms'], d['roofline']['frac'], d['parity']['ok'])" >> $out/bwd.txt
done
MAXSTYLE_BWD_OVERSUB=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_neighbours.py tests/test_gpu_graphed.py -q -x 2>&1 | tail -2 >> $out/bwd.txt
MAXSTYLE_BWD_OVERSUB=3 timeout 300 ncu --section PmSampling --clock-control none -k regex:"bwd_nchw" -s 8 -c 1 -o /tmp/bwd3 -f python tools/kernel_bench.py --sweeps "2,3,4" --iters 10 > $out/bwd3.log 2>&1; python tools/pm_series.py /tmp/bwd3.ncu-rep > $out/bwd3.series.txt 2>&1
cat $out/bwd.txt
