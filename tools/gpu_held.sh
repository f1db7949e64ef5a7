#!/bin/bash
# HISTORY (round 2): produced profiles/r02_fwd_experiments.txt.  The MAXSTYLE_PAIR_DEBUG / _READ_CAP / _STAGGER_GROUP knobs these runs
# used were experiments and have been removed from the library again; MAXSTYLE_PAIR_STAGGER_NS / _MINB / _PIECE_KB / _ORDER remain.
out=gpurun_out/${1:-held}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cluster_fwd.py tests/test_gpu_parity.py tests/test_gpu_fused_neighbours.py -q -x > $out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.txt
S="20,64,224,224,f32"
run() { echo "== $*" >> $out/held.txt; env "$@" python tools/cluster_bench.py --shapes "$S" --variants $V --iters 50 2>>$out/err.txt | cut -c1-330 >> $out/held.txt; }
V=pair_p0,pair_p2,pair_p3,pair_p4,pair_p6 run X=0
V=pair_p0,pair_p3,pair_p4 run MAXSTYLE_PAIR_STAGGER_NS=2000
V=pair_p0 run MAXSTYLE_PAIR_STAGGER_NS=1000
V=pair_p0 run MAXSTYLE_PAIR_STAGGER_NS=3000
V=pair_p0 run MAXSTYLE_PAIR_DEBUG=3
V=pair_p0,pair_p3 run MAXSTYLE_PAIR_MINB=3 MAXSTYLE_PAIR_STAGGER_NS=2000
python tools/cluster_bench.py --shapes "20,64,224,224,bf16;64,64,112,112,f32;32,16,192,192,f32;64,32,512,512,f32" --variants default,pair_p0,pair_p1,pair_p2,pair_p4,pair_p8,pair_p16,resident --iters 30 2>>$out/err.txt | cut -c1-330 >> $out/held.txt
cat $out/held.txt
