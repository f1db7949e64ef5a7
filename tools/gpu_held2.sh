#!/bin/bash
# HISTORY (round 2): produced profiles/r02_fwd_experiments.txt.  The MAXSTYLE_PAIR_DEBUG / _READ_CAP / _STAGGER_GROUP knobs these runs
# used were experiments and have been removed from the library again; MAXSTYLE_PAIR_STAGGER_NS / _MINB / _PIECE_KB / _ORDER remain.
out=gpurun_out/${1:-held2}
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cluster_fwd.py tests/test_gpu_parity.py -q -x > $out/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.txt
S="20,64,224,224,f32"
run() { echo "== $*" >> $out/held.txt; env "$@" python tools/cluster_bench.py --shapes "$S" --variants $V --iters 50 2>>$out/err.txt | cut -c1-330 >> $out/held.txt; }
V=pair_p0,pair_p3 run X=0
V=pair_p0 run MAXSTYLE_PAIR_DEBUG=4
V=pair_p0 run MAXSTYLE_PAIR_DEBUG=8
V=pair_p0,pair_p4 run MAXSTYLE_PAIR_DEBUG=8 MAXSTYLE_PAIR_STAGGER_NS=3000
V=pair_p0 run MAXSTYLE_PAIR_DEBUG=4 MAXSTYLE_PAIR_STAGGER_NS=3000
for d in 3000 4000 5000 6000 7000 8000 10000 12000; do V=pair_p0 run MAXSTYLE_PAIR_STAGGER_GROUP=2 MAXSTYLE_PAIR_STAGGER_NS=$d; done
V=pair_p0,pair_p3 run MAXSTYLE_PAIR_STAGGER_NS=3000
V=pair_p0 run MAXSTYLE_PAIR_STAGGER_NS=3500
V=pair_p0 run MAXSTYLE_PAIR_STAGGER_NS=2500
V=pair_p0 run MAXSTYLE_PAIR_STAGGER_GROUP=2 MAXSTYLE_PAIR_STAGGER_NS=6000 MAXSTYLE_PAIR_DEBUG=4
cat $out/held.txt
