#!/bin/bash
# HISTORY (round 2): produced profiles/r02_fwd_experiments.txt.  The MAXSTYLE_PAIR_DEBUG / _READ_CAP / _STAGGER_GROUP knobs these runs
# used were experiments and have been removed from the library again; MAXSTYLE_PAIR_STAGGER_NS / _MINB / _PIECE_KB / _ORDER remain.
# experiment: do the CTAs of the paired forward run in lockstep?  start them out of phase / drop the waits (timing only)
out=gpurun_out/${1:-stagger}
mkdir -p $out
S="20,64,224,224,f32"
run() { echo "== $*" >> $out/stagger.txt; env "$@" python tools/cluster_bench.py --shapes "$S" --variants pair_p0 --iters 50 2>>$out/err.txt | cut -c1-200 >> $out/stagger.txt; }
run X=0
for ns in 500 1000 2000 3000 4500 6000 9000; do run MAXSTYLE_PAIR_STAGGER_NS=$ns; done
run MAXSTYLE_PAIR_DEBUG=1
run MAXSTYLE_PAIR_DEBUG=3
run MAXSTYLE_PAIR_DEBUG=3 MAXSTYLE_PAIR_STAGGER_NS=3000
run MAXSTYLE_PAIR_MINB=3
run MAXSTYLE_PAIR_MINB=3 MAXSTYLE_PAIR_STAGGER_NS=3000
run MAXSTYLE_PAIR_MINB=3 MAXSTYLE_PAIR_STAGGER_NS=6000
cat $out/stagger.txt
