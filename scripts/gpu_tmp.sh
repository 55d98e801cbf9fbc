#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
for M in 0 2; do
echo "== mode $M (2 = no proxy fence)"
timeout 600 python scripts/cg_prof.py stress $M > $OUT/cg_prof_stress_$M.log 2>&1; tail -30 $OUT/cg_prof_stress_$M.log | grep -A1 "conv:[36] "
timeout 600 python scripts/cg_prof.py frame $M > $OUT/cg_prof_frame_$M.log 2>&1; tail -30 $OUT/cg_prof_frame_$M.log | grep -A1 "conv:[36] "
done
