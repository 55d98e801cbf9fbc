#!/bin/bash
# ncu --set full captures: sparse pair-gather conv at the stress shape (layers 3..6) and the p2 neck conv (3x3 128->128 @200x176)
OUT=gpurun_out/${1:-ncu2}
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_cg_kernel -s 0 -c 4 -f -o $OUT/prof_stress_spconv_cg python scripts/kernel_rooflines.py --shape stress --iters 1 > $OUT/ncu_stress_cg.log 2>&1; echo "ncu stress cg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bev_conv_p2_kernel -s 3 -c 1 -f -o $OUT/prof_p2 python scripts/p2_debug.py timing > $OUT/ncu_p2.log 2>&1; echo "ncu p2 rc=$?"
ls -la $OUT
