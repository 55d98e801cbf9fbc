#!/bin/bash
# usage: bash scripts/gpu_ncu.sh <tag>   -- ncu --set full captures of the two dominant kernels (one GPU)
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for K in bev_conv_tc3_kernel spconv_tc_kernel; do
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-2} -c 2 -f -o $OUT/prof_$K python scripts/profile_frame.py --frames 1 --cloud ring > $OUT/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
