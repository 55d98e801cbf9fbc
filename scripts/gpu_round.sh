#!/bin/bash
# usage: bash scripts/gpu_round.sh <tag>   -- tests + smoke + bench (both arms) + ncu launch list + ncu full captures + per-kernel rooflines
# logs under gpurun_out/<tag>/ ; SKIP_TESTS=1 / SKIP_NCU=1 / SKIP_ROOF=1 to shorten
TAG=${1:-r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 700 $OUT/bench_ref.json
if [ -z "$SKIP_ROOF" ]; then
timeout 600 python scripts/kernel_rooflines.py --shape frame > $OUT/roofline_frame.json 2> $OUT/roofline_frame.err; echo "roof frame rc=$?"; tail -3 $OUT/roofline_frame.err
timeout 900 python scripts/kernel_rooflines.py --shape stress > $OUT/roofline_stress.json 2> $OUT/roofline_stress.err; echo "roof stress rc=$?"; tail -3 $OUT/roofline_stress.err
fi
if [ -z "$SKIP_NCU" ]; then
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_frame.py --frames 2 --cloud ${CLOUD:-ring} > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for K in ${NCU_KERNELS:-bev_conv_p2_kernel spconv_cg_kernel}; do
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o $OUT/prof_$K python scripts/profile_frame.py --frames 1 --cloud ${CLOUD:-ring} > $OUT/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
done
fi
if [ -n "$NCU_STRESS" ]; then
# stress-shape captures (BASELINE config #5 shape on one GPU), first pass of scripts/kernel_rooflines.py:
#  (a) the tensor-core sparse conv: launches 0..3 = layers 3, 4 (32->32), 5 (32->64), 6 (64->64)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spconv_cg_kernel -s 0 -c 4 -f -o $OUT/prof_stress_spconv_cg python scripts/kernel_rooflines.py --shape stress --iters 1 > $OUT/ncu_stress.log 2>&1; echo "ncu stress rc=$?"
#  (b) the bandwidth-bound kernels: voxeliser, rulebook, narrow-layer conv, split, dense(), NMS mask
timeout 900 ncu --set full --clock-control none -k regex:"vox_|hash_build|nbr_kernel|mark_outputs|enumerate_kernel|tile_lists|spconv_rows|dense_gather|post_mask" -c 16 -f -o $OUT/prof_stress_hbm python scripts/kernel_rooflines.py --shape stress --iters 1 > $OUT/ncu_stress_hbm.log 2>&1; echo "ncu stress hbm rc=$?"
fi
