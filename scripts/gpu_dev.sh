#!/bin/bash
# usage: bash scripts/gpu_dev.sh <tag> "<pytest -k expr>" "<sp_h2_debug args>" [roofline split] -- development round (each step under its own timeout)
TAG=${1:-dev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/sp_h2_debug.py $3 > $OUT/sp_h2_debug.log 2>&1; echo "sp_h2_debug rc=$?"; tail -60 $OUT/sp_h2_debug.log
if [ -n "$2" ]; then
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
fi
if [ -n "$4" ]; then
timeout 600 python scripts/kernel_rooflines.py --shape frame --sparse-split $4 > $OUT/roofline_frame_$4.json 2> $OUT/roofline_frame.err; echo "roof frame rc=$?"; tail -3 $OUT/roofline_frame.err
timeout 900 python scripts/kernel_rooflines.py --shape stress --sparse-split $4 > $OUT/roofline_stress_$4.json 2> $OUT/roofline_stress.err; echo "roof stress rc=$?"; tail -3 $OUT/roofline_stress.err
fi
