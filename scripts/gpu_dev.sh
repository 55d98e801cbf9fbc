#!/bin/bash
# usage: bash scripts/gpu_dev.sh <tag> "<pytest -k expr>" [roofline]  -- development round (each step under its own timeout)
TAG=${1:-dev}
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$2" ]; then
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
fi
if [ -n "$3" ]; then
timeout 600 python scripts/kernel_rooflines.py --shape frame > $OUT/roofline_frame.json 2> $OUT/roofline_frame.err; echo "roof frame rc=$?"; tail -3 $OUT/roofline_frame.err
timeout 900 python scripts/kernel_rooflines.py --shape stress > $OUT/roofline_stress.json 2> $OUT/roofline_stress.err; echo "roof stress rc=$?"; tail -3 $OUT/roofline_stress.err
timeout 600 python bench.py --steps 10 --warmup 3 --quick > $OUT/bench_quick.json 2> $OUT/bench.err; cat $OUT/bench_quick.json; tail -3 $OUT/bench.err
fi
