#!/bin/bash
# usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]   -- selected GPU tests + bench + launch list
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "${2:-sparse or engine}" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_frame.py --frames 2 --cloud ${CLOUD:-ring} > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python scripts/summarize_launches.py $OUT/launches.csv 2>&1 | head -30
