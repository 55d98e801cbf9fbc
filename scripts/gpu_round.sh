#!/bin/bash
# usage: bash scripts/gpu_round.sh <tag>     -- tests + smoke + short bench + ncu launch list, logs under gpurun_out/<tag>/
TAG=${1:-r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 600 $OUT/bench_ref.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python scripts/profile_frame.py --frames 2 --cloud ${CLOUD:-ring} > $OUT/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
if [ -n "$NCU_KERNEL" ]; then
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -c ${NCU_COUNT:-3} -f -o $OUT/prof_$NCU_KERNEL python scripts/profile_frame.py --frames 1 --cloud ${CLOUD:-ring} > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
