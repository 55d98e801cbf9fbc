OUT=gpurun_out/${1:-exp}
mkdir -p $OUT
timeout 180 python scripts/tc_debug.py 2>&1 | grep "E1" | tee $OUT/e1.log
timeout 900 python scripts/stress_bench.py 2>&1 | tail -3 | tee $OUT/stress.log
