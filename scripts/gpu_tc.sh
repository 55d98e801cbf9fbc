TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 240 python scripts/tc_debug.py > $OUT/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> $OUT/tc_debug.log; grep -A11 "variant 3" $OUT/tc_debug.log | tail -13
timeout 120 python scripts/conv_bench.py 3 2>&1 | grep -v "^   \|^---\|ablate" | head -9 | tee $OUT/conv_bench_v3.log
