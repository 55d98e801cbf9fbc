TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 180 python scripts/tc_debug.py > $OUT/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> $OUT/tc_debug.log; grep -v "^ out" $OUT/tc_debug.log | tail -32
timeout 120 python scripts/conv_bench.py 2 2>&1 | head -9 | tee $OUT/conv_bench_v2.log
