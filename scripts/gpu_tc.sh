TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 python scripts/tc_debug.py > $OUT/tc_debug.log 2>&1; echo "tc_debug rc=$?" >> $OUT/tc_debug.log; cat $OUT/tc_debug.log | tail -25
timeout 600 python -m pytest tests/test_gpu_bev.py -q --timeout 120 > $OUT/pytest_bev.log 2>&1; echo "rc=$?" >> $OUT/pytest_bev.log; tail -30 $OUT/pytest_bev.log
