OUT=gpurun_out/${1:-exp}
mkdir -p $OUT
timeout 180 python scripts/tc_debug.py 2>&1 | grep "E1" | tee $OUT/e1.log
for S in 1 2 4 8; do timeout 300 python bench.py --quick --steps 6 --warmup 2 --streams $S 2>&1 | tail -1 | tee -a $OUT/streams.log; done
