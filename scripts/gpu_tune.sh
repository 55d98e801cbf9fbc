#!/bin/bash
# quick throughput sweep: concurrent engines x sparse-conv pipeline depth (value only) -> profiles/r2e_tune_streams_deep.log was made with this
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
for D in 0 1; do for S in 8 12 16; do
echo -n "deep $D streams $S: "; timeout 300 python bench.py --quick --steps 4 --warmup 3 --streams $S --cg-deep $D 2>> $OUT/err.log | python -c "import sys,json; print(round(json.loads(sys.stdin.read().strip().splitlines()[-1])['value'],1))" | tee -a $OUT/tune.log
done; done
