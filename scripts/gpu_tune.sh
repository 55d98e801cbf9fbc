#!/bin/bash
# quick throughput sweep: streams x sparse-conv pipeline depth (value only)
OUT=gpurun_out/${1:-tune}; mkdir -p $OUT
for D in 0 1; do for S in 6 8 12 16; do
echo -n "depth $D streams $S: "; timeout 200 python bench.py --quick --steps 8 --warmup 3 --streams $S --sp-h2-depth $D 2>> $OUT/err.log | tee -a $OUT/tune.log
done; done
