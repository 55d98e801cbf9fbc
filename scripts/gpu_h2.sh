#!/bin/bash
OUT=gpurun_out/${1:-h2}
mkdir -p $OUT
timeout 300 python scripts/h2_bench.py > $OUT/h2_bench.log 2>&1; echo "h2_bench rc=$?"; tail -20 $OUT/h2_bench.log
