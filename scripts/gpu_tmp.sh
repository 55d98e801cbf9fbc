#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 300 python scripts/p2_prof.py 0 > $OUT/p2_prof.log 2>&1; cat $OUT/p2_prof.log | head -24
