#!/bin/bash
OUT=gpurun_out/t34; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $OUT/pytest.log; tail -4 $OUT/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 300 python scripts/kernel_rooflines.py --shape frame > $OUT/roof_frame.json 2> $OUT/roof_frame.err; echo "roof rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 200 $OUT/bench.json
