#!/bin/bash
OUT=gpurun_out/t30; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_bev.py -q -x --timeout 150 > $OUT/pytest_bev.log 2>&1; rc=$?; echo "pytest bev rc=$rc" >> $OUT/pytest_bev.log; tail -12 $OUT/pytest_bev.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python scripts/kernel_rooflines.py --shape frame > $OUT/roof_frame.json 2> $OUT/roof_frame.err; echo "roof rc=$?"
timeout 300 python scripts/kernel_rooflines.py --shape frame --p2-derive 0 > $OUT/roof_frame_nd.json 2> $OUT/roof_frame_nd.err; echo "roof nd rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 600 $OUT/bench.json
