#!/bin/bash
OUT=gpurun_out/t29; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log; tail -6 $OUT/pytest.log
timeout 600 python scripts/kernel_rooflines.py --shape frame > $OUT/roof_frame.json 2> $OUT/roof_frame.err; echo "roof rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 1500 $OUT/bench.json
