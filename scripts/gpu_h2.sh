#!/bin/bash
OUT=gpurun_out/${1:-h2}
mkdir -p $OUT
timeout 300 python scripts/h2_bench.py > $OUT/h2_bench.log 2>&1; echo "h2_bench rc=$?"; grep -v "^accuracy" $OUT/h2_bench.log | tail -42
if [ -n "$2" ]; then
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "$2" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
fi
