#!/bin/bash
# round-2 first contact: p2 kernel steps, GPU tests (all, no -x), bench
OUT=gpurun_out/${1:-r2a}
mkdir -p $OUT
python -c "import torch; print(torch.cuda.get_device_name(0))" > $OUT/gpu.txt 2>&1
P2_STEP_TIMEOUT=90 timeout 900 python scripts/p2_debug.py > $OUT/p2_debug.log 2>&1; echo "p2 rc=$?"; tail -60 $OUT/p2_debug.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --durations=10 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -40 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -c 4000 $OUT/bench.json; tail -5 $OUT/bench.err
