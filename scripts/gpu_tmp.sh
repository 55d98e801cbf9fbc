#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
P2_STEP_TIMEOUT=60 timeout 900 python scripts/p2_debug.py > $OUT/p2_debug.log 2>&1; grep -E "SUMMARY|TFLOP|FAIL" $OUT/p2_debug.log
timeout 900 python -m pytest tests/test_gpu_bev.py tests/test_gpu_post_engine.py tests/test_compat.py -m gpu -q --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','stages_ms','latency_single_batch')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['parity_vs_oracle']['pass'], d['parity_vs_oracle'].get('max_rel_score_diff'))"
