#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
P2_STEP_TIMEOUT=60 timeout 900 python scripts/p2_debug.py > $OUT/p2_debug.log 2>&1; grep -E "SUMMARY|auto|FAIL" $OUT/p2_debug.log
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_bev.py tests/test_gpu_post_engine.py -m gpu -q --timeout 300 -x > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
timeout 600 python scripts/kernel_rooflines.py --shape stress --iters 3 > $OUT/roof_stress.json 2> $OUT/roof_stress.err
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 > $OUT/roof_frame.json 2> $OUT/roof_frame.err
python - <<PY
import json
for sh in ("stress","frame"):
    d=json.load(open("$OUT/roof_%s.json"%sh))
    print(sh, "total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:"),g["ms"]) for g in d["groups"] if g["group"].startswith(("conv","neck"))))
PY
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra > $OUT/bench.json 2> $OUT/bench.err; python -c "
import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','stages_ms')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['parity_vs_oracle']['pass'], d['parity_vs_oracle'].get('max_rel_score_diff'))"
