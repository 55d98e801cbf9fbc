#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 > $OUT/roof_frame.json 2> $OUT/roof_frame.err; echo "roof frame rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/roof_frame.json"))
print("frame total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:").replace("rulebook:","rb:"),g["ms"]) for g in d["groups"] if not g["group"].startswith(("neck","conv"))))
PY
timeout 600 python bench.py --steps 6 --warmup 3 --no-extra > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$OUT/bench.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','stages_ms','launches_per_batch')}, d['e2e']['value'], d['latency_single_batch']['median_ms'], d['parity_vs_oracle']['pass'], d['parity_vs_oracle'].get('max_rel_score_diff'))"
