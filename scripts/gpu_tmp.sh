#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_post_engine.py -m gpu -q --timeout 300 -x > $OUT/pytest_sparse.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sparse.log
tail -4 $OUT/pytest_sparse.log
timeout 600 python scripts/kernel_rooflines.py --shape stress --iters 3 > $OUT/roof_stress.json 2> $OUT/roof_stress.err; echo "roof stress rc=$?"; tail -2 $OUT/roof_stress.err
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 > $OUT/roof_frame.json 2> $OUT/roof_frame.err; echo "roof frame rc=$?"
python - <<PY
import json
for sh in ("stress","frame"):
    try:
        d=json.load(open("$OUT/roof_%s.json"%sh))
    except Exception as e:
        print(sh, "no json", e); continue
    print(sh, "total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:").replace("rulebook:","rb:"),g["ms"]) for g in d["groups"] if not g["group"].startswith(("neck"))))
PY
