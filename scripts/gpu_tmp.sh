#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_post_engine.py -m gpu -q --timeout 300 -x > $OUT/pytest_sparse.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sparse.log
tail -6 $OUT/pytest_sparse.log
for M in 0 4; do
timeout 600 python scripts/kernel_rooflines.py --shape stress --iters 3 --cg-l1 $M > $OUT/roof_stress_$M.json 2> $OUT/roof_stress_$M.err; echo "roof stress rc=$?"; tail -2 $OUT/roof_stress_$M.err
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 --cg-l1 $M > $OUT/roof_frame_$M.json 2> $OUT/roof_frame_$M.err; echo "roof frame rc=$?"
python - <<PY
import json
for sh in ("stress","frame"):
    try:
        d=json.load(open("$OUT/roof_%s_$M.json"%sh))
    except Exception as e:
        print(sh, "no json", e); continue
    print(sh, "mode $M total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:").replace("rulebook:","rb:"),g["ms"]) for g in d["groups"] if g["group"].startswith(("conv"))))
PY
done
