#!/bin/bash
# spconv_cg bring-up: sparse tests, then per-group rooflines with the pair-gather kernel (both cache policies) next to the TMA-gather kernel
OUT=gpurun_out/${1:-cg}
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sparse.py -m gpu -q --timeout 300 -x > $OUT/pytest_sparse.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_sparse.log
tail -30 $OUT/pytest_sparse.log
for V in "cg 0" "cg 1" "h2 0"; do set -- $V
timeout 600 python scripts/kernel_rooflines.py --shape stress --iters 3 --sparse-tc $1 --cg-l1 $2 > $OUT/roof_stress_$1_$2.json 2> $OUT/roof_stress_$1_$2.err; echo "roof stress $V rc=$?"; tail -2 $OUT/roof_stress_$1_$2.err
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 --sparse-tc $1 --cg-l1 $2 > $OUT/roof_frame_$1_$2.json 2> $OUT/roof_frame_$1_$2.err; echo "roof frame $V rc=$?"; tail -2 $OUT/roof_frame_$1_$2.err
python - <<PY
import json
for sh in ("stress","frame"):
    try:
        d=json.load(open("$OUT/roof_%s_$1_$2.json"%sh))
    except Exception as e:
        print(sh, "no json", e); continue
    print(sh, "$V", "total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"],g["ms"]) for g in d["groups"] if g["group"].startswith(("conv","split"))))
PY
done
