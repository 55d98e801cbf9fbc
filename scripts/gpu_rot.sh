#!/bin/bash
# de-phased weight streams: neck p2 timing (rot / norot), sparse cg rooflines (rot / norot), sparse + bev tests
OUT=gpurun_out/${1:-rot}
mkdir -p $OUT
timeout 300 python scripts/p2_debug.py timing > $OUT/p2_timing.log 2>&1; tail -20 $OUT/p2_timing.log
for R in 1 0; do
timeout 600 python scripts/kernel_rooflines.py --shape stress --iters 3 --rotate $R > $OUT/roof_stress_rot$R.json 2> $OUT/roof_stress_rot$R.err; echo "roof stress rot$R rc=$?"; tail -2 $OUT/roof_stress_rot$R.err
timeout 600 python scripts/kernel_rooflines.py --shape frame --iters 5 --rotate $R > $OUT/roof_frame_rot$R.json 2> $OUT/roof_frame_rot$R.err; echo "roof frame rot$R rc=$?"
python - <<PY
import json
for sh in ("stress","frame"):
    try:
        d=json.load(open("$OUT/roof_%s_rot$R.json"%sh))
    except Exception as e:
        print(sh, "no json", e); continue
    print(sh, "rot$R", "total_ms", d["total_ms"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:"),g["ms"]) for g in d["groups"] if g["group"].startswith(("conv","neck"))))
PY
done
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_bev.py -m gpu -q --timeout 300 -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
