#!/bin/bash
OUT=gpurun_out/${1:-tmp}
mkdir -p $OUT
for D in 0 1; do
timeout 600 python scripts/kernel_rooflines.py --shape frame-uniform --iters 5 --cg-deep $D > $OUT/roof_uniform_$D.json 2> $OUT/roof_uniform_$D.err; echo "roof rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/roof_uniform_$D.json"))
print("uniform deep=$D total_ms", d["total_ms"], d["active_sites"], " ".join("%s=%.3f"%(g["group"].replace("neck:","n:").replace("rulebook:","rb:"),g["ms"]) for g in d["groups"] if not g["group"].startswith("neck")), "neck=%.3f"%sum(g["ms"] for g in d["groups"] if g["group"].startswith("neck")))
PY
done
timeout 600 python scripts/cg_prof.py frame 0 > /dev/null 2>&1
