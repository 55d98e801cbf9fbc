"""Per-source-line warp-stall samples of an ncu --set full --import-source on capture (cuda,sass view).
    python scripts/ncu_source_hot.py REP [kernel-index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kern = -1
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = {}
fname = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1][:60]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        si = hdr.index("# Samples")
        ii = hdr.index("Instructions Executed")
        key = (func, fname, int(r[0]))
        cur = lines.setdefault(key, [0, 0, r[1].strip()[:110]])
        cur[0] += int(r[si]) if r[si].isdigit() else 0
        cur[1] += int(r[ii]) if r[ii].isdigit() else 0
tot = {}
for (f, fn, ln), v in lines.items():
    tot[f] = tot.get(f, 0) + v[0]
for f in tot:
    print("== %s   total samples %d" % (f, tot[f]))
    items = sorted([(v[0], v[1], fn, ln, v[2]) for (ff, fn, ln), v in lines.items() if ff == f], reverse=True)[:top]
    for smp, ins, fn, ln, src in items:
        print("  %6d (%4.1f%%) inst %9d  %s:%d  %s" % (smp, 100.0 * smp / max(1, tot[f]), ins, fn, ln, src))
