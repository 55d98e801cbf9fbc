"""ncu launch-list CSV -> per-kernel table (count, total us, avg us, share)."""
import collections, csv, sys

def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"]
        k = k.split("(")[0].replace("void ", "")[:70]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        tot[k] = tot.get(k, 0) + v
        cnt[k] += 1
    return tot, cnt

if __name__ == "__main__":
    tot, cnt = load(sys.argv[1])
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    total = sum(tot.values())
    print("%-72s %5s %10s %9s %6s" % ("kernel", "n/fr", "us/frame", "avg us", "share"))
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print("%-72s %5.1f %10.1f %9.1f %5.1f%%" % (k, cnt[k] / frames, v / frames, v / cnt[k], 100 * v / total))
    print("%-72s %5.1f %10.1f" % ("TOTAL", sum(cnt.values()) / frames, total / frames))
