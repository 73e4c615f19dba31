"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    k = re.sub(r"<.*", "", row["Kernel Name"])[:80]
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print(f"# total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches (serialised, cold-cache: compare shares)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t / 1e3:9.3f} ms {100 * t / tot:5.1f}% {n:5d}  {k}")
