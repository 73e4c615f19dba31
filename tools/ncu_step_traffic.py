"""Per-kernel DRAM traffic and duration over ONE whole step from an
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
        --clock-control none --profile-from-start off --csv --log-file <csv> python tools/profile_step.py
capture: every launch of the step is measured (not a sample of a few launches), so `traffic` in bench.py's roofline
block is the mean over exactly the launches whose live duration `achieved` averages.  Writes profiles/ncu_traffic.json
(bytes per launch; the HALO and kx-in-N conv variants are folded into conv3x3_kernel, the name bench.py reports) and prints a table.
    python tools/ncu_step_traffic.py <csv> [<out.txt>] [<workload key of profiles/ncu_traffic.json, default base>]"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
per = collections.defaultdict(dict)          # launch id -> {metric: value, name}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    rec = per[row["ID"]]
    rec["name"] = row["Kernel Name"]
    rec[row["Metric Name"]] = v * scale.get(row["Metric Unit"], 1.0)


def short(name):
    n = re.sub(r"\(.*", "", re.sub(r"<.*", "", name)).replace("void ", "").strip().split("::")[-1]
    return "conv3x3_kernel" if n == "conv3x3_halo_kernel" else n


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])   # launches, us, dram bytes, tensor% x us
for rec in per.values():
    a = agg[short(rec["name"])]
    us = rec.get("gpu__time_duration.sum", 0.0)
    a[0] += 1
    a[1] += us
    a[2] += rec.get("dram__bytes_read.sum", 0.0) + rec.get("dram__bytes_write.sum", 0.0)
    a[3] += rec.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * us
tot = sum(a[1] for a in agg.values())
wl = sys.argv[3] if len(sys.argv) > 3 else "base"
out = [f"# one step of workload '{wl}' (base = 8 clips 432x240 5+3, b1 = one clip), every launch measured; total {tot / 1e3:.2f} ms serialised over "
       f"{sum(a[0] for a in agg.values())} launches"]
traffic = {}
for name, (n, us, b, tp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if not name.startswith(("at", "elementwise", "vectorized")) and "enable_if" not in name:
        traffic[name] = b / n
    out.append(f"{name[:34]:34s} launches={n:4d} total={us / 1e3:8.3f} ms ({100 * us / tot:4.1f}%) dram={b / n / 1e6:9.2f} MB/launch "
               f"({b / max(us, 1e-9) / 1e3:7.1f} GB/s) tensor-pipe active={tp / max(us, 1e-9):5.1f}%")
# bench.py times the conv kernels (tap-table / halo / kx-in-N) under ONE op kind: its traffic entry is the mean over all
# of them; the table above keeps the kx-in-N kernel on its own line
if "conv_kxn_kernel" in agg and "conv3x3_kernel" in agg:
    a, k = agg["conv3x3_kernel"], agg["conv_kxn_kernel"]
    traffic["conv3x3_kernel"] = (a[2] + k[2]) / (a[0] + k[0])
print("\n".join(out))
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("\n".join(out) + "\n")
jpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
old = json.load(open(jpath)) if os.path.exists(jpath) else {}
workload = sys.argv[3] if len(sys.argv) > 3 else "base"       # bench.py reads the entry of the workload it measures
old.setdefault(workload, {}).update(traffic)
json.dump(old, open(jpath, "w"), indent=1)
