"""Summarise one or more .ncu-rep captures (ncu --set full) per kernel: duration, DRAM bytes per launch, tensor-pipe
and memory-pipe utilisation.  Writes profiles/ncu_traffic.json (dram bytes per launch, read by bench.py) and prints
a table.   python tools/ncu_traffic.py <out.txt> a.ncu-rep [b.ncu-rep ...]"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_txt, reps = sys.argv[1], sys.argv[2:]
want = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
agg = collections.defaultdict(list)
for rep in reps:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r[col["Kernel Name"]])).replace("void ", "").strip().split("::")[-1]
        rec = {}
        for m, k in want.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = float(r[col[m]].replace(",", ""))
                rec[k] = v * scale.get(units[col[m]], 1)
        agg[name].append(rec)
lines = []
traffic = {}
for name, recs in agg.items():
    def avg(k):
        vals = [x[k] for x in recs if k in x]
        return sum(vals) / len(vals) if vals else float("nan")
    dram = avg("dram_rd") + avg("dram_wr")
    traffic[name] = dram
    lines.append(f"{name:24s} launches={len(recs):2d} dur={avg('dur_us'):9.1f} us dram={dram / 1e6:9.2f} MB/launch "
                 f"dram%={avg('dram_pct'):5.1f} tensor%={avg('tensor_pct'):5.1f} l1tex%={avg('l1tex_pct'):5.1f} "
                 f"l2%={avg('l2_pct'):5.1f} warps%={avg('warps_pct'):5.1f} regs={avg('regs'):.0f} grid={avg('grid'):.0f}")
open(out_txt, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(traffic)
json.dump(old, open(path, "w"), indent=1)
