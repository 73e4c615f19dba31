"""SASS opcode census of the shipped library: per kernel, how many tcgen05 / TMA / TMEM instructions it contains
(B200_PROFILING.md "What proves a Blackwell-native kernel": tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
TMA -> UTMALDG/UTMASTG/UBLKCP, tcgen05.commit -> UTCBAR, legacy mma.sync -> HMMA).  No GPU needed.
    python tools/sass_census.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "e2fgvi_b200", "libe2fgvi_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "LDGSTS", "HMMA", "SYNCS", "ELECT"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("e2f::", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for op in OPS:
        if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line):
            per[cur][op] += 1
            break
lines = ["# SASS census of e2fgvi_b200/libe2fgvi_b200.so (sm_100a), instructions per kernel; " +
         "kernels with none of these opcodes are plain SIMT (gathers / elementwise)",
         f"{'kernel':70s} " + " ".join(f"{o:>8s}" for o in OPS)]
tot = collections.Counter()
for k, c in per.items():
    tot.update(c)
    if sum(c.values()):
        lines.append(f"{k[:70]:70s} " + " ".join(f"{c.get(o, 0):8d}" for o in OPS))
lines.append(f"{'TOTAL (' + str(len(per)) + ' kernels)':70s} " + " ".join(f"{tot.get(o, 0):8d}" for o in OPS))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(out + "\n")
