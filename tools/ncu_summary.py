"""Text summary of an `ncu --set full` report: per captured launch the metrics DESIGN.md quotes (duration, DRAM bytes and
throughput, tensor-pipe activity, L1/L2 throughput, occupancy, issue utilisation, registers, shared memory, top stall reasons).
    python tools/ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = [f"# {rep}"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    out.append(f"== {name[:150]}")
    for k in KEYS:
        if k in hdr:
            out.append(f"   {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in STALL), reverse=True)[:5]
    out.append("   top stall reasons (warps per issue-active cycle): " +
               ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
