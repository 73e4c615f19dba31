"""Which torch-native (non-e2fgvi_b200) CUDA kernels still run inside one forward, with the Python line that issued
them.  Diagnostic only (torch.profiler); never a timing source."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from e2fgvi_b200.synth import synth_frames, synth_state_dict  # noqa: E402

dev = torch.device("cuda:0")
net = importlib.import_module("model.e2fgvi")
model = net.InpaintGenerator().eval()
model.load_state_dict(synth_state_dict(model, "default", 0))
model.to(dev)
x = synth_frames(8, 8, 240, 432, seed=100).to(dev)
with torch.no_grad():
    for _ in range(2):
        model(x, 5)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        model(x, 5)
        torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_time_total <= 0 or not ev.name.startswith("aten::") or ev.cpu_children and any(
            c.name.startswith("aten::") and c.device_time_total > 0 for c in ev.cpu_children):
        continue
    where = next((s for s in ev.stack if "/e2fgvi_b200/" in s or "/model/" in s), "?")
    rows.append((ev.device_time_total, ev.name, str(ev.input_shapes)[:90], where.replace(ROOT + "/", "")[:110]))
agg = {}
for t, name, shp, where in rows:
    k = (name, shp, where)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"TRACE torch-native device time {tot / 1e3:.2f} ms in {sum(a[0] for a in agg.values())} ops")
for (name, shp, where), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"TRACE {t:8.1f} us n={n:3d} {name:28s} {shp:90s} {where}")
