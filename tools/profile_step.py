"""One InpaintGenerator.forward (B clips of 432x240, 5+3) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off ...  python tools/profile_step.py [--clips B] [--precision strict|tf32]
Never a timing source (numbers under a profiler are not bench values)."""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200.synth import synth_frames, synth_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clips", type=int, default=8)
ap.add_argument("--precision", default="strict")
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()
dev = torch.device("cuda:0")
net = importlib.import_module("model.e2fgvi")
model = net.InpaintGenerator().eval()
model.load_state_dict(synth_state_dict(model, "default", 0))
model.to(dev)
model.precision = args.precision
x = synth_frames(args.clips, 8, 240, 432, seed=100).to(dev)
with torch.no_grad():
    for _ in range(args.warmup):
        model(x, 5)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x, 5)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one forward of", args.clips, "clips")
