"""FusionFeedForward middle at the HQ 720p shape (8 frames, 40 channels, 180x324 -> 6480 tokens x 1960)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
tok = torch.randn(8, 60 * 108, 1960, device=dev)
geo = ((7, 7), (3, 3), (3, 3))
fn = lambda: ops.t2t_fold_unfold(tok, (180, 324), *geo, gelu=True, out="split")  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(f"T2T fused hq720  {us:8.1f} us   {tok.numel() * 8 / us / 1e6:.2f} TB/s of algorithmic bytes")
