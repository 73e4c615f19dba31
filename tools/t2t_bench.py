"""Micro-benchmark: FusionFeedForward middle (fold / fold(ones) -> unfold -> GELU) as two kernels vs the fused one,
B=8 clips (64 frames), 40 channels, 60x108 -> 720 tokens x 1960."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bt = int(sys.argv[1]) if len(sys.argv) > 1 else 64
tok = torch.randn(bt, 720, 1960, device=dev)
geo = ((7, 7), (3, 3), (3, 3))


def pair():
    img = ops.t2t_fold(tok, (60, 108), *geo, normalize=True)
    return ops.t2t_unfold(img, *geo, gelu=True, out="split")


def fused():
    return ops.t2t_fold_unfold(tok, (60, 108), *geo, gelu=True, out="split")


only = sys.argv[2] if len(sys.argv) > 2 else ""
for name, fn in (("fold+unfold", pair), ("fused", fused)):
    if only and only != name:
        continue
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
    print(f"T2T {name:12s} {us:8.1f} us   {tok.numel() * 8 / us / 1e6:.2f} TB/s of the fused kernel's algorithmic bytes")

if only:
    sys.exit(0)
# SoftComp fold (tokens 64 x 720 x 6272 -> 128 x 60 x 108, + bias) and SoftSplit unfold (the reverse, split output)
tok2 = torch.randn(bt, 720, 6272, device=dev)
img2 = torch.randn(bt, 128, 60, 108, device=dev)
bias = torch.randn(128, 60, 108, device=dev)
for name, fn, nbytes in (("sc fold", lambda: ops.t2t_fold(tok2, (60, 108), *geo, bias=bias), tok2.numel() * 4 + img2.numel() * 4),
                         ("ss unfold", lambda: ops.t2t_unfold(img2, *geo, out="split"), tok2.numel() * 4 + img2.numel() * 4)):
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
    print(f"T2T {name:12s} {us:8.1f} us   {nbytes / us / 1e6:.2f} TB/s")
