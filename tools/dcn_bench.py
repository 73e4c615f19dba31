"""Micro-benchmark of the fused deformable alignment kernel (feat_prop.py:41-58) at B clips x 60x108 (one propagation step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
for n in (8, 1):
    h, w = 60, 108
    a = torch.randn(n, 128, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    b = torch.randn(n, 128, h, w, device=dev).contiguous(memory_format=torch.channels_last)
    xg = ops.dcn_pack_input(a, b)
    head = (torch.randn(n, 432, h, w, device=dev) * 0.5).contiguous(memory_format=torch.channels_last)
    f1 = torch.randn(n, 2, h, w, device=dev) * 2
    f2 = torch.randn(n, 2, h, w, device=dev) * 2
    wp = ops.pack_dcn_weight(torch.randn(128, 256, 3, 3, device=dev) / 48, 16)
    bias = torch.randn(128, device=dev)
    run = lambda: ops.deform_align_fused(xg, head, f1, f2, wp, bias, 16, 10.0, out_split=True)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"DCN fused n={n}: {us:8.1f} us  {2.0 * 128 * 2304 * n * h * w / us / 1e6:7.1f} TFLOP/s")
