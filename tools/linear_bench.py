"""Micro-benchmark of the bf16x3 linear kernel on the GEMM shapes of the path at B=8 clips (M = 46080 tokens)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 46080
SHAPES = [("attn.qkv", M, 512, 1536, torch.float16, False), ("attn.proj", M, 512, 512, torch.float32, True),
          ("mlp.conv1", M, 512, 1960, torch.float32, False), ("mlp.conv2", M, 1960, 512, torch.float32, True),
          ("ss.embedding", M, 6272, 512, torch.float32, False), ("sc.embedding", M, 512, 6272, torch.float32, False),
          ("fusion 1x1", 259200, 256, 128, torch.float32, True)]
tot = 0.0
for name, m, k, n, odt, res in SHAPES:
    a = torch.randn(m, k, device=dev)
    hi, lo = ops.split_bf16(a)
    x = ops.SplitMat(hi, lo)
    w = torch.nn.Parameter(torch.randn(n, k, device=dev) / k ** 0.5)
    b = torch.randn(n, device=dev)
    r = torch.randn(m, n, device=dev) if res else None
    run = lambda: ops.linear(x, w, b, residual=r, out_dtype=odt)  # noqa: E731
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    tot += us
    print(f"LINEAR {name:14s} M={m} K={k} N={n}: {us:8.1f} us  {2.0 * m * n * k / us / 1e6:7.1f} TFLOP/s")
    del a, hi, lo, x, r
print(f"LINEAR total {tot / 1e3:.2f} ms")

# SoftSplit / SoftComp as gather convs (round 2), same FLOPs as ss.embedding / sc.embedding above + the unfold / fold passes
x = torch.randn(64, 128, 60, 108, device=dev).contiguous(memory_format=torch.channels_last)
xs = ops.split_nhwc(x)
ss_w = torch.nn.Parameter(torch.randn(512, 6272, device=dev) / 6272 ** 0.5)
ss_b = torch.randn(512, device=dev)
tok = torch.randn(64, 20, 36, 512, device=dev)
hi, lo = ops.split_bf16(tok)
toks = ops.SplitMat(hi, lo)
sc_w = torch.nn.Parameter(torch.randn(6272, 512, device=dev) / 512 ** 0.5)
sc_b = torch.nn.Parameter(torch.randn(6272, device=dev))
sc_map = torch.nn.Parameter(torch.randn(128, 60, 108, device=dev))
for name, run in (("soft_split conv", lambda: ops.soft_split(xs, ss_w, ss_b, 7, 3, 3)),
                  ("soft_comp conv", lambda: ops.soft_comp(toks, sc_w, sc_b, (60, 108), 7, 3, 3, bias_map_extra=sc_map, residual=x))):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"LINEAR {name:14s} 64 frames 128<->512 k7 s3: {us:8.1f} us  {2.0 * 46080 * 6272 * 512 / us / 1e6:8.1f} TFLOP/s")
