"""Micro-benchmark: the small-Cout layers on the plain implicit-GEMM conv (window-packed K where it applies) vs the
kx-in-N kernel, at B=8 clips (112 SPyNet pairs at 64x128, 64 decoder frames at 240x432)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, n, cin, cout, ks, h, w in (("spynet 64->32 k7", 112, 64, 32, 7, 64, 128), ("spynet 32->16 k7", 112, 32, 16, 7, 64, 128),
                                     ("spynet 16->2 k7", 112, 16, 2, 7, 64, 128), ("dec3 64->3 k3", 64, 64, 3, 3, 240, 432)):
    x = torch.randn(n, cin, h, w, device=dev)
    wt = torch.nn.Parameter(torch.randn(cout, cin, ks, ks, device=dev) * 0.05)
    b = torch.nn.Parameter(torch.randn(cout, device=dev))
    xs = ops.split_nhwc(x)
    src_old = ops.pack_rows(x, lead=ks // 2) if cin <= 32 else [xs]
    flops = 2.0 * n * h * w * cout * cin * ks * ks
    t_old = timeit(lambda: ops.conv3x3(src_old, wt, b, negative_slope=0.0, out="f32"))
    t_new = timeit(lambda: ops.conv_kxn(xs, wt, b, negative_slope=0.0, out="f32"))
    print(f"KXN {name:18s} plain {t_old:8.1f} us ({flops / t_old / 1e6:6.1f} TFLOP/s)   kx-in-N {t_new:8.1f} us ({flops / t_new / 1e6:6.1f} TFLOP/s)")

# encoder conv 7: 640 -> 256, groups 8, two sources (256 + 384 channels), 64 frames at 60x108
x0 = torch.randn(64, 256, 60, 108, device=dev)
x1 = torch.randn(64, 384, 60, 108, device=dev)
s0, s1 = ops.split_nhwc(x0), ops.split_nhwc(x1)
wt = torch.nn.Parameter(torch.randn(256, 80, 3, 3, device=dev) * 0.05)
b = torch.nn.Parameter(torch.randn(256, device=dev))
flops = 2.0 * 64 * 60 * 108 * 256 * 80 * 9
t_old = timeit(lambda: ops.conv3x3([s0, s1], wt, b, groups=8, negative_slope=0.2, out="split"))
t_new = timeit(lambda: ops.conv_kxn([s0, s1], wt, b, negative_slope=0.2, out="split", groups=8))
print(f"KXN enc7 640->256 g8      plain {t_old:8.1f} us ({flops / t_old / 1e6:6.1f} TFLOP/s)   kx-in-N {t_new:8.1f} us ({flops / t_new / 1e6:6.1f} TFLOP/s)")
