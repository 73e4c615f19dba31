"""Dump the clock64 event trace of CTA (0,0,0) of the focal attention kernel (perf experiments)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(0)
qkv = torch.randn(B, 8, 20, 36, 1536, generator=g).half().to(dev)
pooled = torch.randn(B, 8, 4, 4, 1536, generator=g).half().to(dev)
buf = torch.zeros(3 * 64, dtype=torch.int64, device=dev)
run = lambda: ops.focal_window_attention(qkv, pooled, 4, (5, 9), (2, 4), (5, 9), 128 ** -0.5, out_dtype=torch.float32)  # noqa
for _ in range(3):
    run()
torch.cuda.synchronize()
os.environ["E2F_ATTN_TRACE"] = hex(buf.data_ptr())
run()
torch.cuda.synchronize()
del os.environ["E2F_ATTN_TRACE"]
t = buf.cpu().view(3, 64)
t0 = int(t[t > 1].min())
names = ["softmax: start Sready Sregs math Pstored arrived", "loader: idx free issued Kdone Vdone",
         "mma: waitP Pready PVissued Sissued"]
fine = bool(int(os.environ.get("E2F_ATTN_DEBUG", "0")) & 8)
per = [7 if fine else 5, 5, 4]
if fine:
    names[0] = "softmax: start Sready maxdone decided math Pstored arrived"
    print("rescale fired in tiles:", [k for k in range(10) if int(t[2][63 - k]) == 1])
for role in range(3):
    print(names[role])
    row = [int(v) - t0 if v > 0 else -1 for v in t[role]]
    for i in range(0, 64, per[role]):
        chunk = row[i:i + per[role]]
        if all(c < 0 for c in chunk):
            break
        print(f"  tile {i // per[role]:2d}: " + " ".join(f"{c:7d}" for c in chunk))
