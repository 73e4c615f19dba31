"""Print max-abs error of the GPU forward against the reference goldens under different library-precision flags."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from e2fgvi_b200.synth import synth_frames, synth_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
dev = torch.device("cuda:0")
for name in ["e2e_hq_tiny_stress", "e2e_hq_small_stress", "e2e_base_stress", "e2e_base_default"]:
    g = torch.load(os.path.join(GOLDEN, name + ".pt"))
    c = g["case"]
    net = importlib.import_module("model." + ("e2fgvi_hq" if c["hq"] else "e2fgvi"))
    m = net.InpaintGenerator().eval()
    m.load_state_dict(synth_state_dict(m, c["family"], c["weight_seed"]))
    m.to(dev)
    x = synth_frames(1, c["T"], c["H"], c["W"], seed=c["frame_seed"]).to(dev)
    for conv_tf32 in (True, False):
        for mm_tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = conv_tf32
            torch.backends.cuda.matmul.allow_tf32 = mm_tf32
            with torch.no_grad():
                pred, (ff, fb) = m(x, c["l_t"])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pred, (ff, fb) = m(x, c["l_t"])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            s = g["subsample"]
            err = (pred[:, :, ::s, ::s].cpu() - g["pred"]).abs().max().item()
            ferr = (ff.cpu() - g["flows_forward"]).abs().max().item()
            print(f"{name:22s} conv_tf32={conv_tf32!s:5s} matmul_tf32={mm_tf32!s:5s} max|err|={err:.3e} "
                  f"flow_err={ferr:.2e} (flow max {g['flows_forward'].abs().max():.1f}) {dt*1e3:.1f} ms", flush=True)
