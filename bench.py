#!/usr/bin/env python
"""bench.py — frames/sec of InpaintGenerator.forward on synthetic 432x240 5+3 clips (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips-per-gpu B]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1, one rank per GPU, NCCL)

A step = one forward over B clips per GPU (default 8 = BASELINE configs[3]'s per-GPU share: 64 clips over 8 GPUs)
followed, for N > 1, by the single all-gather output stitch.  Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, T, L_T = 240, 432, 8, 5
METRIC = "frames/sec InpaintGenerator.forward 432x240x(5+3)"
# other BASELINE.json configs, selectable with --workload (the default "base" is the headline one):
#   name -> (model module, H, W (mirror-padded to multiples of 60 / 108 like test.py:156-165), T, l_t, clips per GPU)
WORKLOADS = {
    "base": ("model.e2fgvi", 240, 432, 8, 5, 8),
    "hq720": ("model.e2fgvi_hq", 720, 1296, 8, 5, 1),       # configs[2]: 720x1280, 5+3
    "hq1080": ("model.e2fgvi_hq", 1080, 1944, 16, 10, 1),   # configs[4]: 1080x1920, 10+6, one clip per GPU
}


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_cores():
    """CPU threads this process may really use: min(affinity mask, cgroup cpu quota) — os.cpu_count() alone
    reports the host's cores inside a quota-limited container and oversubscribes the oracle by 10-100x."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
            out, _ = self.p.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_model(device, module="model.e2fgvi"):
    net = importlib.import_module(module)
    from e2fgvi_b200.synth import synth_state_dict
    model = net.InpaintGenerator().eval()
    sd = synth_state_dict(model, "default", 0)          # the reference's own init family (BASELINE config)
    model.load_state_dict(sd, strict=True)
    return model.to(device), sd


def cpu_oracle_fps(sd, steps=1, warmup=0):
    """The reference's CPU implementation of the path: the oracle port (oracle/restate.py) on all host threads,
    one 5+3 clip per step (the reference itself is single-process, b=1: test.py:108,152-166)."""
    from e2fgvi_b200.synth import synth_frames
    from oracle import restate
    torch.set_num_threads(host_cores())
    x = synth_frames(1, T, H, W, seed=3)
    with torch.no_grad():
        for _ in range(warmup):
            restate.inpaint_generator_forward(sd, x, L_T)
        t0 = time.perf_counter()
        for _ in range(steps):
            restate.inpaint_generator_forward(sd, x, L_T)
        dt = time.perf_counter() - t0
    return steps * T / dt, dt / steps, torch.get_num_threads()


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port; /root/reference does not exist on the GPU box)."""
    if rank != 0:
        return
    net = importlib.import_module("model.e2fgvi")
    from e2fgvi_b200.synth import synth_state_dict
    sd = synth_state_dict(net.InpaintGenerator(), "default", 0)
    fps, s_per_step, cores = cpu_oracle_fps(sd, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "e2fgvi 432x240, 5 local + 3 ref frames, 1 clip per step, CPU", "frames_per_clip": T},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} x one 5+3 clip forward (oracle/restate.py, torch CPU fp32; DCN = "
                                   "explicit restatement, mmcv is not installable offline)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips-per-gpu", type=int, default=None)
    ap.add_argument("--workload", default="base", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="strict", choices=["strict", "tf32"],
                    help="library conv/linear precision: strict = fp32 (TF32 off), tf32 = PyTorch GPU defaults")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    global H, W, T, L_T, METRIC
    module, H, W, T, L_T, default_b = WORKLOADS[args.workload]
    if args.clips_per_gpu is None:
        args.clips_per_gpu = default_b
    if args.workload != "base":
        METRIC = f"frames/sec InpaintGenerator.forward {W}x{H}x({L_T}+{T - L_T}) [{args.workload}]"

    from e2fgvi_b200 import clips as C
    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    args.warmup = max(args.warmup, 3)
    rank, world, local_rank = C.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm"
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from e2fgvi_b200 import build as _build
    from e2fgvi_b200 import ops
    from e2fgvi_b200.synth import synth_frames
    _build.build()
    model, sd = make_model(dev, module)
    model.precision = args.precision
    log(f"rank {rank}/{world}: model ready, precision={args.precision}, host cores={host_cores()}")
    B = args.clips_per_gpu
    num_clips = B * world

    # rotating input sets: 4 x B clips (4 x 80 MB at B=8) > 126 MB L2, so no step re-reads a cached input
    n_sets = 4
    host_sets = [synth_frames(B, T, H, W, seed=100 + rank * n_sets + i).pin_memory() for i in range(n_sets)]
    dev_sets = [h.to(dev) for h in host_sets]
    out_host = torch.empty((B * T, 3, H, W), dtype=torch.float32).pin_memory()

    def step(i):
        pred, _ = model(dev_sets[i % n_sets], L_T)
        return C.gather_outputs(pred, num_clips, T, rank, world) if world > 1 else pred

    def sync():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(args.warmup):
            step(i)
        sync()
        log("warm-up done")
        # ---------------- device-resident timing (value) + live kernel timing for the roofline
        prof = ops.profile_kernels(True)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        sync()
        launches = ops.launch_count() - n0
        clocks = sampler.stop() if sampler else None
        ops.profile_kernels(False)
        ms = e0.elapsed_time(e1)
        log(f"timed region: {ms / args.steps:.2f} ms/step")
        # ---------------- end-to-end through the public API with HOST buffers (pinned H2D in, D2H of the result)
        for i in range(2):
            pred, _ = model(host_sets[i].to(dev, non_blocking=True), L_T)
            out_host.copy_(pred, non_blocking=True)
        sync()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            x = host_sets[i % n_sets].to(dev, non_blocking=True)
            pred, _ = model(x, L_T)
            if world > 1:
                pred = C.gather_outputs(pred, num_clips, T, rank, world)[rank * B * T:(rank + 1) * B * T]
            out_host.copy_(pred, non_blocking=True)
        f1.record()
        sync()
        ms_e2e = f0.elapsed_time(f1)
        log(f"e2e: {ms_e2e / args.steps:.2f} ms/step")
        # single-clip latency (BASELINE configs[1] shape, b=1)
        one = dev_sets[0][:1]
        n_sets_b1 = 1
        for _ in range(3):
            model(one, L_T)
        sync()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            model(one, L_T)
        g1.record()
        sync()
        ms_b1 = g0.elapsed_time(g1) / 5
        # the same single clip replayed from a CUDA graph (launch-bound case)
        ms_b1_graph = None
        try:
            from e2fgvi_b200.graph import GraphedGenerator
            graphed = GraphedGenerator(model, one, L_T)
            for _ in range(3):
                graphed(one)
            sync()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
            for _ in range(10):
                graphed(one)
            h1.record()
            sync()
            ms_b1_graph = h0.elapsed_time(h1) / 10
        except Exception as exc:      # reported, never fatal for the headline numbers
            log(f"CUDA-graph latency run failed: {exc!r}")
        # video-level driver (test.py's sliding-window loop): a synthetic 60-frame 432x240 video
        # from pinned uint8 host buffers to the composited uint8 result back on the host
        video = None
        if args.workload == "base" and rank == 0:
            try:
                import numpy as np
                from e2fgvi_b200.synth import synth_video
                from e2fgvi_b200.video import VideoInpainter
                vf, vm = synth_video(60, H, W, 21)
                vf_t, vm_t = torch.from_numpy(vf).pin_memory(), torch.from_numpy(vm).pin_memory()
                drv = VideoInpainter(model, clips_per_call=4)
                drv(vf_t, vm_t)
                torch.cuda.synchronize()          # rank-0-only section: no collective (sync() holds a barrier)
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record()
                comp = drv(vf_t, vm_t).cpu()
                v1.record()
                torch.cuda.synchronize()
                ms_v = v0.elapsed_time(v1)
                sched = drv.schedule(60)
                video = {"video_frames": 60, "windows": len(sched),
                         "network_frames": sum(len(nb) + len(rf) for _, nb, rf in sched), "ms": ms_v,
                         "video_frames_per_s": 60 / (ms_v * 1e-3),
                         "network_frames_per_s": sum(len(nb) + len(rf) for _, nb, rf in sched) / (ms_v * 1e-3),
                         "untouched_pixels_exact": bool(np.array_equal(comp.numpy()[vm == 0], vf[vm == 0]))}
            except Exception as exc:
                log(f"video driver run failed: {exc!r}")

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    frames = num_clips * T * args.steps
    value = frames / (ms * 1e-3)
    e2e = frames / (ms_e2e * 1e-3)

    if rank == 0:
        hbm, tf_burst, tf_sust, src = peaks()
        # per-kernel live timing (CUDA events around each launch inside the timed region)
        kinds = {"focal_window_attention": ("focal_attn_kernel", "tensor", 1.0),
                 "deform_align_fused": ("dcn_kernel", "tensor", 1.0),
                 # bf16x3 kernels: algorithmic fp32 FLOPs; each costs 3 bf16 MMAs, so <= 1/3 of the bf16 peak
                 "conv3x3_bf16x3": ("conv3x3_kernel", "tensor", 3.0),
                 "linear_bf16x3": ("linear_kernel", "tensor", 3.0),
                 "t2t_fold": ("t2t_fold_kernel", "hbm", 1.0), "t2t_unfold": ("t2t_unfold733_kernel", "hbm", 1.0)}
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes / launch from `ncu --set full`
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        kernels = {}
        for key, (kname, bound, mult) in kinds.items():
            ev = prof.get(key, [])
            if not ev:
                continue
            durs = [a.elapsed_time(b) for a, b, _ in ev]
            work = sum(w for _, _, w in ev)
            tot_ms = sum(durs)
            if bound == "tensor":
                ach = work / (tot_ms * 1e-3) / 1e12
                peak, unit = tf_sust, "TFLOP/s"
            else:
                ach = work / (tot_ms * 1e-3) / 1e9
                peak, unit = hbm, "GB/s"
            kernels[kname] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                              "launches_timed": len(durs), "avg_launch_ms": tot_ms / len(durs),
                              "share_of_step": tot_ms / ms, "traffic": traffic.get(kname)}
            if mult != 1.0:
                kernels[kname]["tensor_work_multiplier"] = mult
                kernels[kname]["tensor_pipe_frac"] = mult * ach / peak
        roof = None
        if kernels:
            dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
            roof = dict(kernels[dom], kernel=dom,
                        peak_source=f"{src} " + ("bf16_tflops_sustained" if kernels[dom]["bound"] == "tensor" else "hbm_gbs"),
                        note="achieved = algorithmic work (SURVEY 8d) / live CUDA-event launch time inside the timed region")
        cpu = None
        if not args.no_cpu_baseline and args.workload == "base":
            log("timing the CPU oracle (1 warm-up + 2 clips) ...")
            fps, s_per, cores = cpu_oracle_fps(sd, steps=2, warmup=1)
            log(f"CPU oracle: {s_per:.2f} s/clip on {cores} threads")
            cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": "2 x one 5+3 clip forward after 1 warm-up (oracle/restate.py, torch CPU fp32)"}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (DCN, attention kernels); bf16 3-term split operands / f32 "
                     "accumulate = fp32-level accuracy (conv, linear kernels)"
                     + ("" if args.precision == "strict" else "; torch library ops tf32"),
            "data": "synthetic",
            "config": {"workload": (f"e2fgvi 432x240, 5 local + 3 ref frames, {B} clips per GPU per step "
                                    "(BASELINE configs[3] per-GPU share)") if args.workload == "base" else
                                   f"{module.split('.')[-1]} {W}x{H} (mirror-padded), {L_T} local + {T - L_T} ref frames, "
                                   f"{B} clip(s) per GPU per step",
                       "global_batch_clips": num_clips, "frames_per_clip": T, "parallelism": f"clip-dp{world}",
                       "l2": f"inputs rotate over {n_sets} sets x {B * T * 3 * H * W * 4 / 1e6:.0f} MB (> 126 MB L2)",
                       "precision": args.precision,
                       "weights": "random-init, reference default family (e2fgvi_b200.synth 'default', seed 0)"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * T * 3 * H * W * 4 * world,
                    "d2h_bytes_per_step": B * T * 3 * H * W * 4 * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_kernels": kernels,
            "cpu_baseline": cpu, "latency_b1_ms": ms_b1, "fps_b1": T / (ms_b1 * 1e-3),
            "latency_b1_cuda_graph_ms": ms_b1_graph, "video_driver": video,
            "fps_b1_cuda_graph": None if not ms_b1_graph else T / (ms_b1_graph * 1e-3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
