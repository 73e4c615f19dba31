"""Clip-level data parallelism: clips are independent units, so they shard across ranks with no
data-path collective; the ONLY collective is one all-gather of the output frames for the stitch.

One process per GPU (torchrun); ``nccl`` on GPUs, ``gloo`` in the CPU tests.

Round 2 (VERDICT r01 "Multi-GPU"):
* clips shard in contiguous BLOCKS (rank r owns clips [r*share, (r+1)*share)), so the all-gather lands in global clip
  order and the 637 MB transpose-reshape copy of the round-robin layout is gone;
* the stitch is asynchronous and double-buffered (``ClipStitcher``): the all-gather of step i runs on the
  communication stream while the forward of step i+1 computes;
* the payload can be fp32 (exact, default), fp16, or the uint8 frames test.py actually keeps
  (``((x + 1) / 2 * 255)`` truncated, test.py:168-169) — 2x / 4x fewer bytes over NVLink.
"""
import os

import torch
import torch.distributed as dist

PAYLOADS = ("fp32", "fp16", "uint8")


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local_rank


def padded_share(num_clips, world):
    """Clips per rank after padding to an equal share (all_gather needs equal sizes)."""
    return (num_clips + world - 1) // world


def shard_clips(num_clips, rank, world):
    """Clip ids owned by ``rank``: the contiguous block [rank*share, (rank+1)*share) clipped to num_clips.  Block
    (not round-robin) ownership makes the all-gather result already ordered by clip id."""
    share = padded_share(num_clips, world)
    return list(range(min(rank * share, num_clips), min((rank + 1) * share, num_clips)))


def encode_payload(pred, payload="fp32"):
    """What a rank sends into the stitch: fp32 predictions, their fp16 rounding, or test.py's uint8 frames
    (``(x + 1) / 2 * 255`` truncated like ``astype(np.uint8)``, test.py:168-169)."""
    if payload == "fp32":
        return pred.contiguous()
    if payload == "fp16":
        return pred.to(torch.float16).contiguous()
    if payload == "uint8":
        return ((pred + 1) / 2 * 255).clamp_(0, 255).to(torch.uint8).contiguous()
    raise ValueError(f"payload must be one of {PAYLOADS}")


class _Pending:
    """Handle of one in-flight stitch: ``wait()`` orders the CURRENT stream after the collective and returns the
    gathered (num_clips*T, 3, H, W) tensor (a view of one of the stitcher's two landing buffers: consume it before
    the stitch after next is started)."""

    def __init__(self, work, out, n_valid, keep):
        self.work, self.out, self.n_valid, self.keep = work, out, n_valid, keep

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        self.keep = None
        return self.out[: self.n_valid]


class ClipStitcher:
    """Double-buffered asynchronous output stitch.  ``start(local_pred)`` launches the all-gather (NCCL: on the
    process group's communication stream, ordered after everything already enqueued on the current stream) and
    returns immediately, so the next forward overlaps it; ``_Pending.wait()`` joins."""

    def __init__(self, num_clips, frames_per_clip, rank, world, payload="fp32"):
        if payload not in PAYLOADS:
            raise ValueError(f"payload must be one of {PAYLOADS}")
        self.num_clips, self.T, self.rank, self.world, self.payload = num_clips, frames_per_clip, rank, world, payload
        self.share = padded_share(num_clips, world)
        self._bufs = [None, None]
        self._n = 0

    def start(self, local_pred):
        """local_pred: (share * T, 3, H, W) — this rank's clips in ``shard_clips`` order, zero-padded to the share."""
        n_valid = self.num_clips * self.T
        send = encode_payload(local_pred, self.payload)
        if self.world == 1:
            return _Pending(None, send, n_valid, None)
        if send.shape[0] != self.share * self.T:
            raise ValueError(f"local_pred has {send.shape[0]} frames, expected share*T = {self.share * self.T}")
        k = self._n & 1
        self._n += 1
        shape = (self.world * send.shape[0],) + tuple(send.shape[1:])
        buf = self._bufs[k]
        if buf is None or buf.shape != shape or buf.dtype != send.dtype or buf.device != send.device:
            buf = self._bufs[k] = torch.empty(shape, dtype=send.dtype, device=send.device)
        work = dist.all_gather_into_tensor(buf, send, async_op=True)
        return _Pending(work, buf, n_valid, send)


def gather_outputs(local_pred, num_clips, frames_per_clip, rank, world, payload="fp32"):
    """Blocking stitch: all-gather per-rank predictions; block sharding means the result is already in clip order.

    local_pred: (share * frames_per_clip, 3, H, W) for this rank's clips, zero-padded to ``padded_share`` clips.
    Returns (num_clips * frames_per_clip, 3, H, W) on every rank."""
    return ClipStitcher(num_clips, frames_per_clip, rank, world, payload).start(local_pred).wait()


@torch.no_grad()
def run_clips(model, clips, num_local_frames, rank=0, world=1, device=None, clips_per_call=8, payload="fp32"):
    """Run ``model`` over this rank's share of ``clips`` (num_clips, T, 3, H, W; same on every rank or only the
    local share is read) and return the stitched (num_clips*T, 3, H, W) predictions on every rank."""
    num_clips, T = clips.shape[0], clips.shape[1]
    mine = shard_clips(num_clips, rank, world)
    share = padded_share(num_clips, world)
    outs = []
    for i in range(0, len(mine), clips_per_call):
        batch = clips[mine[i:i + clips_per_call]]
        if device is not None:
            batch = batch.to(device, non_blocking=True)
        pred, _ = model(batch, num_local_frames)
        outs.append(pred)
    local = torch.cat(outs) if outs else clips.new_zeros((0, 3) + tuple(clips.shape[-2:]))
    if device is not None and local.device != torch.device(device):
        local = local.to(device)
    if local.shape[0] < share * T:
        pad = local.new_zeros((share * T - local.shape[0],) + tuple(local.shape[1:]))
        local = torch.cat([local, pad])
    return gather_outputs(local, num_clips, T, rank, world, payload)
