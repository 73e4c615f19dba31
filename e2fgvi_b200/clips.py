"""Clip-level data parallelism: clips are independent units, so they shard across ranks with no
data-path collective; the ONLY collective is one all-gather of the output frames for the stitch.

One process per GPU (torchrun); ``nccl`` on GPUs, ``gloo`` in the CPU tests.
"""
import os

import torch
import torch.distributed as dist


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


def shard_clips(num_clips, rank, world):
    """Clip ids owned by ``rank``: clip_id % world == rank (round-robin keeps neighbouring windows of one video on
    different GPUs, so a video's windows finish together)."""
    return list(range(rank, num_clips, world))


def padded_share(num_clips, world):
    """Clips per rank after padding to an equal share (all_gather needs equal sizes)."""
    return (num_clips + world - 1) // world


def gather_outputs(local_pred, num_clips, frames_per_clip, rank, world):
    """All-gather per-rank predictions and restore global clip order.

    local_pred: (share * frames_per_clip, 3, H, W) for this rank's clips in ``shard_clips`` order, zero-padded to
    ``padded_share`` clips.  Returns (num_clips * frames_per_clip, 3, H, W) on every rank."""
    if world == 1:
        return local_pred[: num_clips * frames_per_clip]
    share = padded_share(num_clips, world)
    assert local_pred.shape[0] == share * frames_per_clip, (local_pred.shape, share, frames_per_clip)
    gathered = torch.empty((world,) + tuple(local_pred.shape), dtype=local_pred.dtype, device=local_pred.device)
    dist.all_gather_into_tensor(gathered.view(-1, *local_pred.shape[1:]), local_pred.contiguous())
    # gathered[r, j] holds clip j*world + r  ->  order by clip id
    g = gathered.view(world, share, frames_per_clip, *local_pred.shape[1:]).transpose(0, 1)
    return g.reshape(share * world * frames_per_clip, *local_pred.shape[1:])[: num_clips * frames_per_clip]


@torch.no_grad()
def run_clips(model, clips, num_local_frames, rank=0, world=1, device=None, clips_per_call=8):
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
    return gather_outputs(local, num_clips, T, rank, world)
