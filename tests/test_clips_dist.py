"""CPU: the N>1 path (clip sharding + the single all-gather stitch) with world_size 2 over gloo."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _ToyModel(torch.nn.Module):
    """Stands in for InpaintGenerator: per-clip independent, so sharded == unsharded exactly."""

    def forward(self, x, l_t):
        b, t, c, h, w = x.shape
        return (x * 0.5 + x.mean(dim=(1, 2, 3, 4), keepdim=True)).reshape(b * t, c, h, w), None


def _worker(rank, world, port, num_clips, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from e2fgvi_b200 import clips as C
    r, w, _ = C.init_from_env("gloo")
    g = torch.Generator().manual_seed(0)
    clips = torch.randn(num_clips, 4, 3, 6, 8, generator=g)
    out = C.run_clips(_ToyModel(), clips, 3, rank=r, world=w, clips_per_call=2)
    want, _ = _ToyModel()(clips, 3)
    ok = bool(torch.equal(out, want))
    # compact payloads (what test.py keeps is the uint8 frame, test.py:168-169)
    out8 = C.run_clips(_ToyModel(), clips.clamp(-1, 1), 3, rank=r, world=w, clips_per_call=2, payload="uint8")
    want8 = ((_ToyModel()(clips.clamp(-1, 1), 3)[0] + 1) / 2 * 255).clamp(0, 255).to(torch.uint8)
    ok = ok and out8.dtype == torch.uint8 and bool(torch.equal(out8, want8))
    # asynchronous double-buffered stitch: two stitches in flight, joined in order
    share, T = C.padded_share(num_clips, w), clips.shape[1]
    st = C.make_stitcher(num_clips, T, r, w)              # no CUDA here: every rank must agree on the all-gather class
    ok = ok and type(st) is C.ClipStitcher
    pend = []
    for k in range(3):
        mine = C.shard_clips(num_clips, r, w)
        local, _ = _ToyModel()(clips[mine] + k, 3)
        if local.shape[0] < share * T:
            local = torch.cat([local, local.new_zeros((share * T - local.shape[0],) + tuple(local.shape[1:]))])
        pend.append((k, st.start(local)))
        if len(pend) == 2:
            kk, h = pend.pop(0)
            ok = ok and bool(torch.equal(h.wait(), _ToyModel()(clips + kk, 3)[0]))
    for kk, h in pend:
        ok = ok and bool(torch.equal(h.wait(), _ToyModel()(clips + kk, 3)[0]))
    q.put((rank, ok, tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


def _run(num_clips, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, num_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_two_rank_stitch_even():
    for rank, ok, shape in _run(6, 29611):
        assert ok and shape == (24, 3, 6, 8)


def test_two_rank_stitch_ragged():
    """5 clips over 2 ranks: rank 1 pads its share; the stitch drops the padding."""
    for rank, ok, shape in _run(5, 29612):
        assert ok and shape == (20, 3, 6, 8)


def test_shard_helpers():
    from e2fgvi_b200 import clips as C
    assert C.shard_clips(5, 0, 2) == [0, 1, 2] and C.shard_clips(5, 1, 2) == [3, 4]      # contiguous blocks
    assert C.shard_clips(64, 3, 8) == list(range(24, 32)) and C.shard_clips(3, 3, 4) == []
    assert C.padded_share(5, 2) == 3 and C.padded_share(64, 8) == 8
    x = torch.arange(12.).view(12, 1, 1, 1)
    assert torch.equal(C.gather_outputs(x, 3, 4, 0, 1), x)


# ---- GPU: the peer-memory stitch (csrc/peer.cu) between two PROCESSES sharing cuda:0 — CUDA IPC, DMA pushes and the
# stream-memop flags are the real ones; only the process group (IPC-handle exchange, barriers) runs over gloo because
# NCCL refuses two ranks on one device.  On a multi-GPU box `torchrun bench.py --gpus N` exercises the NVLink route.
def _peer_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK="0")
    from e2fgvi_b200 import clips as C
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda:0")
    num_clips, T = 5, 4                                     # ragged: rank 1 pads its share
    share = C.padded_share(num_clips, world)
    g = torch.Generator().manual_seed(0)
    clips = torch.randn(num_clips, T, 3, 20, 24, generator=g).to(dev)
    want = lambda k: (clips + k).reshape(num_clips * T, 3, 20, 24)             # noqa: E731
    ok = True
    for payload in ("fp32", "uint8"):
        st = C.PeerStitcher(num_clips, T, rank, world, payload)
        pend = []
        for k in range(5):                                  # both landing buffers reused twice
            mine = C.shard_clips(num_clips, rank, world)
            local = (clips[mine] + k).reshape(-1, 3, 20, 24)
            if local.shape[0] < share * T:
                local = torch.cat([local, local.new_zeros((share * T - local.shape[0],) + tuple(local.shape[1:]))])
            if rank == 1 and k == 2:
                torch.cuda._sleep(200_000_000)              # a slow rank: the fast one must not overwrite unread data
            pend.append((k, st.start(local)))
            if len(pend) == 2:
                kk, h = pend.pop(0)
                got = h.wait().clone()
                ok = ok and bool(torch.equal(got, C.encode_payload(want(kk), payload)))
        for kk, h in pend:
            ok = ok and bool(torch.equal(h.wait().clone(), C.encode_payload(want(kk), payload)))
        st.close()
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_peer_memory_stitch_two_processes_one_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, 29631, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
