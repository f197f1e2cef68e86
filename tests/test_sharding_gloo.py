"""The N > 1 layout on CPU: two gloo ranks, each owning its block of streams; the gather puts the
frames back in global stream order.  (The kernels themselves need no collective.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _frames_for(streams, nbytes):
    # a recognisable stand-in for decoded pictures: byte k of stream s is (s * 31 + k) & 0xff
    s = torch.tensor(list(streams), dtype=torch.int64).unsqueeze(1)
    k = torch.arange(nbytes, dtype=torch.int64).unsqueeze(0)
    return ((s * 31 + k) & 0xFF).to(torch.uint8)


def _worker(rank, world, port, per_gpu, nbytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mpeg_b200.sharding import gather_frames, owner_of, stream_range
    mine = stream_range(rank, world, per_gpu)
    assert all(owner_of(s, per_gpu) == rank for s in mine)
    local = _frames_for(mine, nbytes)
    out = gather_frames(local, dst=0)
    if rank == 0:
        want = _frames_for(range(world * per_gpu), nbytes)
        q.put(bool(torch.equal(out, want)))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_sharding_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, 1000, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_stream_range_partitions_everything():
    from mpeg_b200.sharding import owner_of, stream_range
    world, per = 8, 256
    seen = []
    for r in range(world):
        seen += list(stream_range(r, world, per))
    assert seen == list(range(world * per))           # BASELINE config 5: 2048 streams over 8 GPUs
    assert owner_of(2047, per) == 7 and owner_of(256, per) == 1
    with pytest.raises(ValueError):
        stream_range(8, 8, per)
