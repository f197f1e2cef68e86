"""Stream-parallel layout over the GPUs of one box: streams are independent, rank r owns a
contiguous block of them, nothing is exchanged on the data path.  The only collective is the optional
gather of decoded frames to one rank."""
import torch
import torch.distributed as dist


def stream_range(rank: int, world: int, streams_per_gpu: int):
    """Global stream ids owned by `rank` (contiguous blocks: stream s lives on GPU s // streams_per_gpu)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return range(rank * streams_per_gpu, (rank + 1) * streams_per_gpu)


def owner_of(stream: int, streams_per_gpu: int) -> int:
    return stream // streams_per_gpu


def gather_frames(local: torch.Tensor, dst: int = 0):
    """Gather every rank's decoded pictures ([streams_per_gpu, picture_bytes] uint8) to `dst`.
    Returns a [world * streams_per_gpu, picture_bytes] tensor in global stream order on dst, None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    recv = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, recv, dst=dst)
    return torch.cat(recv, dim=0) if rank == dst else None
