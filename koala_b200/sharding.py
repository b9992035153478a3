"""Stream partition across the GPUs of one box.

Streams are independent units with O(1) state, so the path shards with no data-path collective: rank r owns a contiguous
block of streams for their whole lifetime (state, PCM in/out stay on the owning GPU).  torch.distributed is used only
off the per-frame path -- barriers and the reduction of counters / timings -- over NCCL on GPUs and gloo in CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_streams(total_streams: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition: returns (first_stream, num_streams) of `rank`; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size) or total_streams < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(total_streams, world_size)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def owner_of(stream: int, total_streams: int, world_size: int) -> int:
    """Inverse of shard_streams: the rank that owns `stream`."""
    base, extra = divmod(total_streams, world_size)
    boundary = extra * (base + 1)
    if stream < boundary:
        return stream // (base + 1)
    return extra + (stream - boundary) // base


def reduce_job_stats(frames_local: int, seconds_local: float, device=None):
    """Whole-job (total frames, max seconds) over all ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return frames_local, seconds_local
    t = torch.tensor([float(frames_local)], dtype=torch.float64, device=device)
    s = torch.tensor([float(seconds_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return int(round(t.item())), s.item()


__all__ = ['shard_streams', 'owner_of', 'reduce_job_stats']
