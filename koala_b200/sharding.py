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


class ShardedKoala(object):
    """The multi-GPU front door: one process per GPU (launch with torchrun / torch.distributed.run), rank r owns the
    contiguous stream block `shard_streams(total_streams, r, world)` on `cuda:LOCAL_RANK` for the streams' whole lifetime.

    There is no data-path collective: `process` only touches the local engine.  torch.distributed (NCCL on GPUs, gloo in the
    CPU tests) is used for what happens off the per-frame path: `gather` (collect enhanced PCM on rank 0, e.g. to write
    files), `job_stats` (SUM of frames, MAX of seconds) and `barrier`.  Works un-launched too (world size 1).

    `engine_factory(num_streams, device_index)` builds the per-rank engine; the default is `BatchKoala` on the local GPU.
    """

    def __init__(self, total_streams: int, model_path=None, precision: str = 'bf16', library_path=None, engine_factory=None):
        import os
        import torch.distributed as dist
        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = self._dist.get_rank() if self._dist else 0
        self.world_size = self._dist.get_world_size() if self._dist else 1
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.total_streams = int(total_streams)
        self.first_stream, self.num_streams = shard_streams(self.total_streams, self.rank, self.world_size)
        if self.num_streams < 1:
            raise ValueError('rank %d of %d owns no stream of %d' % (self.rank, self.world_size, self.total_streams))
        if engine_factory is None:
            from ._batch import BatchKoala

            def engine_factory(n, dev):
                return BatchKoala(n, model_path=model_path, device='gpu:%d' % dev, precision=precision, library_path=library_path)
        self.engine = engine_factory(self.num_streams, self.local_rank)
        self.frame_length = self.engine.frame_length
        self.delay_sample = self.engine.delay_sample
        self.sample_rate = self.engine.sample_rate

    def local_slice(self, global_array):
        """This rank's rows of a [total_streams, ...] array (a view)."""
        return global_array[self.first_stream:self.first_stream + self.num_streams]

    def process(self, pcm_local, out=None, time_major: bool = False):
        """One or more steps for the streams this rank owns; same arguments as `BatchKoala.process`."""
        if out is None and not time_major:
            return self.engine.process(pcm_local)
        return self.engine.process(pcm_local, out=out, time_major=time_major)

    def reset(self, global_stream_ids=None):
        """`pv_koala_reset` for every stream, or for the listed GLOBAL stream ids that this rank owns (others are ignored)."""
        if global_stream_ids is None:
            return self.engine.reset()
        mine = [s - self.first_stream for s in global_stream_ids if owner_of(s, self.total_streams, self.world_size) == self.rank]
        if mine:
            self.engine.reset(mine)

    def gather(self, out_local):
        """Collect every rank's [num_streams_r, ...] numpy result on rank 0 as one [total_streams, ...] array (None elsewhere).
        Off the hot path: object gather through the process group."""
        import numpy as np
        if not self._dist:
            return np.asarray(out_local)
        parts = [None] * self.world_size if self.rank == 0 else None
        self._dist.gather_object(np.asarray(out_local), parts, dst=0)
        return np.concatenate(parts, axis=0) if self.rank == 0 else None

    def job_stats(self, frames_local: int, seconds_local: float, device=None):
        return reduce_job_stats(frames_local, seconds_local, device)

    def barrier(self):
        if self._dist:
            self._dist.barrier()

    def delete(self):
        self.engine.delete()


__all__ = ['shard_streams', 'owner_of', 'reduce_job_stats', 'ShardedKoala']
