"""world_size-2 gloo test of the N>1 path: `ShardedKoala` partitions the streams by contiguous block, there is no
data-path collective, and the gathered result equals the single-process result bit for bit.
The per-rank engine is injected (an oracle-backed stand-in for `BatchKoala`): what is under test is the host-side
partition / gather / statistics logic, which is the same code the GPU ranks run."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, synth_pcm


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _OracleEngine:
    """`BatchKoala`'s surface (process / reset / delete + constants) over the CPU oracle."""
    frame_length, delay_sample, sample_rate = 256, 256, 16000

    def __init__(self, model_path, n):
        from oracle import OracleBatch, OracleModel
        self._ob, self.num_streams = OracleBatch(OracleModel(model_path), n, "bf16"), n

    def process(self, pcm, out=None, time_major=False):
        return self._ob.process(np.ascontiguousarray(pcm))

    def reset(self, ids=None):
        assert ids is None
        self._ob.reset()

    def delete(self):
        pass


def _worker(rank, world, port, model_path, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from koala_b200 import ShardedKoala, owner_of, shard_streams
    total, frames = 11, 3
    pcm = synth_pcm(total, frames, seed=21)
    sk = ShardedKoala(total, engine_factory=lambda n, dev: _OracleEngine(model_path, n))
    assert (sk.first_stream, sk.num_streams) == shard_streams(total, rank, world)
    assert all(owner_of(s, total, world) == rank for s in range(sk.first_stream, sk.first_stream + sk.num_streams))
    out = sk.process(sk.local_slice(pcm))
    whole = sk.gather(out)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), whole)
    else:
        assert whole is None
    frames_total, seconds = sk.job_stats(sk.num_streams * frames, 0.25 * (rank + 1))
    assert frames_total == total * frames and abs(seconds - 0.25 * world) < 1e-9    # SUM of units, MAX of time
    sk.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process(random_model_path, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), random_model_path, str(tmp_path)), nprocs=world, join=True)
    from oracle import OracleBatch, OracleModel
    pcm = synth_pcm(11, 3, seed=21)
    whole = OracleBatch(OracleModel(random_model_path), 11, "bf16").process(pcm)
    assert (np.load(tmp_path / "gathered.npy") == whole).all()


def test_unlaunched_sharded_koala_is_one_rank(random_model_path):
    from koala_b200 import ShardedKoala
    sk = ShardedKoala(5, engine_factory=lambda n, dev: _OracleEngine(random_model_path, n))
    assert (sk.rank, sk.world_size, sk.first_stream, sk.num_streams) == (0, 1, 0, 5)
    pcm = synth_pcm(5, 2, seed=3)
    assert (sk.gather(sk.process(pcm)) == _OracleEngine(random_model_path, 5).process(pcm)).all()
