"""world_size-2 gloo test of the N>1 path: streams shard by contiguous block, no data-path collective, and the gathered
result equals the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, synth_pcm


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, model_path, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from koala_b200 import reduce_job_stats, shard_streams
    from oracle import OracleBatch, OracleModel
    total, frames = 11, 3
    pcm = synth_pcm(total, frames, seed=21)
    start, count = shard_streams(total, rank, world)
    out = OracleBatch(OracleModel(model_path), count, "bf16").process(pcm[start:start + count])
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), out)
    frames_total, seconds = reduce_job_stats(count * frames, 0.25 * (rank + 1))
    assert frames_total == total * frames and abs(seconds - 0.25 * world) < 1e-9    # SUM of units, MAX of time
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one_process(random_model_path, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), random_model_path, str(tmp_path)), nprocs=world, join=True)
    from oracle import OracleBatch, OracleModel
    pcm = synth_pcm(11, 3, seed=21)
    whole = OracleBatch(OracleModel(random_model_path), 11, "bf16").process(pcm)
    gathered = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert (gathered == whole).all()
