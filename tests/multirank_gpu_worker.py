"""Per-rank body of tests/test_gpu_multirank.py; launched with `python -m torch.distributed.run --nproc-per-node N`.

Every rank runs the product's multi-GPU entry (`koala_b200.ShardedKoala`, NCCL process group) on ITS shard of one global
PCM tensor, checks four of its own streams against the CPU oracle and reports a checksum; rank 0 collects the verdicts, the
gathered output, and checks that shard boundaries and rank order are right."""
import hashlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import koala_b200 as kb
    from koala_b200 import spec
    from oracle import OracleBatch, OracleModel
    from test_gpu_baseline_sizes import distinct_pcm

    total, frames, out_path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    model = os.path.join(ROOT, "gpurun_out", f"multirank_{os.getpid()}.kpv")
    os.makedirs(os.path.dirname(model), exist_ok=True)
    spec.save_model(model, spec.random_model())
    pcm = distinct_pcm(total, frames, seed=99)                 # the same global tensor on every rank (seeded)
    sk = kb.ShardedKoala(total, model_path=model, precision="bf16")
    assert sk.engine.device_index == local_rank
    mine = np.ascontiguousarray(sk.local_slice(pcm))
    out = sk.process(mine)
    picks = sorted(set([0, sk.num_streams // 3, (2 * sk.num_streams) // 3, sk.num_streams - 1]))
    ref = OracleBatch(OracleModel(model), len(picks), "bf16").process(np.ascontiguousarray(mine[picks]), threads=4)
    worst = int(np.abs(out[picks].astype(np.int32) - ref.astype(np.int32)).max())
    verdict = {"rank": rank, "device": torch.cuda.get_device_name(local_rank), "first_stream": sk.first_stream,
               "num_streams": sk.num_streams, "oracle_checked_streams": [sk.first_stream + p for p in picks],
               "max_lsb_vs_oracle": worst, "sha1": hashlib.sha1(out.tobytes()).hexdigest()}
    verdicts = [None] * world if rank == 0 else None
    dist.gather_object(verdict, verdicts, dst=0)
    whole = sk.gather(out)
    frames_total, _ = sk.job_stats(sk.num_streams * frames, 1.0, device=torch.device("cuda", local_rank))
    if rank == 0:
        ok = all(v["max_lsb_vs_oracle"] <= 1 for v in verdicts) and frames_total == total * frames
        ok = ok and whole.shape == pcm.shape and [v["first_stream"] for v in verdicts] == sorted(v["first_stream"] for v in verdicts)
        # every rank's block sits where shard_streams says, and equals what that rank hashed
        for v in verdicts:
            blk = whole[v["first_stream"]:v["first_stream"] + v["num_streams"]]
            ok = ok and hashlib.sha1(np.ascontiguousarray(blk).tobytes()).hexdigest() == v["sha1"]
        # streams that carry identical input on different ranks (none by construction) would be a data bug: all distinct
        ok = ok and len({v["sha1"] for v in verdicts}) == world
        with open(out_path, "w") as f:
            json.dump({"ok": bool(ok), "world": world, "total_streams": total, "frames": frames, "ranks": verdicts}, f, indent=1)
    sk.barrier()
    sk.delete()
    try:
        os.remove(model)
    except OSError:
        pass
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
