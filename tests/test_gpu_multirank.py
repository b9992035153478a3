"""N > 1 on real GPUs: one process per GPU through `ShardedKoala`, every rank's OWN output oracle-checked (run on a box
with >= 2 B200s: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`; skipped on a 1-GPU box)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_every_rank_output_is_oracle_checked(library_path, tmp_path):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    report = os.path.join(ROOT, "gpurun_out", f"multirank_{world}gpu.json")
    os.makedirs(os.path.dirname(report), exist_ok=True)
    total = 600 * world + 37                      # uneven split: some ranks own one stream more; three tiles per rank
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multirank_gpu_worker.py"), str(total), "12", report]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-4000:]
    rep = json.load(open(report))
    assert rep["ok"] and rep["world"] == world and len(rep["ranks"]) == world, rep
    assert sum(r["num_streams"] for r in rep["ranks"]) == total
