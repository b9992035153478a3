"""Where do the GPU and the oracle differ?  cfg5-shaped run (1024 streams, trained weights), chunked vs frame-by-frame."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import koala_b200 as kb
from koala_b200 import default_model_path
from oracle import OracleBatch, OracleModel
from test_gpu_baseline_sizes import distinct_pcm, picked_streams

n, frames, chunk = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 512, 64
model = default_model_path()
pcm = distinct_pcm(n, frames, seed=5, pool=32)
eng = kb.BatchKoala(n, model_path=model, precision="bf16")
out = np.concatenate([eng.process(np.ascontiguousarray(pcm[:, t:t + chunk])) for t in range(0, frames, chunk)], axis=1)
eng.delete()
eng = kb.BatchKoala(n, model_path=model, precision="bf16")
one = np.stack([eng.process(np.ascontiguousarray(pcm[:, t])) for t in range(frames)], axis=1)
eng.delete()
print("chunked == frame-by-frame:", bool((out == one).all()), "differing samples", int((out != one).sum()))
picks = picked_streams(n, 4, seed=3)
ref = OracleBatch(OracleModel(model), len(picks), "bf16").process(np.ascontiguousarray(pcm[picks]), threads=os.cpu_count() or 8)
for name, o in (("chunked", out), ("frame-by-frame", one)):
    d = np.abs(o[picks].astype(np.int32) - ref.astype(np.int32))
    print(name, "hist", {int(k): int((d == k).sum()) for k in np.unique(d)})
    per_frame = (d >= 1).sum(axis=(0, 2))
    print("  frames mod 8 share of >=1 LSB:", [int(per_frame[i::8].sum()) for i in range(8)])
    print("  per stream >=1:", (d >= 1).sum(axis=(1, 2)).tolist())
    w = np.argwhere(d >= 2)
    print("  2-LSB at (stream, frame, sample):", w[:12].tolist(), "amplitude there", [int(ref[a, b, c]) for a, b, c in w[:12]])
    print("  first quarter vs last quarter >=1:", int(per_frame[:frames // 4].sum()), int(per_frame[-frames // 4:].sum()))
