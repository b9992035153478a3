#!/bin/bash
# Round 2: fixed-point variant (tcgen05 kind::i8) -- GPU tests and a first timing.
mkdir -p gpurun_out
echo "== fixed-point tests";  timeout 900 python -m pytest tests/test_gpu_fixed_point.py -m gpu -x -q 2>&1 | tail -25
echo "== timing"
timeout 300 python - <<'PY'
import numpy as np, torch, time, os, sys
sys.path.insert(0, '.')
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; spec.save_model(m, spec.random_model())
for n in (256, 4096, 8192):
    eng = kb.BatchKoala(n, model_path=m, precision="int8")
    pcm = torch.from_numpy((np.random.default_rng(0).standard_normal((n, 16, 256)) * 2000).astype(np.int16)).cuda()
    out = torch.empty_like(pcm)
    for _ in range(3): eng.process(pcm, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(10): eng.process(pcm, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 160 * 1e3
    eng.profile(True); eng.process(pcm, out=out); prof = eng.profile_read(); eng.profile(False)
    print(f"int8 {n} streams: {us:.1f} us/step = {n / us:.2f} M frames/s | " + " ".join(f"{k} {v[0] / 16 * 1e3:.1f}" for k, v in prof.items() if v[1]))
    eng.delete()
PY
