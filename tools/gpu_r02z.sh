#!/bin/bash
mkdir -p gpurun_out
echo skip-pytest
timeout 300 python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; spec.save_model(m, spec.random_model())
for n in (8192, 4096, 1024):
    eng = kb.BatchKoala(n, model_path=m, precision="bf16")
    for steps in (64, 128, 256):
        h_in = torch.from_numpy((np.random.default_rng(0).standard_normal((steps, n, 256)) * 2000).astype(np.int16)).pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        eng.process(h_in, out=h_out, time_major=True)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); eng.process(h_in, out=h_out, time_major=True); ts.append(time.perf_counter() - t0)
        print(f"{n} streams {steps}-frame time-major host call: {min(ts)*1e3:.3f} ms = {n*steps/min(ts)/1e6:.2f} M frames/s (median {n*steps/sorted(ts)[2]/1e6:.2f})")
    eng.delete()
PY
