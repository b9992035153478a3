#!/bin/bash
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python - <<'PY'
import numpy as np, torch, time, sys
sys.path.insert(0, '.')
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/r.kpv"; spec.save_model(m, spec.random_model())
for n, prec in ((128, "bf16"), (256, "fp32"), (4096, "bf16"), (8192, "bf16")):
    eng = kb.BatchKoala(n, model_path=m, precision=prec)
    for steps in (16, 64, 512) if n <= 256 else (32, 128):
        h_in = torch.from_numpy((np.random.default_rng(0).standard_normal((steps, n, 256)) * 2000).astype(np.int16)).pin_memory()
        h_out = torch.empty_like(h_in).pin_memory()
        eng.process(h_in, out=h_out, time_major=True)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); eng.process(h_in, out=h_out, time_major=True); ts.append(time.perf_counter() - t0)
        print(f"{n} streams {prec} {steps}-frame time-major host call: {min(ts)*1e3:.3f} ms = {n*steps/min(ts)/1e6:.2f} M frames/s (median {sorted(ts)[2]*1e3:.3f} ms)")
    eng.delete()
PY
