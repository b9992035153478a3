#!/bin/bash
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import numpy as np, koala_b200 as kb
k = kb.create(access_key=kb.ANY_ACCESS_KEY, device="gpu:0")
frame = [int(v) for v in (np.random.default_rng(0).standard_normal(256) * 2000).astype(np.int16)]
for _ in range(50): k.process(frame)
t0 = time.perf_counter()
for _ in range(1000): k.process(frame)
print("Koala.process (python list in/out): %.1f us per call" % ((time.perf_counter() - t0) / 1000 * 1e6))
k.delete()
b = kb.BatchKoala(1)
x = np.zeros((1, 1, 256), np.int16); o = np.empty_like(x)
for _ in range(50): b.process(x, out=o)
t0 = time.perf_counter()
for _ in range(1000): b.process(x, out=o)
print("BatchKoala(1).process (numpy in/out): %.1f us per call" % ((time.perf_counter() - t0) / 1000 * 1e6))
b.delete()
PY
