#!/bin/bash
# First-contact GPU script: each stage under its own timeout, logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== fp32 smoke" ; KOALA_SMOKE_ONLY=fp32 timeout 300 python - <<'PY' 2>&1 | tail -20
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import koala_b200 as kb
from koala_b200 import spec
from oracle import OracleBatch, OracleModel
m = 'gpurun_out/r.kpv'; spec.save_model(m, spec.random_model())
rng = np.random.default_rng(0)
pcm = (rng.standard_normal((5, 8, 256)) * 3000).astype(np.int16)
for prec in ('fp32',):
    eng = kb.BatchKoala(5, model_path=m, precision=prec)
    out = eng.process(pcm)
    ref = OracleBatch(OracleModel(m), 5, prec).process(pcm, threads=2)
    d = np.abs(out.astype(int) - ref.astype(int))
    print(prec, 'max LSB diff', d.max(), 'per-frame max', d.max(axis=(0, 2)))
PY
echo "== bf16 smoke" ; set -o pipefail; timeout 90 python - <<'PY' 2>&1 | tail -30
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import koala_b200 as kb
from koala_b200 import spec
from oracle import OracleBatch, OracleModel
m = 'gpurun_out/r.kpv'
rng = np.random.default_rng(0)
pcm = (rng.standard_normal((5, 8, 256)) * 3000).astype(np.int16)
eng = kb.BatchKoala(5, model_path=m, precision='bf16')
out = eng.process(pcm)
ob = OracleBatch(OracleModel(m), 5, 'bf16'); ref = ob.process(pcm, threads=2)
d = np.abs(out.astype(int) - ref.astype(int))
print('bf16 max LSB diff', d.max(), 'per-frame max', d.max(axis=(0, 2)))
assert d.max() <= 2, 'bf16 parity broken'
mask = eng.debug_read('mask', (5, 256), np.float32); rm = np.stack([ob.stream(s).last_mask for s in range(5)])
print('mask max abs diff', np.abs(mask - rm).max(), 'mask range', mask.min(), mask.max())
for l in range(2):
    h = eng.debug_read(f'h{l}', (5, 512), np.float32); rh = np.stack([ob.stream(s).h[l] for s in range(5)])
    print('h', l, 'max abs diff', np.abs(h - rh).max(), 'first bad cols', np.argwhere(np.abs(h - rh) > 1e-3)[:6].tolist())
PY
rc=$?; if [ $rc -ne 0 ]; then echo "bf16 smoke failed or hung (rc=$rc): stopping here"; exit 1; fi
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench bf16" ; timeout 600 python bench.py --steps 100 --warmup 10 --cpu-seconds 5 2>&1 | tail -5
echo "== bench fp32" ; timeout 600 python bench.py --workload cfg2_256_fp32 --steps 100 --warmup 10 --no-cpu-baseline 2>&1 | tail -3
