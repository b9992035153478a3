#!/bin/bash
# Round 2: whole GPU suite + smoke + default bench line (all BASELINE workloads + the fixed-point variant under `others`) + compute-sanitizer logs.
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench";   timeout 900 python bench.py --cpu-seconds 8 2> gpurun_out/bench_r02n.err | tail -1 > gpurun_out/bench_r02n_default.json; tail -3 gpurun_out/bench_r02n.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02n_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'))
print(d.get('kernel_ms_per_step'), d['clocks']); print(d.get('host_link'))
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','steps','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
echo "== compute-sanitizer memcheck (smoke: all three precisions)"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02n_sanitizer_memcheck.log python __graft_entry__.py smoke > gpurun_out/r02n_sanitizer_memcheck.out 2>&1; tail -3 gpurun_out/r02n_sanitizer_memcheck.log
echo "== compute-sanitizer racecheck (fp32 CUDA-core + fixed-point + STFT kernels: small engine run)"
cat > gpurun_out/san_small.py <<'PY'
import sys, os, numpy as np
sys.path.insert(0, '.')
import koala_b200 as kb
from koala_b200 import spec
m = "gpurun_out/s.kpv"; spec.save_model(m, spec.random_model(hidden=256, layers=1), hidden=256, layers=1)
pcm = (np.random.default_rng(0).standard_normal((3, 2, 256)) * 3000).astype(np.int16)
for prec, env in (("fp32", "1"), ("int8", "0"), ("bf16", "0")):
    os.environ["KOALA_FP32_CUDA_CORES"] = env
    eng = kb.BatchKoala(3, model_path=m, precision=prec); out = eng.process(pcm); eng.delete(); print(prec, int(np.abs(out).sum()))
PY
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02n_sanitizer_racecheck.log python gpurun_out/san_small.py > gpurun_out/r02n_sanitizer_racecheck.out 2>&1; tail -3 gpurun_out/r02n_sanitizer_racecheck.log
echo "== compute-sanitizer initcheck"
timeout 900 compute-sanitizer --tool initcheck --log-file gpurun_out/r02n_sanitizer_initcheck.log python gpurun_out/san_small.py > gpurun_out/r02n_sanitizer_initcheck.out 2>&1; tail -3 gpurun_out/r02n_sanitizer_initcheck.log
