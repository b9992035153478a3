#!/bin/bash
# Round 2: fp32-on-tensor-cores tests + default bench line with the host-link probe.
mkdir -p gpurun_out
echo "== pytest (chunked + parity)";  timeout 1800 python -m pytest tests/test_gpu_chunked.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
echo "== bench";   timeout 900 python bench.py --cpu-seconds 6 2> gpurun_out/bench_r02k.err | tail -1 > gpurun_out/bench_r02k_default.json; tail -3 gpurun_out/bench_r02k.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02k_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'))
print(d.get('kernel_ms_per_step'), d['clocks']); print(d.get('host_link'))
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
