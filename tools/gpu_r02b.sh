#!/bin/bash
# Round 2, time-persistent fused kernel: smoke, the chunked-launch tests, the whole GPU suite, a short bench line.
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== chunked"; timeout 900 python -m pytest tests/test_gpu_chunked.py -x -q 2>&1 | tail -15
echo "== pytest";  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench";   timeout 900 python bench.py --steps 640 --cpu-seconds 4 2> gpurun_out/bench_r02b.err | tail -1 > gpurun_out/bench_r02b_default.json; tail -3 gpurun_out/bench_r02b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02b_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['avg_launch_ms'], 'one-frame', d.get('one_frame_per_call'))
print(d['kernel_ms_per_step'], d['clocks'])
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
