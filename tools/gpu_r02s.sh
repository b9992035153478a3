#!/bin/bash
# Round 2: big batches as L2-sized partitions -- whole GPU suite, default bench line.
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench";   timeout 900 python bench.py --cpu-seconds 8 2> gpurun_out/bench_r02s.err | tail -1 > gpurun_out/bench_r02s_default.json; tail -3 gpurun_out/bench_r02s.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02s_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'), d['roofline']['kernel'])
print(d.get('kernel_ms_per_step'), d['clocks'], d['gpu_launches'], d['process_calls'], d['one_frame_per_call']); print(d.get('host_link'))
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','steps','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
