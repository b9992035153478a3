#!/bin/bash
mkdir -p gpurun_out
echo "== fixed-point tests";  timeout 900 python -m pytest tests/test_gpu_fixed_point.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --workload fixed_point_4096_int8 --steps 1024 --no-cpu-baseline --no-others --e2e-steps 128 2>gpurun_out/err_i8.txt | tail -1 > gpurun_out/bench_r02w_fixed_point.json
python -c "import json; d=json.load(open('gpurun_out/bench_r02w_fixed_point.json')); print('int8 4096 value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['kernel_ms_per_step'])" || tail -5 gpurun_out/err_i8.txt
echo "== bench (default, short)"; timeout 900 python bench.py --cpu-seconds 4 2> gpurun_out/bench_r02w.err | tail -1 > gpurun_out/bench_r02w_default.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02w_default.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','us_per_call','steps','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
