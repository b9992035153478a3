#!/bin/bash
# Round 2 re-entry: state of the time-persistent kernel build: smoke, whole GPU suite, bench line, ncu launch list.
mkdir -p gpurun_out
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench";   timeout 900 python bench.py --cpu-seconds 6 2> gpurun_out/bench_r02h.err | tail -1 > gpurun_out/bench_r02h_default.json; tail -3 gpurun_out/bench_r02h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02h_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'))
print(d.get('kernel_ms_per_step'), d['clocks'])
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
echo "== launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 32 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 8 > gpurun_out/r02h_ncu_bench.log 2>&1; tail -2 gpurun_out/r02h_ncu_bench.log | cut -c1-300
