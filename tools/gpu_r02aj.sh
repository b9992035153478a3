#!/bin/bash
# The driver's own invocation: python3 bench.py --gpus 1 --steps 20 --warmup 5 (and the reference arm the same way).
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/bench_r02aj.err | tail -1 > gpurun_out/bench_r02aj_$i.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02aj_$i.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'), d['clocks'], d['gpu_launches'], d['process_calls'])
PY
done
timeout 600 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-300
