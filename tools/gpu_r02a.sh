#!/bin/bash
# Round 2, first GPU call: the new parity tests at BASELINE sizes on the unchanged kernels, the restructured bench line, and
# the fused kernel alone (L2-warm, no STFT kernels around it).
mkdir -p gpurun_out
echo "== pytest";  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== smoke";   timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench";   timeout 900 python bench.py 2> gpurun_out/bench_r02a.err | tail -1 > gpurun_out/bench_r02a_default.json; cut -c1-600 gpurun_out/bench_r02a_default.json; tail -3 gpurun_out/bench_r02a.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02a_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline']['peak'])
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
echo "== fused only"; KOALA_B200_ONLY_MASKNET=1 timeout 300 python bench.py --steps 500 --no-cpu-baseline --no-others --e2e-steps 8 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused-only ms/step', d['ms_per_step'], d['kernel_ms_per_step'])"
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 3 2>&1 | tail -1 | cut -c1-300
