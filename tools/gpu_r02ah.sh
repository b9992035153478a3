#!/bin/bash
# Round 2, final build: tests, smoke, default bench line, reference arm, ncu launch list, full-set capture of one steady-state chunk, trace, memcheck.
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench";   timeout 900 python bench.py 2> gpurun_out/bench_r02ah.err | tail -1 > gpurun_out/bench_r02ah_default.json; tail -2 gpurun_out/bench_r02ah.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02ah_default.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],d['roofline'].get('avg_launch_ms'), d['roofline']['kernel'], d['roofline']['traffic_source'])
print(d.get('kernel_ms_per_step'), d['clocks'], d['gpu_launches'], d['process_calls'], d['one_frame_per_call']); print(d.get('host_link')); print(d['cpu_baseline'])
for k,v in (d.get('others') or {}).items(): print(k, {x:v.get(x) for x in ('value','ms_per_step','steps','error')}, (v.get('e2e') or {}).get('value'), (v.get('roofline') or {}).get('frac'))
PY
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02ah_launches.csv \
    python bench.py --steps 128 --warmup 32 --no-cpu-baseline --no-others --e2e-steps 8 > gpurun_out/ncu_launch_r02ah.log 2>&1
echo "== ncu full";     timeout 900 ncu --set full --clock-control none --import-source on -s 15 -c 6 -f -o gpurun_out/step_r02ah \
    python bench.py --steps 128 --warmup 32 --no-cpu-baseline --no-others --e2e-steps 8 > gpurun_out/ncu_full_r02ah.log 2>&1; tail -1 gpurun_out/ncu_full_r02ah.log | cut -c1-120
echo "== trace"; timeout 100 python tools/gpu_trace.py 4096 gpurun_lib_TRACE.so 32 40 2>&1 | head -24 > gpurun_out/trace_r02ah.txt; head -4 gpurun_out/trace_r02ah.txt
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02ah_sanitizer_memcheck.log python __graft_entry__.py smoke > /dev/null 2>&1; tail -1 gpurun_out/r02ah_sanitizer_memcheck.log
