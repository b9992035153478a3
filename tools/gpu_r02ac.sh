#!/bin/bash
mkdir -p gpurun_out
echo "== fixed-point tests";  timeout 900 python -m pytest tests/test_gpu_fixed_point.py -m gpu -x -q 2>&1 | tail -4
for n in 4096 8192; do
timeout 300 python bench.py --workload fixed_point_4096_int8 --streams $n --steps 1024 --no-cpu-baseline --no-others --e2e-steps 64 2>gpurun_out/err_i8.txt | tail -1 > gpurun_out/bench_r02ac_fixed_point_$n.json
python -c "import json; d=json.load(open('gpurun_out/bench_r02ac_fixed_point_$n.json')); print('int8 $n value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['kernel_ms_per_step'])" || tail -5 gpurun_out/err_i8.txt
done
