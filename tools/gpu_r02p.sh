#!/bin/bash
# Round 2: host-ingest chunk scaling for small batches, fixed-point epilogue parameters in shared memory.
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for w in cfg5_128_per_gpu_bf16 cfg2_256_fp32 fixed_point_4096_int8; do
  timeout 300 python bench.py --workload $w --steps 2048 --no-cpu-baseline --no-others --e2e-steps 512 2>gpurun_out/err_$w.txt | tail -1 > gpurun_out/bench_r02p_$w.json
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_r02p_$w.json')); print('$w', 'value', d['value'], 'us/step', d['ms_per_step']*1e3, 'e2e', d['e2e']['value'], d['e2e'].get('stream_major_call_value_rank0'), d['kernel_ms_per_step'])" || tail -5 gpurun_out/err_$w.txt
done
