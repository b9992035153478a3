#!/bin/bash
# Round 2, eight GPUs (final build): BASELINE configs[3] (65 536 streams) under torchrun, the reference arm, and configs[4] (1024 streams x 10-minute clips).
mkdir -p gpurun_out
echo "== bench N=8"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 2000 --warmup 32 2> gpurun_out/bench_r02v_8gpu.err | tail -1 > gpurun_out/bench_r02v_8gpu.json; tail -3 gpurun_out/bench_r02v_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02v_8gpu.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac']); print(d.get('host_link')); print(d['clocks'])
PY
echo "== cfg5 N=8"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload cfg5_128_per_gpu_bf16 --steps 37504 --warmup 64 --e2e-steps 512 2> gpurun_out/bench_r02v_cfg5_8gpu.err | tail -1 > gpurun_out/bench_r02v_cfg5_8gpu.json; tail -2 gpurun_out/bench_r02v_cfg5_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02v_cfg5_8gpu.json'))
print('cfg5 value',d['value'],'ms/step',d['ms_per_step'],'rtf_x',d['rtf_x'],'e2e',d['e2e']['value'], d['config']['total_streams'])
PY
