#!/bin/bash
# Eight GPUs, final build: driver-style (--steps 20 --warmup 5) and the longer default run; configs[4] under torchrun.
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" 2> gpurun_out/$2.err | tail -1 > gpurun_out/$2.json
python - <<PY
import json
d=json.load(open('gpurun_out/$2.json'))
print('$2: value',d['value'],'ms/step',d['ms_per_step'],'rtf_x',d['rtf_x'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['config']['total_streams'], d['clocks']['reasons'], round(d['host_link']['e2e_fraction_of_slowest_rank_ceiling'],3))
PY
}
run 29551 bench_r02an_8gpu_driver_style --steps 20 --warmup 5
run 29552 bench_r02an_8gpu --steps 2000 --warmup 32
run 29553 bench_r02an_cfg5_8gpu --workload cfg5_128_per_gpu_bf16 --steps 37504 --warmup 64 --e2e-steps 512
