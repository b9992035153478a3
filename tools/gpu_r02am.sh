#!/bin/bash
# Driver-style two-GPU invocation (torchrun, --steps 20 --warmup 5), both arms.
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench_r02am_2gpu.err | tail -1 > gpurun_out/bench_r02am_2gpu.json; tail -2 gpurun_out/bench_r02am_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02am_2gpu.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['repetitions'],'frac',d['roofline']['frac'], d['n_gpus'], d['config']['total_streams'], d['clocks'])
print(d['host_link']['e2e_fraction_of_ceiling'], d['host_link']['e2e_fraction_of_slowest_rank_ceiling'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -2
