#!/bin/bash
# Round 2, two GPUs: the sharded GPU test and the bench line under torchrun (host-link probe with both ranks copying at once).
mkdir -p gpurun_out
echo "== multirank test"; timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1000 --warmup 32 2> gpurun_out/bench_r02l_2gpu.err | tail -1 > gpurun_out/bench_r02l_2gpu.json; tail -3 gpurun_out/bench_r02l_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02l_2gpu.json'))
print('value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac']); print(d.get('host_link')); print(d['clocks'])
PY
echo "== reference arm N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
