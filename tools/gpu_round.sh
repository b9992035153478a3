#!/bin/bash
# One GPU call that refreshes everything under profiles/ for the current build: tests, smoke, the default bench line, the other
# BASELINE workloads, the ncu launch list, one full-set capture of the fused kernel and of the STFT kernels, a clock64 trace.
TAG=${1:-r01c}
mkdir -p gpurun_out
echo "== pytest";  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke";   timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench";   timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_${TAG}_default.json; cut -c1-400 gpurun_out/bench_${TAG}_default.json
for w in cfg3_4096_bf16 cfg2_256_fp32 cfg5_128_per_gpu_bf16; do
  timeout 300 python bench.py --workload $w --steps 500 --no-cpu-baseline --e2e-steps 32 2>&1 | tail -1 > gpurun_out/bench_${TAG}_$w.json
  python -c "import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 8"
echo "== ncu launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 9 -c 30 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu full";     timeout 600 ncu --set full --clock-control none --import-source on -s 9 -c 3 -f -o gpurun_out/step_${TAG} $CMD > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-120
echo "== trace"; [ -f gpurun_lib_TRACE.so ] || python -m koala_b200._build -DKOALA_FU_TRACE=1 -ogpurun_lib_TRACE.so > /dev/null 2>&1
timeout 100 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 2>&1 | head -14 > gpurun_out/trace_${TAG}.txt; head -3 gpurun_out/trace_${TAG}.txt
