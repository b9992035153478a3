#!/bin/bash
# One GPU call that refreshes everything under profiles/ for the current build (run under gpurun on one B200):
#   bash tools/gpu_round.sh r03a
# tests, smoke, the default bench line (all BASELINE workloads + the fixed-point variant under `others`), the reference arm, the ncu
# launch list, one full-set capture of a steady-state chunk (-> profiles/traffic.json via tools/ncu_traffic.py, run afterwards where
# the .ncu-rep landed), a clock64 trace of the fused kernel, compute-sanitizer memcheck.  Copy what should be judged from
# gpurun_out/ to profiles/.
TAG=${1:-r03a}
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke";   timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench";   timeout 900 python bench.py 2> gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}_default.json; cut -c1-400 gpurun_out/bench_${TAG}_default.json
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-300
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 32 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 8 > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "== ncu full";     timeout 900 ncu --set full --clock-control none --import-source on -s 9 -c 3 -f -o gpurun_out/step_${TAG} \
    python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-others --e2e-steps 8 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-120
echo "== trace"; [ -f gpurun_lib_TRACE.so ] || python -m koala_b200._build -DKOALA_FU_TRACE=1 -ogpurun_lib_TRACE.so > /dev/null 2>&1
timeout 100 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 40 2>&1 | head -24 > gpurun_out/trace_${TAG}.txt; head -3 gpurun_out/trace_${TAG}.txt
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_${TAG}.log python __graft_entry__.py smoke > /dev/null 2>&1; tail -1 gpurun_out/sanitizer_memcheck_${TAG}.log
