#!/bin/bash
# Round 2: ncu full-set capture of one steady-state chunk (analysis / fused mask estimator over 16 steps / synthesis) + clock64 trace.
mkdir -p gpurun_out
CMD="python bench.py --steps 64 --warmup 16 --no-cpu-baseline --no-others --e2e-steps 8"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -s 9 -c 3 -f -o gpurun_out/step_r02i $CMD > gpurun_out/ncu_full_r02i.log 2>&1; tail -1 gpurun_out/ncu_full_r02i.log | cut -c1-160
ls -la gpurun_out/*.ncu-rep
echo "== trace"; timeout 100 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 40 44 2>&1 | head -40 > gpurun_out/trace_r02i.txt; head -30 gpurun_out/trace_r02i.txt
