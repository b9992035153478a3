#!/bin/bash
# old-build comparison of the cfg5 test, new trace of the multi-step kernel, full test suite
mkdir -p gpurun_out
echo "== old build cfg5"; (cd _old && timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k config5 2>&1 | tail -4)
echo "== trace steady"; timeout 200 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 60 66 2>&1 | tail -32 | tee gpurun_out/trace_r02c_steady.txt
echo "== trace head"; timeout 200 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 0 2>&1 | head -16 | tee gpurun_out/trace_r02c_head.txt
echo "== trace small"; timeout 200 python tools/gpu_trace.py 128 gpurun_lib_TRACE.so 64 4 2>&1 | head -24 | tee gpurun_out/trace_r02c_small.txt
echo "== pytest";  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
