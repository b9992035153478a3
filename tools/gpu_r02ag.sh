#!/bin/bash
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
L=koala_b200/lib/libpv_koala_b200.so
for i in 1 2; do FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60; done
FPC=32 timeout 200 python tools/variant_bench.py $L 8192 60
FPC=64 timeout 200 python tools/variant_bench.py $L 128 60
echo "== trace"; timeout 100 python tools/gpu_trace.py 4096 gpurun_lib_TRACE.so 32 40 2>&1 | head -16
