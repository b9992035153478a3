#!/bin/bash
mkdir -p gpurun_out
echo "== chunked"; timeout 900 python -m pytest tests/test_gpu_chunked.py -x -q 2>&1 | tail -5
echo "== trace steady"; timeout 200 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 60 2>&1 | head -24 | tee gpurun_out/trace_r02d_steady.txt
echo "== variant"; timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 100
FPC=1 timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 300
FPC=64 timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 128 100
echo "== pytest";  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
