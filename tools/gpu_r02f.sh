#!/bin/bash
mkdir -p gpurun_out
echo "== chunked"; timeout 900 python -m pytest tests/test_gpu_chunked.py -x -q 2>&1 | tail -3
echo "== variants"
timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 100
timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 100
FPC=1 timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 300
FPC=64 timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 128 100
FPC=32 timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 4096 100
echo "== trace steady"; timeout 200 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 60 2>&1 | head -24 | tee gpurun_out/trace_r02f_steady.txt
