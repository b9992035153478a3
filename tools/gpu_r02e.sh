#!/bin/bash
mkdir -p gpurun_out
echo "== trace (clock calibration)"; timeout 200 python tools/gpu_trace.py 8192 gpurun_lib_TRACE.so 16 60 2>&1 | head -4
echo "== variants"
for i in 1 2; do
timeout 200 python tools/variant_bench.py koala_b200/lib/libpv_koala_b200.so 8192 100
timeout 200 python tools/variant_bench.py gpurun_lib_FAKEKB.so 8192 100
done
echo "== cfg5 test"; timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -x -q -k config5 2>&1 | tail -3
