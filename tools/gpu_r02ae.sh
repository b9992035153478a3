#!/bin/bash
# Timing experiment (wrong data): both GEMM operands fetched as contiguous 16 KB / 12 KB boxes of a K-blocked view, on the r02g build.
for i in 1 2; do
FPC=16 timeout 200 python tools/variant_bench.py gpurun_lib_OLD.so 8192 100
FPC=16 timeout 200 python tools/variant_bench.py gpurun_lib_OLDFAKEKB.so 8192 100
done
FPC=32 timeout 200 python tools/variant_bench.py gpurun_lib_OLD.so 4096 60
FPC=32 timeout 200 python tools/variant_bench.py gpurun_lib_OLDFAKEKB.so 4096 60
