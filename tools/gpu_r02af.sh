#!/bin/bash
# Timing experiments (wrong results): the fused kernel without its weight loads / without its activation operand loads.
L=koala_b200/lib/libpv_koala_b200.so
FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
FPC=32 timeout 200 python tools/variant_bench.py gpurun_lib_SKIPB.so 4096 60
FPC=32 timeout 200 python tools/variant_bench.py gpurun_lib_SKIPA.so 4096 60
FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
