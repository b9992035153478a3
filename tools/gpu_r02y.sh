#!/bin/bash
mkdir -p gpurun_out
echo "== pytest";  timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
L=koala_b200/lib/libpv_koala_b200.so
FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
FPC=32 timeout 200 python tools/variant_bench.py $L 8192 60
FPC=64 timeout 200 python tools/variant_bench.py $L 128 60
FPC=1 timeout 200 python tools/variant_bench.py $L 8192 200
FPC=16 timeout 200 python tools/variant_bench.py $L 2048 100
