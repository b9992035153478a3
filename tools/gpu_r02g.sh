#!/bin/bash
mkdir -p gpurun_out
L=koala_b200/lib/libpv_koala_b200.so
echo "== variants"
timeout 200 python tools/variant_bench.py $L 8192 100
timeout 200 python tools/variant_bench.py gpurun_lib_NOPF.so 8192 100
KOALA_E_RING=2 timeout 200 python tools/variant_bench.py $L 8192 100
KOALA_E_RING=2 timeout 200 python tools/variant_bench.py gpurun_lib_NOPF.so 8192 100
KOALA_CHUNK_FRAMES=8 FPC=8 timeout 200 python tools/variant_bench.py $L 8192 200
KOALA_CHUNK_FRAMES=32 FPC=32 timeout 200 python tools/variant_bench.py $L 8192 50
KOALA_FU_XFIRST=1 timeout 200 python tools/variant_bench.py $L 8192 100
timeout 200 python tools/variant_bench.py $L 8192 100
