#!/bin/bash
# Does a 4096-stream batch run faster per frame than 8192 because of L2 residency or because of the longer launch?
L=koala_b200/lib/libpv_koala_b200.so
FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_CHUNK_FRAMES=16 FPC=16 timeout 200 python tools/variant_bench.py $L 4096 120
FPC=64 timeout 200 python tools/variant_bench.py $L 2048 40
KOALA_CHUNK_FRAMES=16 FPC=16 timeout 200 python tools/variant_bench.py $L 2048 160
FPC=16 timeout 200 python tools/variant_bench.py $L 8192 100
FPC=16 timeout 200 python tools/variant_bench.py $L 6144 100
FPC=16 timeout 200 python tools/variant_bench.py $L 16384 50
