#!/bin/bash
# Wave quantisation of the STFT kernels: frontend grid 911 CTAs (888 fit at once) vs 886.
L=koala_b200/lib/libpv_koala_b200.so
FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_STFT_PER_WARP=37 FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_STFT_PER_WARP=40 FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_STFT_PER_WARP=19 FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_STFT_PER_WARP=3 FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
KOALA_STFT_PER_WARP=1 FPC=32 timeout 200 python tools/variant_bench.py $L 4096 60
