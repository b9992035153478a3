#!/bin/bash
# Round 2: ncu full-set captures of the other numeric modes' kernels (fixed-point layers at 4096 streams; fp32-planes fused kernel at 256 streams).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:i8_layer -s 8 -c 4 -f -o gpurun_out/i8_r02q python tools/ncu_modes.py int8 4096 4 > gpurun_out/ncu_i8.log 2>&1; tail -1 gpurun_out/ncu_i8.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -f -o gpurun_out/fp32_r02q python tools/ncu_modes.py fp32 256 64 > gpurun_out/ncu_fp32.log 2>&1; tail -1 gpurun_out/ncu_fp32.log | cut -c1-120
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_fused -s 2 -c 1 -f -o gpurun_out/cfg5_r02q python tools/ncu_modes.py bf16 128 64 > gpurun_out/ncu_cfg5.log 2>&1; tail -1 gpurun_out/ncu_cfg5.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
