#!/bin/bash
# full-set ncu capture of the GRU tensor-core kernel only (2 launches of a steady-state step), with source counters
mkdir -p gpurun_out
TAG=${1:-gru}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_masknet_kernel -s 12 -c 4 -f -o gpurun_out/${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log | cut -c1-300
ls -la gpurun_out/${TAG}.ncu-rep
