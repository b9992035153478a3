#!/bin/bash
# ncu captures for profiles/: (1) per-launch durations of the bench command, (2) full-set capture of one steady-state step.
mkdir -p gpurun_out
TAG=${1:-r01}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 18 -c 6 -f -o gpurun_out/step_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/ncu_full_${TAG}.log
