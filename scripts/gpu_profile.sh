#!/bin/bash
# ncu launch list of one eager step + full captures of the dominant kernels (run under gpurun; outputs in gpurun_out/)
set -x
TAG=${1:-cur}
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-roofline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/prof_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 150 -c 3 -o gpurun_out/halo_${TAG} -f $B >> gpurun_out/prof_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_row -s 10 -c 3 -o gpurun_out/wgrad_${TAG} -f $B >> gpurun_out/prof_${TAG}.log 2>&1
ls -la gpurun_out
