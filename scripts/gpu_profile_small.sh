#!/bin/bash
# ncu --set full captures of the non-GEMM kernels (run under gpurun; outputs in gpurun_out/)
TAG=${1:-cur}
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-roofline"
for k in pfn_scatter pfn_moments bn_train_act channel_reduce bn_relu_bwd_apply; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/${k}_${TAG} -f $B > gpurun_out/prof_small_${TAG}.log 2>&1
done
ls gpurun_out | grep ${TAG}
