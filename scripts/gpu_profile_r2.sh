#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE GPU; outputs in gpurun_out/):
#   1. launch list of one eager headline step;
#   2. --set full of the front end (voxeliser, PillarVFE), the Where2comm mask / fusion kernels and the dominant GEMM kernels
#      inside the headline step;
#   3. --set full of the kernels the headline step does not run at full size (ego-warp, mask compaction, AttentionFusion).
set -x
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-roofline --no-extra"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2_launches_final.csv $B > gpurun_out/r2_prof.log 2>&1
# the bench runs 5 eager steps; ~33 front-end / mask kernels per step: skip the first 4 steps
ncu --set full --clock-control none --import-source on -k regex:"pfn_|vox_|vr_|topk_mask|gauss_mask|conf_map|att_fuse" -s 140 -c 40 \
    -o gpurun_out/r2_frontend -f $B >> gpurun_out/r2_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tapgemm" -s 330 -c 6 -o gpurun_out/r2_tapgemm -f $B >> gpurun_out/r2_prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"warp_affine|mask_compact|mask_decompact|att_fuse" -s 8 -c 8 \
    -o gpurun_out/r2_small -f python scripts/ncu_small_kernels.py >> gpurun_out/r2_prof.log 2>&1
ls -la gpurun_out/r2_*
