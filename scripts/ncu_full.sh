#!/bin/bash
# `ncu --set full` of the hot kernels of ONE train step (B200_PROFILING.md recipe); run under gpurun.  The report is too
# large to ship (> 64 MiB): its raw page is exported to csv and condensed by scripts/ncu_raw_summary.py.
# usage: scripts/ncu_full.sh <tag> [skip] [count]
TAG=${1:-r1}; SKIP=${2:-168}; COUNT=${3:-56}
mkdir -p gpurun_out
ncu --set full --clock-control none \
    -k regex:'gemm_tc_kernel|meanpool|bn_elu_apply_rows|pool_bwd_apply|bn_bwd_apply_t|pointnet_l1|adam_flat|chamfer_fwd|chamfer_bwd' \
    --import-source on \
    --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu --graph off > gpurun_out/full_$TAG.log 2>&1
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/full_${TAG}*
