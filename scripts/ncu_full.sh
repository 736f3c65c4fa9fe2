#!/bin/bash
# `ncu --set full` captures of the PointNet kernels of ONE train step (B200_PROFILING.md recipe); run under gpurun.
# usage: scripts/ncu_full.sh <tag> [skip] [count]
TAG=${1:-r1}; SKIP=${2:-132}; COUNT=${3:-33}
mkdir -p gpurun_out
# (1) every PointNet kernel of the step, metrics only (raw page -> csv; the report itself is dropped: too large to ship)
ncu --set full --clock-control none \
    -k regex:'gemm_tc_kernel|meanpool|bn_elu_apply_t|pool_bwd_apply|bn_bwd_apply_t|pointnet_l1' \
    --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/full_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/full_$TAG.log 2>&1
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_raw.csv 2>/dev/null
# (2) the data-gradient GEMM with source correlation (3 launches of the step)
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel<256, 1, 1, 9>' \
    --launch-skip 9 --launch-count 3 -f -o gpurun_out/dgrad_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/dgrad_$TAG.log 2>&1
ncu -i gpurun_out/dgrad_$TAG.ncu-rep --page source --csv > gpurun_out/dgrad_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/ /tmp/full_$TAG.ncu-rep
