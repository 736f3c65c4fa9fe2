#!/bin/bash
# `ncu --set full` of the PointNet kernels of ONE open-set inference step (eval forward: layer-1 kernel, two affine+ELU
# GEMMs, the pooled last-layer GEMM); run under gpurun.  usage: scripts/ncu_infer.sh <tag>
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'pointnet_l1_fwd_t|gemm_tc_kernel' \
    --launch-skip 30 --launch-count 4 -f -o /tmp/infer_$TAG \
    python bench.py --workload infer --steps 1 --warmup 3 --no-cpu > gpurun_out/infer_$TAG.log 2>&1
ncu -i /tmp/infer_$TAG.ncu-rep --page raw --csv > gpurun_out/infer_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/infer_${TAG}*
