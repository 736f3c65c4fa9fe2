#!/bin/bash
# `ncu --set full` with source correlation of the PointNet backward GEMMs (dgrad L4, wgrad L3, dgrad L3, wgrad L2, dgrad L2)
TAG=${1:-r1}; SKIP=${2:-43}; COUNT=${3:-5}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel' \
    --launch-skip $SKIP --launch-count $COUNT -f -o gpurun_out/dgrad_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/dgrad_$TAG.log 2>&1
ncu -i gpurun_out/dgrad_$TAG.ncu-rep --page source --csv > gpurun_out/dgrad_${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/dgrad_$TAG.ncu-rep --page raw --csv > gpurun_out/dgrad_${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/
