#!/bin/bash
# `ncu --set full` with source correlation for selected gemm_tc launches: usage ncu_src.sh <tag> <skip> <count>
TAG=$1; SKIP=$2; COUNT=$3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel' \
    --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/src_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/src_$TAG.log 2>&1
ncu -i /tmp/src_$TAG.ncu-rep --page source --csv > gpurun_out/src_${TAG}_source.csv 2>/dev/null
ncu -i /tmp/src_$TAG.ncu-rep --page raw --csv > gpurun_out/src_${TAG}_raw.csv 2>/dev/null
