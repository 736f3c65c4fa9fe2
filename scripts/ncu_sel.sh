#!/bin/bash
# metrics-only `ncu --set full` of selected gemm_tc launches: usage ncu_sel.sh <tag> <skip> <count>
TAG=$1; SKIP=$2; COUNT=$3
ncu --set full --clock-control none -k regex:'gemm_tc_kernel' --launch-skip $SKIP --launch-count $COUNT -f -o /tmp/sel_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/sel_$TAG.log 2>&1
ncu -i /tmp/sel_$TAG.ncu-rep --page raw --csv > gpurun_out/sel_${TAG}_raw.csv 2>/dev/null
