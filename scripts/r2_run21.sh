#!/bin/bash
# round-2 GPU session 21 (1 GPU): A/B of PCAA_LATE_UPDATE (decoder Adam forked behind the HBM-bound head of the encoder backward)
mkdir -p gpurun_out
PCAA_LATE_UPDATE=1 timeout 200 python -m pytest tests/test_gpu_step.py -m gpu -q -x -k "oracle or graphed or golden" > gpurun_out/r2_t21.log 2>&1; echo "pytest (late update) rc=$?"; tail -1 gpurun_out/r2_t21.log
for i in 1 2 3 4; do
  v=$(( i % 2 ))
  PCAA_LATE_UPDATE=$v timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_late_ab_${i}_late$v.log 2>&1
  echo "run $i late=$v rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_late_ab_${i}_late$v.log | head -1) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/r2_late_ab_${i}_late$v.log | head -1)"
done
