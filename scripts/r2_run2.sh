#!/bin/bash
# round-2 GPU session 2: whole GPU test tier, A/B benches of the fused paths, ncu launch list of one step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP > gpurun_out/r2_t2.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t2.log
for cfg in "default:" "tcn_unfused:PCAA_TCN_FUSED=0" "split:PCAA_SPLIT_GRAPHS=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --steps 20 --warmup 5 --no-cpu --no-infer --phases > gpurun_out/r2_ab_$name.log 2>&1
  echo "$name rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_ab_$name.log | head -1)"
done
python bench.py --steps 20 --warmup 5 --no-cpu --no-infer --batch 32 > gpurun_out/r2_ab_b32.log 2>&1
echo "b32 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_ab_b32.log | head -1)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_B256_v1.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_B256_v1.csv > gpurun_out/r2_launch_shares_step_B256_v1.txt 2>&1
head -30 gpurun_out/r2_launch_shares_step_B256_v1.txt
