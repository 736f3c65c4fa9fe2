#!/bin/bash
# round-2 GPU session 5 (1 GPU): kernel / step tests after the packed layer-1 weight gradient and the 128-wide decoder tiles,
# launch list, bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_baseline_sizes.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2_t5.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t5.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_B256_v3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_B256_v3.csv > gpurun_out/r2_launch_shares_step_B256_v3.txt 2>&1
head -24 gpurun_out/r2_launch_shares_step_B256_v3.txt
python bench.py --steps 20 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_bench5.log 2>&1; echo "bench rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench5.log | head -1)"
for b in 32 64 128 512; do
python bench.py --steps 20 --warmup 5 --no-cpu --no-infer --batch $b > gpurun_out/r2_bench5_b$b.log 2>&1; echo "B=$b $(grep -o '"value": [0-9.]*' gpurun_out/r2_bench5_b$b.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench5_b$b.log | head -1)"
done
