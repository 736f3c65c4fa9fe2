#!/bin/bash
# round-2 GPU session 9 (1 GPU): tests after the vectorised TCN kernels, ncu --set full with source of the eval layer-1 kernel,
# inference launch list + bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_inference.py tests/test_gpu_baseline_sizes.py tests/test_gpu_blocks_and_convergence.py -m gpu -q > gpurun_out/r2_t9.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t9.log
ncu --set full --clock-control none --import-source on -k regex:'pointnet_l1_fwd_bn_t' --launch-skip 3 --launch-count 1 -f -o /tmp/src_l1 \
    python bench.py --workload infer --steps 1 --warmup 3 --no-cpu > gpurun_out/src_l1.log 2>&1
ncu -i /tmp/src_l1.ncu-rep --page source --csv > gpurun_out/src_l1_source.csv 2>/dev/null
ncu -i /tmp/src_l1.ncu-rep --page raw --csv > gpurun_out/src_l1_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/src_l1_raw.csv
python scripts/ncu_source_top.py gpurun_out/src_l1_source.csv 0 25 > gpurun_out/src_l1_top.txt 2>&1; head -40 gpurun_out/src_l1_top.txt | cut -c1-220
python bench.py --workload infer --steps 100 --warmup 5 --no-cpu > gpurun_out/r2_infer9.log 2>&1; echo "infer $(grep -o '"value": [0-9.]*' gpurun_out/r2_infer9.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_infer9.log | head -1)"
