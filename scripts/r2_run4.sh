#!/bin/bash
# round-2 GPU session 4 (1 GPU): whole GPU test tier, default bench line, ncu launch list + --set full capture of one step
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rP > gpurun_out/r2_t11.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t11.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench11.log 2>&1; echo "bench rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench11.log | head -1)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_B256_v5.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_B256_v5.csv > gpurun_out/r2_launch_shares_step_B256_v5.txt 2>&1
head -5 gpurun_out/r2_launch_shares_step_B256_v5.txt
# one eager step = 127 launches, 55 of them match the filter; 3 warm-up steps precede the timed one: capture that 4th step
ncu --set full --clock-control none \
    -k regex:'gemm_tc_kernel|meanpool|bn_elu_apply_rows|pool_bwd_apply|bn_bwd_apply_t|pointnet_l1|adam_flat|chamfer_fwd|chamfer_bwd' \
    --launch-skip ${SKIP:-165} --launch-count ${COUNT:-55} -f -o /tmp/full_r2v3 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/full_r2v3.log 2>&1
ncu -i /tmp/full_r2v3.ncu-rep --page raw --csv > gpurun_out/full_r2v3_raw.csv 2>/dev/null
python scripts/ncu_raw_summary.py gpurun_out/full_r2v3_raw.csv > gpurun_out/full_r2v3_summary.txt 2>&1
python scripts/roofline_table.py gpurun_out/full_r2v3_raw.csv --tag r2v3 > gpurun_out/r2_roofline_table.md 2>&1
tail -30 gpurun_out/r2_roofline_table.md
