#!/bin/bash
# round-2 GPU session 7 (2 GPUs): data-parallel tests (incl. SyncBN with identical shards), reproducibility probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -rP > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?"
grep "^dp_parity" gpurun_out/r2_t7.log | cut -c1-600
tail -2 gpurun_out/r2_t7.log
python scripts/noise_probe.py 8 50 > gpurun_out/r2_noise_probe.txt 2>&1
python scripts/noise_probe.py 256 150 >> gpurun_out/r2_noise_probe.txt 2>&1
cat gpurun_out/r2_noise_probe.txt
