#!/bin/bash
# round-2 GPU session 18 (2 GPUs): the data-parallel test file on the final tree (three exchanges + SyncBN)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q -rP > gpurun_out/r2_t18.log 2>&1; echo "pytest dp rc=$?"
grep "^dp_parity" gpurun_out/r2_t18.log | cut -c1-330; tail -2 gpurun_out/r2_t18.log
