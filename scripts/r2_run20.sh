#!/bin/bash
# round-2 GPU session 20 (1 GPU): full GPU suite on the final tree
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_t20.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t20.log
