#!/bin/bash
# round-2 GPU session 3 (2 GPUs): data-parallel parity tests, 2-GPU bench with parity + phases, single-graph capture experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_reference_scripts.py -m gpu -q -rP > gpurun_out/r2_t3.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_t3.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-infer --phases > gpurun_out/r2_dp2_split.log 2>&1
echo "dp2 split rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_split.log | head -1)"
PCAA_DP_ONE_GRAPH=1 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-infer > gpurun_out/r2_dp2_onegraph.log 2>&1
echo "dp2 onegraph rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_onegraph.log | head -1)"
PCAA_DP_EXCHANGE=nccl timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-infer > gpurun_out/r2_dp2_nccl.log 2>&1
echo "dp2 nccl rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_nccl.log | head -1)"
python bench.py --steps 20 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_dp1.log 2>&1
echo "dp1 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp1.log | head -1)"
tail -c 600 gpurun_out/r2_dp2_onegraph.log
