#!/bin/bash
# round-2 GPU session 16 (2 GPUs): decoder span's Adam update sharded over the ranks (reduce-scatter, update 1/world, all-gather weights)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dp.py -m gpu -q -x -rP > gpurun_out/r2_t16.log 2>&1; echo "pytest dp rc=$?"
grep "^dp_parity" gpurun_out/r2_t16.log | cut -c1-400; tail -2 gpurun_out/r2_t16.log
timeout 600 python -m pytest tests/test_gpu_step.py -m gpu -q -x > gpurun_out/r2_t16b.log 2>&1; echo "pytest step rc=$?"; tail -1 gpurun_out/r2_t16b.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29821 bench.py --gpus 2 --steps 30 --warmup 5 --no-infer --phases > gpurun_out/r2_dp2_shard1.log 2>&1
echo "dp2 sharded rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_shard1.log | head -1) $(grep -o '"ok": [a-z]*' gpurun_out/r2_dp2_shard1.log | head -1)"
PCAA_DP_SHARD_ADAM=0 timeout 600 $TR --master-port 29822 bench.py --gpus 2 --steps 30 --warmup 5 --no-infer --no-dp-parity > gpurun_out/r2_dp2_shard0.log 2>&1
echo "dp2 all-reduce rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_shard0.log | head -1)"
timeout 600 $TR --master-port 29823 bench.py --gpus 2 --steps 30 --warmup 5 --no-infer --no-dp-parity > gpurun_out/r2_dp2_shard1b.log 2>&1
echo "dp2 sharded (2nd) rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_shard1b.log | head -1)"
tail -c 1500 gpurun_out/r2_dp2_shard1.log
