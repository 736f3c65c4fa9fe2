#!/bin/bash
# round-2 GPU session 13 (2 GPUs): encoder gradient exchange in two buckets (upper part overlapped with the layer 2/1 backward)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_step.py -m gpu -q -x > gpurun_out/r2_t13.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_t13.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29811 bench.py --gpus 2 --steps 40 --warmup 5 --no-infer --phases > gpurun_out/r2_dp2_bucket.log 2>&1
echo "dp2 split rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_bucket.log | head -1)"
PCAA_DP_ONE_GRAPH=1 timeout 300 $TR --master-port 29812 bench.py --gpus 2 --steps 40 --warmup 5 --no-infer > gpurun_out/r2_dp2_bucket_onegraph.log 2>&1
echo "dp2 onegraph rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_bucket_onegraph.log | head -1)"
PCAA_DP_EXCHANGE=nccl timeout 600 $TR --master-port 29813 bench.py --gpus 2 --steps 40 --warmup 5 --no-infer > gpurun_out/r2_dp2_bucket_nccl.log 2>&1
echo "dp2 nccl rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_bucket_nccl.log | head -1)"
python bench.py --steps 40 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_dp2_bucket_n1.log 2>&1
echo "n1 $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp2_bucket_n1.log | head -1)"
grep -h "phase\|dp_parity" gpurun_out/r2_dp2_bucket.log | tail -20
