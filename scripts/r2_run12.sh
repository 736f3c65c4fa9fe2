#!/bin/bash
# round-2 GPU session 12 (8 GPUs): final data-parallel record (critic exchange overlapped with the decoder work)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 --phases > gpurun_out/r2_dp8_final.log 2>&1
echo "dp8 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_final.log | head -1)"
timeout 300 $TR --master-port 29712 bench.py --gpus 8 --steps 40 --warmup 5 --no-infer --no-dp-parity > gpurun_out/r2_dp8_final_b.log 2>&1
echo "dp8 (40 steps) rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_final_b.log | head -1)"
python bench.py --steps 40 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_dp8_final_n1.log 2>&1
echo "n1 on this box $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_final_n1.log | head -1)"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR4 --master-port 29713 bench.py --gpus 4 --steps 40 --warmup 5 --no-infer > gpurun_out/r2_dp4_final.log 2>&1
echo "dp4 rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp4_final.log | head -1)"
