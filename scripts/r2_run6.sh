#!/bin/bash
# round-2 GPU session 6 (8 GPUs): data-parallel bench with parity through the graph-replayed path, per-phase timings, inference
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 --phases > gpurun_out/r2_dp8_split.log 2>&1
echo "dp8 split rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_split.log | head -1)"
PCAA_DP_ONE_GRAPH=1 timeout 300 $TR --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 --no-infer > gpurun_out/r2_dp8_onegraph.log 2>&1
echo "dp8 onegraph rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_onegraph.log | head -1)"
PCAA_DP_ONE_GRAPH=1 timeout 300 $TR --master-port 29613 bench.py --gpus 8 --steps 40 --warmup 5 --no-infer --no-dp-parity > gpurun_out/r2_dp8_onegraph_b.log 2>&1
echo "dp8 onegraph (40 steps) rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_onegraph_b.log | head -1)"
timeout 300 $TR --master-port 29614 bench.py --gpus 8 --steps 40 --warmup 5 --no-infer --no-dp-parity > gpurun_out/r2_dp8_split_b.log 2>&1
echo "dp8 split (40 steps) rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_split_b.log | head -1)"
python bench.py --steps 40 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_dp8_n1.log 2>&1
echo "n1 on this box $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_n1.log | head -1)"
tail -c 400 gpurun_out/r2_dp8_onegraph.log
