#!/bin/bash
# round-2 GPU session 17 (8 GPUs): A/B of the sharded decoder update (PCAA_DP_SHARD_ADAM=1, default) vs the plain all-reduce + full Adam (=0)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
i=0
for v in 1 0 1 0; do
  i=$((i+1))
  extra="--no-dp-parity"; [ $i -eq 1 ] && extra="--phases"; [ $i -eq 2 ] && extra="--phases --no-dp-parity"
  PCAA_DP_SHARD_ADAM=$v timeout 400 $TR --master-port $((29940+i)) bench.py --gpus 8 --steps 40 --warmup 5 --no-infer $extra > gpurun_out/r2_dp8_shard_ab_${i}_shard$v.log 2>&1
  echo "run $i shard=$v rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_dp8_shard_ab_${i}_shard$v.log | head -1) $(grep -o '"ok": [a-z]*' gpurun_out/r2_dp8_shard_ab_${i}_shard$v.log | head -1)"
done
