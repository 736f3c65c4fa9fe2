#!/bin/bash
# round-2 GPU session 10 (1 GPU): layer-1 kernels after the block-order change: kernel tests, launch lists (train + inference), benches
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py tests/test_gpu_inference.py -m gpu -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2_t10.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_launches_B256_v4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-infer --graph off > gpurun_out/r2_ncu_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/r2_launches_B256_v4.csv > gpurun_out/r2_launch_shares_step_B256_v4.txt 2>&1
grep -i "l1\|im2col\|tcn_bn\|launches" gpurun_out/r2_launch_shares_step_B256_v4.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_infer_v2.csv \
    python bench.py --workload infer --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_infer.log 2>&1
grep -c . gpurun_out/r2_launches_infer_v2.csv
python - <<'PY'
import collections, csv, re
rows = [l for l in open("gpurun_out/r2_launches_infer_v2.csv") if not l.startswith("==")]
r = list(csv.DictReader(rows))
names = [x["Kernel Name"] for x in r]; vals = [float(x["Metric Value"].replace(",", "")) for x in r]
starts = [i for i, n in enumerate(names) if "pointnet_l1_fwd" in n]
s, e = starts[3], starts[4]
agg = collections.OrderedDict()
for n, v in zip(names[s:e], vals[s:e]):
    k = re.sub(r"\(.*", "", n)[:90]; a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
out = [f"# one inference step (1020 crops): {e - s} launches, {tot / 1e6:.3f} ms (cold-cache, serialised under ncu: compare SHARES)"]
out += [f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}%  x{c:3d}  {k}" for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
open("gpurun_out/r2_launch_shares_infer_v2.txt", "w").write("\n".join(out) + "\n"); print("\n".join(out[:9]))
PY
python bench.py --workload infer --steps 100 --warmup 5 --no-cpu > gpurun_out/r2_infer10.log 2>&1; echo "infer $(grep -o '"value": [0-9.]*' gpurun_out/r2_infer10.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_infer10.log | head -1)"
python bench.py --steps 20 --warmup 5 --no-cpu --no-infer > gpurun_out/r2_bench10.log 2>&1; echo "train $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench10.log | head -1)"
