#!/bin/bash
# round-2 GPU session 8 (1 GPU): whole GPU test tier after the phase restructuring / eval layer-1 kernel, inference launch list, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_t8.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_t8.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_infer_v1.csv \
    python bench.py --workload infer --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_ncu_infer.log 2>&1
python - <<'PY'
import collections, csv, re
rows = [l for l in open("gpurun_out/r2_launches_infer_v1.csv") if not l.startswith("==")]
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
open("gpurun_out/r2_launch_shares_infer_v1.txt", "w").write("\n".join(out) + "\n"); print("\n".join(out[:16]))
PY
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench8.log 2>&1; echo "bench rc=$? $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench8.log | head -2 | tr '\n' ' ')"
