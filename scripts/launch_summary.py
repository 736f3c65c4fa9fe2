"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time share of ONE train step."""
import collections, csv, re, sys
path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else -1       # which step (the one ending at the step-th layer-1 forward launch)
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
vals = [float(r["Metric Value"].replace(",", "")) for r in rows]
# a step starts with the layer-1 PointNet forward kernel
starts = [i for i, n in enumerate(names) if "pointnet_l1_fwd" in n]
s, e = starts[step - 1], starts[step]
agg = collections.OrderedDict()
for n, v in zip(names[s:e], vals[s:e]):
    k = re.sub(r"\(.*", "", n)[:100]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
print(f"# one train step: {e - s} launches, {tot / 1e6:.3f} ms (cold-cache, serialised under ncu: compare SHARES)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}%  x{c:3d}  {k}")
