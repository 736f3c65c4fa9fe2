"""Measurement sweeps of SURVEY.md 8(d) on one B200 (run under gpurun): batch sweep of the train step (config 2), points-per-frame
sweep (config 4), open-set inference throughput (config 5, single GPU leg).  Writes gpurun_out/sweep_<tag>.jsonl (one bench.py
JSON line per configuration) and a table."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["batch", "nmax", "infer"]
runs = []
if "batch" in which:
    runs += [("train", ["--batch", str(b), "--steps", str(max(5, min(40, 2560 // b)))]) for b in (32, 64, 128, 256, 512, 1024)]
if "nmax" in which:
    runs += [("train", ["--batch", "256", "--nmax", str(n), "--steps", "10"]) for n in (50, 70, 90, 110, 130)]
if "infer" in which:
    runs += [("infer", ["--workload", "infer", "--batch", str(b), "--steps", "10"]) for b in (6, 96, 1020, 4092)]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.jsonl")
lines = []
with open(out, "w") as f:
    for kind, extra in runs:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--warmup", "3"] + extra, capture_output=True, text=True)
        js = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not js:
            print("FAILED", extra, r.stderr[-800:])
            continue
        d = json.loads(js[-1])
        lines.append(d)
        f.write(json.dumps(d) + "\n")
        f.flush()
print(f"{'workload':48s} {'B/gpu':>6s} {'value':>10s} {'unit':>9s} {'ms/step':>8s} {'e2e':>10s} {'tc TF/s':>8s} {'frac':>5s} {'share':>6s} {'step frac':>9s}")
for d in lines:
    r = d["roofline"]
    print(f"{d['config']['workload']:48s} {d['config']['batch_per_gpu']:6d} {d['value']:10.1f} {d['unit']:>9s} {d['ms_per_step']:8.3f} "
          f"{d['e2e']['value']:10.1f} {r['achieved']:8.1f} {r['frac']:5.2f} {r['share_of_step']:6.2f} {r['whole_step_tensor_frac']:9.3f}")
