"""Turn an `ncu --set full --page raw --csv` export of ONE train step (scripts/ncu_full.sh) into the per-kernel roofline table of
DESIGN.md section 4 and into profiles/traffic.json (the `roofline.traffic` figure of bench.py), so that neither is hand-copied.

    python scripts/roofline_table.py gpurun_out/full_<tag>_raw.csv --batch 256 --nmax 150 --tag r2v1 \
        [--write-traffic] > profiles/r2_roofline_table.md

For every hot launch: measured time, measured DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the ALGORITHMIC bytes /
FLOPs of that launch (formulas below), achieved GB/s or TFLOP/s on the algorithmic work and the fraction of the
measured peak (MEASURED_PEAKS.json: HBM copy bandwidth, sustained dense bf16).  ncu times are cold-cache and serialised: the
fractions here are per-kernel ceilings-in-isolation; bench.py's `roofline` is the in-step number.
"""
import argparse
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POINTNET = [(4, 512), (512, 512), (512, 1024), (1024, 1024)]          # (Cin, Cout) of layers 1..4, models.py:86-98


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "MEASURED_PEAKS.json"
    return 6650.0, 1400.0, "B200_PROFILING.md fallback"


def work(name, occ, B, N):
    """(bound, algorithmic bytes, algorithmic flops, what) of the occ-th launch (0-based) of kernel `name` in one step."""
    P = B * 30 * N
    S = 120 * N
    dec = [64, S // 16, S // 8, S // 4, S // 2, S]
    m = re.search(r"gemm_tc_kernel<256, (\w+), (\w+), (\d+)>", name)
    if m:
        a_mn, b_mn, mode = m.group(1) in ("1", "true"), m.group(2) in ("1", "true"), int(m.group(3))
        if mode == 7 and occ < 3:                               # PointNet forward, layers 2..4
            ci, co = POINTNET[occ + 1]
            return "tensor", 2 * P * (ci + co) + 2 * ci * co, 2.0 * P * ci * co, f"PointNet L{occ + 2} forward (+ BN statistics)"
        if mode == 9 and occ < 3:                               # data gradient towards layers 3, 2, 1
            ci, co = POINTNET[3 - occ]
            return "tensor", 2 * P * (co + 2 * ci) + 2 * ci * co, 2.0 * P * ci * co, f"PointNet L{4 - occ} data gradient (+ ELU', BN-backward sums)"
        if mode == 4 and not a_mn and not b_mn and occ < 3:     # weight gradient, k = points
            ci, co = POINTNET[3 - occ]
            return "tensor", 2 * P * (co + ci) + 4 * ci * co, 2.0 * P * ci * co, f"PointNet L{4 - occ} weight gradient"
        if mode == 4 and a_mn and b_mn:                         # TCN weight gradients (6) interleaved with decoder's (5): by size
            return None
        if mode in (1, 2, 5) or mode == 0:
            return None
        return None
    if "bn_elu_apply_rows" in name and occ < 2:
        c = POINTNET[occ + 1][1]
        return "hbm", 4 * P * c, 0.0, f"BatchNorm + ELU apply, layer {occ + 2} (read y, write a)"
    if "bn_bwd_apply_t" in name and occ < 2:
        c = POINTNET[2 - occ][1]
        return "hbm", 6 * P * c, 0.0, f"BatchNorm backward apply, layer {3 - occ} (read dz, y; write dy)"
    if "pool_bwd_apply" in name:
        return "hbm", 4 * P * 1024, 0.0, "mean-pool' o ELU' o BN' of layer 4 (read y4, write dy4)"
    if "meanpool_staged" in name:
        return "hbm", 2 * P * 1024, 0.0, "BN + ELU + mean pool over points, layer 4 (read y4)"
    if "pointnet_l1_fwd" in name:
        return "hbm", 16 * P + 4 * P * 512, 0.0, "layer 1 forward: read x, write y1 and a1"
    if "pointnet_l1_wgrad_t" in name:
        return "hbm", 16 * P + 4 * P * 512, 0.0, "layer 1 weight gradient (+ BN backward): read x, dz1, y1"
    if "adam_flat" in name:
        n_dec = sum((dec[i] + 7) // 8 * 8 * dec[i + 1] + dec[i + 1] for i in range(5)) + 32 * 64 + 64
        n = [4481, n_dec, 2439236][occ] if occ < 3 else 0
        return "hbm", 30 * n, 0.0, ["Adam, critic", "Adam, decoder span (+ bf16 shadow)", "Adam, encoder span"][occ] if occ < 3 else "Adam"
    if "chamfer_fwd4" in name:
        return "fp32", 2 * 16 * P, 2.0 * B * 30 * N * N * 2 * 4, "Chamfer forward: both nearest-neighbour searches"
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--nmax", type=int, default=150)
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--write-traffic", action="store_true")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, data = rows[0], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    hbm, tf, src = peaks()
    occ = collections.Counter()
    out, gemm_traffic, gemm_alg, gemm_ns = [], [], [], 0.0
    total_dram = 0.0
    for d in data:
        name = d[ix["Kernel Name"]]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("pcaa::", "")
        t_ns = float(d[ix["gpu__time_duration.sum"]].replace(",", ""))
        dram = float(d[ix["dram__bytes_read.sum"]].replace(",", "")) + float(d[ix["dram__bytes_write.sum"]].replace(",", ""))
        unit_r = rows[1][ix["dram__bytes_read.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit_r, 1.0)
        dram *= scale
        tu = rows[1][ix["gpu__time_duration.sum"]]
        t_ns *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(tu, 1.0)
        total_dram += dram
        w = work(short, occ[short], args.batch, args.nmax)
        occ[short] += 1
        if w is None:
            continue
        bound, ab, fl, what = w
        if bound == "tensor":
            ach, frac, unit = fl / t_ns / 1e3, fl / t_ns / 1e3 / tf, "TFLOP/s"
            gemm_traffic.append(dram), gemm_alg.append(ab)
            gemm_ns += t_ns
        elif bound == "hbm":
            ach, frac, unit = ab / t_ns, ab / t_ns / hbm, "GB/s"
        else:
            ach, frac, unit = fl / t_ns / 1e3, float("nan"), "TFLOP/s fp32"
        out.append((short, what, bound, t_ns / 1e6, ab / 1e9, dram / 1e9, ach, unit, frac))
    print(f"# per-kernel roofline, one train step at B={args.batch}, N={args.nmax} ({args.tag}); peaks from {src}: HBM {hbm:.0f} GB/s, bf16 {tf:.0f} TFLOP/s sustained")
    print("# generated by scripts/roofline_table.py from the ncu --set full raw export (cold-cache, serialised launches)\n")
    print("| kernel | what | bound | time ms | algorithmic GB | measured DRAM GB | achieved | fraction of peak |")
    print("|---|---|---|---|---|---|---|---|")
    for short, what, bound, ms, ab, dram, ach, unit, frac in out:
        print(f"| `{short[:60]}` | {what} | {bound} | {ms:.3f} | {ab:.3f} | {dram:.3f} | {ach:.0f} {unit} | {frac:.2f} |")
    print(f"\nDRAM traffic of all captured launches: {total_dram / 1e9:.1f} GB")
    if gemm_traffic:
        rec = {"bytes_per_launch": sum(gemm_traffic) / len(gemm_traffic), "launches": len(gemm_traffic),
               "algorithmic_bytes_per_launch": sum(gemm_alg) / len(gemm_alg), "ncu_time_ms_sum": gemm_ns / 1e6,
               "_source": f"scripts/roofline_table.py on {os.path.basename(args.csv)} ({args.tag})"}
        print(f"PointNet tcgen05 GEMMs: {len(gemm_traffic)} launches, measured {rec['bytes_per_launch'] / 1e9:.3f} GB/launch vs algorithmic "
              f"{rec['algorithmic_bytes_per_launch'] / 1e9:.3f} GB/launch, {gemm_ns / 1e6:.3f} ms in total")
        if args.write_traffic:
            p = os.path.join(ROOT, "profiles", "traffic.json")
            d = json.load(open(p)) if os.path.exists(p) else {}
            d[f"train_B{args.batch}_N{args.nmax}"] = rec
            json.dump(d, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
