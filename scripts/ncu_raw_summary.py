"""Condense an `ncu --page raw --csv` export into one line per kernel launch (time, DRAM traffic, pipe utilisation)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16_mma%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma_act%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("smsp__cycles_active.avg", "cycles")]
cols = [(c, n) for c, n in COLS if c in idx]
print("# " + " | ".join(["kernel", "grid"] + [f"{n}[{units[idx[c]]}]" for c, n in cols]))
for d in data:
    name = d[idx["Kernel Name"]]
    name = name[:name.index("(")] if "(" in name else name
    print(" | ".join([name.replace("void ", "").replace("pcaa::", ""), d[idx["Grid Size"]]] + [d[idx[c]] for c, _ in cols]))
