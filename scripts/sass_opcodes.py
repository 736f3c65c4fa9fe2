"""Blackwell evidence: histogram of the tcgen05 / TMEM / TMA SASS opcodes in the built library, per kernel.

    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt

Runs `cuobjdump -sass` on opensetgaitrecognition_pcaa_b200/libpcaa_sm100.so (no GPU needed) and counts, for every kernel,
the mnemonics B200_PROFILING.md names as proof of the sm_100a programming model: UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit),
LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG / UTMAPF (TMA load / store / reduce / prefetch), SYNCS (mbarrier),
plus the packed fp32x2 arithmetic (FFMA2 / FADD2) and MUFU.EX2 of the elementwise kernels.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opensetgaitrecognition_pcaa_b200", "libpcaa_sm100.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTMACCTL", "SYNCS", "FFMA2", "FADD2", "MUFU.EX2",
       "HMMA", "ATOMG", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for key in OPS:
                if op == key or op.startswith(key + ".") or (key == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    cur[key] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: architectures {arch}, {len(per)} kernels")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("# library totals: " + ", ".join(f"{k} {tot[k]}" for k in OPS if tot[k]))
    print(f"{'kernel':110s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k in OPS))
    for (name, c), dn in zip(per.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("pcaa::", "")
        if not any(c[k] for k in OPS):
            continue
        print(f"{short[:110]:110s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in OPS))


if __name__ == "__main__":
    sys.exit(main())
