#!/usr/bin/env python
"""Per-kernel SASS evidence that the hot path is Blackwell-native: counts of the tcgen05 / TMA / TMEM instructions in every
kernel of libraft_b200.so (cuobjdump -sass).  UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = TMA tensor load, UTMASTG = TMA tensor
store, LDTM / STTM = tcgen05.ld / tcgen05.st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, ACQBULK/UBLKCP = bulk copies.

    python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.txt        (runs without a GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raft-tf_b200", "lib", "libraft_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "LDG", "STG", "LDS", "STS", "MUFU", "FFMA", "HMMA"]


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kern, counts, total = None, {}, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = collections.Counter()
            total[kern] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            total[kern] += 1
            base = op.split(".")[0]
            counts[kern][base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                counts[kern]["UTCHMMA.2CTA"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static SASS, sm_100a)")
    print("# " + " ".join(f"{o:>8}" for o in ["insts"] + OPS) + "  kernel")
    agg = collections.Counter()
    for k in sorted(counts, key=lambda k: -counts[k]["UTCHMMA"] * 10000 - counts[k]["UTMALDG"] * 100 - total[k] / 1e4):
        c = counts[k]
        agg.update(c)
        print("  " + " ".join(f"{v:>8}" for v in [total[k]] + [c[o] for o in OPS]) + "  " + demangle(k)[:110])
    print("# total: " + ", ".join(f"{o}={agg[o]}" for o in OPS[:8]))


if __name__ == "__main__":
    main()
