#!/usr/bin/env python
"""Instruction census of every kernel in the built library (cuobjdump -sass): global load/store widths,
FP64 multiplies / adds / fused multiply-adds, shuffles, barriers, atomics.  Backs the claims in DESIGN.md
(256-bit accesses in the streaming kernels, no DFMA in the exact-order FIR, no barrier in the chain-free
passes).  Usage: scripts/sass_census.py [library.so] > profiles/r1_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["instr", "LDG.256", "LDG.128", "LDG.other", "STG.256", "STG.128", "STG.other", "DFMA", "DMUL", "DADD", "SHFL", "BAR", "ATOM/RED", "LDS", "STS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "genodsp_b200", "lib", "libgdsp_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    stats, fn, k = collections.OrderedDict(), None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = re.sub(r"\(.*", "", names[k]).replace("void ", ""); k += 1
            stats[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if not m or fn is None:
            continue
        op, c = m.group(1), stats[fn]
        c["instr"] += 1
        if op.startswith("LDG"):
            c["LDG.256" if ".256" in op else "LDG.128" if ".128" in op else "LDG.other"] += 1
        elif op.startswith("STG"):
            c["STG.256" if ".256" in op else "STG.128" if ".128" in op else "STG.other"] += 1
        elif op.startswith(("ATOM", "RED")):
            c["ATOM/RED"] += 1
        else:
            for key in ("DFMA", "DMUL", "DADD", "SHFL", "BAR", "LDS", "STS"):
                if op.startswith(key):
                    c[key] += 1
    print("%-46s" % "kernel" + "".join("%10s" % c for c in COLS))
    for fn in sorted(stats):
        print("%-46s" % fn[:45] + "".join("%10d" % stats[fn][c] for c in COLS))


if __name__ == "__main__":
    main()
