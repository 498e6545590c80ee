# scratch: dump both stderr streams of one CLI parity case
import os, sys, subprocess, pathlib, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import test_gpu_cli as t
d = pathlib.Path(tempfile.mkdtemp())
rng = np.random.default_rng(5)
with open(d / "g.chroms", "w") as f:
    for n, l in t.CHROMS: f.write("%s %d\n" % (n, l))
with open(d / "reads.iv", "w") as f:
    for n, l in t.CHROMS:
        m = l * 5 // 100
        s = rng.integers(0, max(1, l - 150), m); ln = rng.integers(50, 151, m)
        for a, b in zip(s, ln): f.write("%s\t%d\t%d\n" % (n, a, min(l, a + b)))
with open(d / "trackB.iv", "w") as f:
    for n, l in t.CHROMS:
        if n == "chrC": continue
        pos = int(rng.integers(0, 100))
        while pos < l:
            e = min(l, pos + int(rng.integers(1, 600)))
            f.write("%s\t%d\t%d\t%s\n" % (n, pos, e, repr(float(rng.integers(1, 4096)) / 1024)))
            pos = e + int(rng.integers(1, 600))
args = t.C + ["--novalue", "--precision=17", "--progress=operations", "=", "smooth", "--window=31",
        "=", "percentile", "90", "--preserve=scratch.pres", "--precision=12", "=", "multiply", "trackB.iv",
        "=", "percentile", "10..90by20", "--window=7", "--min=0.25", "--preserve=scratch2.pres", "=", "clip", "--max=percentile90"]
(a, b) = t.both(d, args)
open("gpurun_out/dbg_ref.err", "wb").write(a[2]); open("gpurun_out/dbg_ours.err", "wb").write(b[2])
print(a[0], b[0], a[1] == b[1])
