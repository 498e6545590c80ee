#!/usr/bin/env python
"""Wall-clock of the drop-in CLI (genodsp_b200/bin/genodsp) against the unmodified reference binary
(oracle/_ref/genodsp, when built) on the same text input: BASELINE cfg1 (one 10 Mbp chromosome,
1 M reads, `sum --window=101 = localmax --neighborhood=11`) and the hg38/--scale depth pipeline
(`smooth --window=101`, collapsed text output).  Outputs are compared byte for byte."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

OURS = os.path.join(ROOT, "genodsp_b200", "bin", "genodsp")
REF = os.path.join(ROOT, "oracle", "_ref", "genodsp")


def write_case(d, chroms, seed):
    rng = np.random.default_rng(seed)
    with open(os.path.join(d, "g.chroms"), "w") as f:
        for n, l in chroms:
            f.write("%s %d\n" % (n, l))
    path = os.path.join(d, "reads.iv")
    with open(path, "w") as f:
        for n, l in chroms:
            m = int(round(l * 5 / 100.0))
            s = np.sort(rng.integers(0, max(1, l - 150), m))
            e = np.minimum(l, s + rng.integers(50, 151, m))
            f.write("".join("%s\t%d\t%d\n" % (n, a, b) for a, b in zip(s.tolist(), e.tolist())))
    return path


def run(binary, args, stdin_path, cwd, tag):
    """stdout goes to a file in `cwd` (a pipe into this Python process would be the bottleneck for GB-sized
    outputs); it is hashed afterwards, outside the timed region.  GENODSP_TIMING makes our CLI print its phases."""
    out_path = os.path.join(cwd, "out.%s.txt" % tag)
    env = dict(os.environ, GENODSP_TIMING="1")
    t0 = time.perf_counter()
    with open(stdin_path, "rb") as fin, open(out_path, "wb") as fout:
        p = subprocess.run([binary] + args, stdin=fin, stdout=fout, stderr=subprocess.PIPE, cwd=cwd, env=env)
    dt = time.perf_counter() - t0
    h = hashlib.md5()
    nbytes = 0
    with open(out_path, "rb") as f:
        for block in iter(lambda: f.read(1 << 24), b""):
            h.update(block); nbytes += len(block)
    os.remove(out_path)
    phases = {}
    for line in p.stderr.decode(errors="replace").splitlines():
        if line.startswith("[timing]"):
            f = line.split()
            phases[" ".join(f[1:-2])] = float(f[-2])
    return dt, p.returncode, h.hexdigest(), nbytes, phases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=16)
    args = ap.parse_args()
    out = {}
    cases = [("cfg1", [("chr1", 10000000)], ["--chromosomes=g.chroms", "--novalue", "=", "sum", "--window=101", "=", "localmax", "--neighborhood=11"]),
             ("hg38_div%d_smooth" % args.scale, bench.scaled_genome(args.scale), ["--chromosomes=g.chroms", "--novalue", "--precision=3", "=", "smooth", "--window=101"])]
    for name, chroms, cmd in cases:
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
            reads = write_case(d, chroms, 1)
            bases = sum(l for _, l in chroms)
            row = {"bases": bases, "input_bytes": os.path.getsize(reads)}
            run(OURS, cmd, reads, d, "warm")                          # warm-up: page cache
            t, rc, md5, nbytes, phases = run(OURS, cmd, reads, d, "ours")
            row["ours"] = {"s": round(t, 3), "rc": rc, "out_bytes": nbytes, "gbp_s": round(bases / t / 1e9, 4), "phases_s": phases}
            if os.path.exists(REF):
                tr, rcr, md5r, nbr, _ = run(REF, cmd, reads, d, "ref")
                row["reference"] = {"s": round(tr, 3), "rc": rcr, "out_bytes": nbr, "gbp_s": round(bases / tr / 1e9, 4)}
                row["identical_output"] = (md5 == md5r)
                row["speedup"] = round(tr / t, 2)
            out[name] = row
            print(name, row, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
