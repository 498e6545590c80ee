#!/usr/bin/env python
"""Regenerates the golden fixtures from the UNMODIFIED reference binary.

    make -C oracle ref && python tests/golden/make_golden.py

Inputs (g.chroms, reads.iv, vals.iv, trackB.iv) are generated from a fixed seed;
every case in CASES is run through oracle/_ref/genodsp and its stdout/stderr are
stored as <case>.out / <case>.err.  The reference ships no test vectors of its
own; these files, produced by its own binary, are what pins parity on
machines where /root/reference is absent.
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "genodsp")

CHROMS = [("chrA", 3000), ("chrB", 7000), ("chrC", 1100)]

C = ["--chromosomes=g.chroms"]
CASES = {
    "depth": (C + ["--novalue"], "reads.iv"),
    "depth_show_nocollapse": (C + ["--novalue", "--uncovered:show", "--nocollapse"], "reads.iv"),
    "depth_NA_origin1": (C + ["--novalue", "--uncovered:NA", "--origin=one"], "reads.iv"),
    "cfg1_sum_localmax": (C + ["--novalue", "=", "sum", "--window=101", "=", "localmax", "--neighborhood=11"], "reads.iv"),
    "slidingsum": (C + ["--novalue", "--precision=4", "=", "slidingsum", "--window=100", "--denom=W"], "reads.iv"),
    "smooth101": (C + ["--novalue", "--precision=15", "=", "smooth", "--window=101"], "reads.iv"),
    "bestmax_localmin": (C + ["--novalue", "=", "bestmax", "--window=31", "=", "localmin", "--neighborhood=5", "--infinity=99"], "reads.iv"),
    "cumulative": (C + ["--novalue", "=", "cumulativesum"], "reads.iv"),
    "percentile_binarize": (C + ["--novalue", "=", "sum", "--window=100", "--denom=100", "=", "percentile", "99", "--precision=3",
                                 "=", "binarize", "--threshold=percentile99"], "reads.iv"),
    "five_stage": (C + ["--novalue", "--precision=9", "=", "smooth", "--window=101", "=", "localmax", "--neighborhood=11", "=",
                        "percentile", "99", "--precision=6", "=", "binarize", "--threshold=percentile99"], "reads.iv"),
    "morphology": (C + ["--novalue", "=", "binarize", "6", "=", "open", "31", "=", "close", "201", "=", "clump", "0.5",
                        "--length=300"], "reads.iv"),
    "dilate_erode": (C + ["--novalue", "=", "dilate", "31", "--threshold=8", "=", "erode", "11"], "reads.iv"),
    "anticlump": (C + ["--novalue", "=", "anticlump", "3", "--length=200", "--one=2"], "reads.iv"),
    "valued": (C + ["--precision=2"], "vals.iv"),
    "multi_signal": (C + ["--novalue", "--precision=6", "=", "add", "trackB.iv", "=", "multiply", "trackB.iv", "=", "and",
                          "trackB.iv", "=", "binarize", "0.5"], "reads.iv"),
    "pointwise": (C + ["--novalue", "--precision=3", "=", "addconst", "-4.5", "=", "abs", "=", "clip", "--min=1", "--max=3.5",
                       "=", "invert", "=", "erase", "--min=2", "--max=3"], "reads.iv"),
    # the interval-file operators (add.c, multiply.c, mask.c, logical.c, minmax.c, map.c)
    "subtract_divide": (C + ["--novalue", "--precision=4", "=", "subtract", "trackB.iv", "--value=4", "=", "divide", "trackB.iv",
                             "--value=4", "--infinity=999"], "reads.iv"),
    "or_mask_masknot": (C + ["--novalue", "=", "or", "trackB.iv", "--value=4", "=", "mask", "maskM.iv", "--mask=7", "=", "masknot",
                             "trackB.iv", "--mask=-2"], "reads.iv"),
    "minover": (C + ["--novalue", "=", "minover", "trackB.iv", "--infinity=99"], "reads.iv"),
    "maxover": (C + ["--novalue", "=", "maxover", "trackB.iv", "--zero=-1"], "reads.iv"),
    "minwith_maxwith": (C + ["--novalue", "--precision=3", "=", "minwith", "trackB.iv", "--value=4", "=", "maxwith", "vals.iv", "--value=4"], "reads.iv"),
    "map": (C + ["--novalue", "--precision=4", "=", "map", "depth.map"], "reads.iv"),
    # under --novalue an operator's file is read without values too (every value 1) unless it says --value=
    "add_multiply_values": (C + ["--novalue", "--precision=5", "=", "add", "trackB.iv", "--value=4", "=", "multiply", "trackB.iv",
                                 "--value=4"], "reads.iv"),
    # the state a --window / --min / --max percentile leaves behind (the collect permutation, percentile.c:547-580,
    # then the per-chromosome sorts and bubble passes) is visible to the next operator
    "percentile_collect_window_min": (C + ["--novalue", "=", "percentile", "50", "--window=7", "--min=1", "=", "addconst", "0"], "reads.iv"),
    "percentile_collect_max": (C + ["--novalue", "--precision=3", "=", "slidingsum", "--window=25", "--denom=W", "=", "percentile",
                                    "20..80by30", "--max=6", "--quiet", "=", "addconst", "0.5"], "reads.iv"),
    "percentile_collect_high_rank": (C + ["--novalue", "=", "percentile", "99", "--min=2", "=", "binarize", "--threshold=percentile99"], "reads.iv"),
    "percentile_general_rank": (C + ["--novalue", "=", "percentile", "30", "=", "addconst", "0"], "reads.iv"),
    "subtract_novalue_inherited": (C + ["--novalue", "=", "subtract", "trackB.iv"], "reads.iv"),
    # the bubble passes on real values (every chromosome sorted on its own, then combine_sorted_vectors steps)
    "percentile_general_rank_reals": (C + ["--novalue", "--precision=9", "=", "smooth", "--window=11", "=", "percentile", "30",
                                           "--precision=9", "=", "addconst", "0"], "reads.iv"),
    # read_intervals' per-cell rule (genodsp.c:1307-1330) when the running value comes back to missingVal: it reads as
    # "not yet covered" and the next interval overwrites it -- zero-valued rows under --overlap=min|max, a depth that
    # reaches --missing, a sum that passes through it
    "input_overlap_min_zero_rows": (C + ["=", "input", "quirks.iv", "--overlap=min"], "vals.iv"),
    "input_overlap_max_missing2": (C + ["--uncovered:show", "=", "input", "quirks.iv", "--overlap=max", "--missing=2"], "vals.iv"),
    "input_overlap_min_missing_neg1": (C + ["--uncovered:show", "=", "input", "quirks.iv", "--overlap=minimum", "--missing=-1"], "vals.iv"),
    "input_depth_missing3": (C + ["--uncovered:show", "=", "input", "reads.iv", "--novalue", "--missing=3"], "vals.iv"),
    "input_sum_missing5": (C + ["--uncovered:show", "=", "input", "quirks.iv", "--missing=5"], "vals.iv"),
    # add.c:280-281 adds interval after interval: overlapping integer values on a signal that holds non-integers are
    # ((v+a)+b), not v+(a+b)
    "add_overlapping_ints_on_reals": (C + ["--novalue", "--precision=17", "=", "smooth", "--window=11", "=", "add", "quirks.iv",
                                           "--value=4", "=", "subtract", "quirks.iv", "--value=4"], "reads.iv"),
    "add_overlapping_ints_on_ints": (C + ["--novalue", "=", "add", "quirks.iv", "--value=4"], "reads.iv"),
}


def write_inputs():
    rng = np.random.default_rng(20261018)
    with open(os.path.join(HERE, "g.chroms"), "w") as f:
        for n, l in CHROMS:
            f.write("%s %d\n" % (n, l))
    with open(os.path.join(HERE, "reads.iv"), "w") as f:
        f.write("track name=golden\n# reads\n")
        for n, l in CHROMS:
            m = l * 5 // 100
            s = rng.integers(0, l - 150, m)
            ln = rng.integers(50, 151, m)
            for a, b in zip(s, ln):
                f.write("%s\t%d\t%d\n" % (n, a, a + b))
    with open(os.path.join(HERE, "vals.iv"), "w") as f:
        for n, l in CHROMS:
            for a in np.sort(rng.integers(0, l - 400, l // 300)):
                f.write("%s %d %d %s\n" % (n, a, a + int(rng.integers(1, 400)), repr(float(rng.integers(-8, 9)) / 4)))
    with open(os.path.join(HERE, "trackB.iv"), "w") as f:
        for n, l in CHROMS[:2]:
            pos = int(rng.integers(0, 100))
            while pos < l:
                e = min(l, pos + int(rng.integers(1, 600)))
                f.write("%s\t%d\t%d\t%s\n" % (n, pos, e, repr(float(rng.integers(1, 4096)) / 1024)))
                pos = e + int(rng.integers(1, 600))
    # (generated after the files above so that those keep their bytes)
    with open(os.path.join(HERE, "maskM.iv"), "w") as f:           # unsorted, overlapping, no value column
        for _ in range(60):
            n, l = CHROMS[int(rng.integers(0, 3))]
            a = int(rng.integers(0, l - 300))
            f.write("%s\t%d\t%d\n" % (n, a, a + int(rng.integers(1, 300))))
    with open(os.path.join(HERE, "depth.map"), "w") as f:          # piecewise-linear map of the depth values
        f.write("# depth -> score\n")
        for x in rng.permutation(np.arange(0, 26, 2)):
            f.write("%r %r\n" % (float(x), float(rng.integers(-40, 41)) / 8))
    with open(os.path.join(HERE, "quirks.iv"), "w") as f:          # unsorted, deeply overlapping, small integer values
        for _ in range(400):                                       # (many zeros, twos, fives and minus ones)
            n, l = CHROMS[int(rng.integers(0, 3))]
            a = int(rng.integers(0, l - 200))
            f.write("%s\t%d\t%d\t%d\n" % (n, a, a + int(rng.integers(1, 200)), int(rng.choice([0, 0, 2, 5, -1, 3, 7, -4]))))


def main():
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    write_inputs()
    for name, (args, stdin) in CASES.items():
        with open(os.path.join(HERE, stdin), "rb") as fin:
            p = subprocess.run([REF] + args, stdin=fin, capture_output=True, cwd=HERE)
        assert p.returncode == 0, (name, p.stderr)
        open(os.path.join(HERE, name + ".out"), "wb").write(p.stdout)
        open(os.path.join(HERE, name + ".err"), "wb").write(p.stderr)
    json.dump({k: {"args": v[0], "stdin": v[1]} for k, v in CASES.items()}, open(os.path.join(HERE, "cases.json"), "w"), indent=1)
    print("wrote %d golden cases" % len(CASES))


if __name__ == "__main__":
    main()
