"""Drop-in check of the host CLI: genodsp_b200/bin/genodsp (C host + CUDA library) must write
byte-identical stdout (and the same variable / threshold messages on stderr) as the unmodified
reference binary oracle/_ref/genodsp for the same command line and input."""
import json
import os
import subprocess

import numpy as np
import pytest

from checkers import REF_BIN, ROOT, have_ref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="oracle/_ref/genodsp not built")]

OURS = os.path.join(ROOT, "genodsp_b200", "bin", "genodsp")
CHROMS = [("chrA", 300000), ("chrB", 150000), ("chrC", 70000), ("chrD", 1234)]


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    rng = np.random.default_rng(5)
    with open(d / "g.chroms", "w") as f:
        for n, l in CHROMS:
            f.write("%s %d\n" % (n, l))
    # coverage-like reads, mean depth ~5, a few unknown chromosomes, comments and a track line
    with open(d / "reads.iv", "w") as f:
        f.write("track name=reads\n# a comment\n\n")
        for n, l in CHROMS:
            m = l * 5 // 100
            s = rng.integers(0, max(1, l - 150), m)
            ln = rng.integers(50, 151, m)
            for a, b in zip(s, ln):
                f.write("%s\t%d\t%d\n" % (n, a, min(l, a + b)))
        f.write("chrUn\t5\t10\n")
    # valued intervals (dyadic values), column 4 and a 5th column
    with open(d / "vals.iv", "w") as f:
        for n, l in CHROMS:
            m = l // 300
            s = np.sort(rng.integers(0, l - 400, m))
            for a in s:
                f.write("%s %d %d %s %s\n" % (n, a, a + int(rng.integers(1, 400)), repr(float(rng.integers(-8, 9)) / 4),
                                              repr(float(rng.integers(0, 5)))))
    # sorted disjoint second track (~50 % covered), values k/1024; chrC absent
    with open(d / "trackB.iv", "w") as f:
        for n, l in CHROMS:
            if n == "chrC":
                continue
            pos = int(rng.integers(0, 100))
            while pos < l:
                e = min(l, pos + int(rng.integers(1, 600)))
                f.write("%s\t%d\t%d\t%s\n" % (n, pos, e, repr(float(rng.integers(1, 4096)) / 1024)))
                pos = e + int(rng.integers(1, 600))
    # unsorted, overlapping mask intervals
    with open(d / "maskM.iv", "w") as f:
        for _ in range(400):
            n, l = CHROMS[int(rng.integers(0, 3))]
            a = int(rng.integers(0, l - 2000))
            f.write("%s\t%d\t%d\n" % (n, a, a + int(rng.integers(1, 2000))))
    return d


def run(binary, args, stdin_path, cwd):
    with open(stdin_path, "rb") as fin:
        p = subprocess.run([binary] + args, stdin=fin, capture_output=True, cwd=cwd, timeout=600)
    return p.returncode, p.stdout, p.stderr


def both(data, args, stdin="reads.iv"):
    a = run(REF_BIN, args, data / stdin, data)
    b = run(OURS, args, data / stdin, data)
    return a, b


def assert_same(data, args, stdin="reads.iv", stderr_too=True):
    (rc_r, out_r, err_r), (rc_o, out_o, err_o) = both(data, args, stdin)
    assert rc_r == 0, err_r.decode()[-500:]
    assert rc_o == 0, err_o.decode()[-2000:]
    if out_r != out_o:
        lr, lo = out_r.split(b"\n"), out_o.split(b"\n")
        for i, (x, y) in enumerate(zip(lr, lo)):
            assert x == y, "stdout differs at line %d: ref %r ours %r (args %s)" % (i + 1, x, y, args)
        assert len(lr) == len(lo), "stdout has %d lines, reference %d (args %s)" % (len(lo), len(lr), args)
    if stderr_too:
        assert err_r == err_o, "stderr differs:\nref:  %r\nours: %r" % (err_r[-400:], err_o[-400:])


C = ["--chromosomes=g.chroms"]

GOLDEN = os.path.join(ROOT, "tests", "golden")
with open(os.path.join(GOLDEN, "cases.json")) as _f:
    GOLDEN_CASES = json.load(_f)


@pytest.mark.parametrize("case", sorted(GOLDEN_CASES))
def test_cli_matches_committed_golden(case):
    """our CLI against the fixtures the reference binary produced (tests/golden/make_golden.py): the same
    check as the live comparisons below, but pinned to committed bytes"""
    spec = GOLDEN_CASES[case]
    rc, out, err = run(OURS, spec["args"], os.path.join(GOLDEN, spec["stdin"]), GOLDEN)
    assert rc == 0, err.decode()[-2000:]
    want = open(os.path.join(GOLDEN, case + ".out"), "rb").read()
    if out != want:
        lo, lw = out.split(b"\n"), want.split(b"\n")
        for i, (x, y) in enumerate(zip(lo, lw)):
            assert x == y, "%s: stdout differs at line %d: ours %r golden %r" % (case, i + 1, x, y)
        assert len(lo) == len(lw), "%s: %d lines, golden %d" % (case, len(lo), len(lw))
    assert err == open(os.path.join(GOLDEN, case + ".err"), "rb").read(), case



def test_cfg1_depth_sum_localmax(data):
    assert_same(data, C + ["--novalue", "=", "sum", "--window=101", "=", "localmax", "--neighborhood=11"])


def test_cfg2_depth_smooth(data):
    assert_same(data, C + ["--novalue", "--precision=12", "=", "smooth", "--window=101"])


def test_cfg3_percentile_pipeline(data):
    assert_same(data, C + ["--novalue", "=", "sum", "--window=100", "--denom=100", "=", "percentile", "99",
                           "--precision=3", "=", "binarize", "--threshold=percentile99"])


def test_five_stage_target_pipeline(data):
    assert_same(data, C + ["--novalue", "--precision=9", "=", "smooth", "--window=101", "=", "localmax",
                           "--neighborhood=11", "=", "percentile", "99", "--precision=6", "=", "binarize",
                           "--threshold=percentile99"])


def test_cfg4_morphology_clump(data):
    assert_same(data, C + ["--novalue", "=", "binarize", "6", "=", "open", "101", "=", "close", "1001", "=", "clump",
                           "0.5", "--length=1000"])
    assert_same(data, C + ["--novalue", "=", "dilate", "31", "--threshold=8", "=", "erode", "11"])
    assert_same(data, C + ["--novalue", "=", "anticlump", "3", "--length=200", "--one=2"])


def test_cfg5_multi_signal_fused_chain(data):
    assert_same(data, C + ["--novalue", "--precision=6", "=", "add", "trackB.iv", "=", "multiply", "trackB.iv", "=", "mask",
                           "maskM.iv", "--mask=0", "=", "and", "trackB.iv", "=", "binarize", "0.5"])
    assert_same(data, C + ["--novalue", "--precision=6", "=", "subtract", "trackB.iv", "=", "divide", "trackB.iv",
                           "--infinity=1000", "=", "masknot", "trackB.iv", "--mask=-1", "=", "abs", "=", "addconst", "0.25"])
    assert_same(data, C + ["--novalue", "=", "erase", "--min=3", "--max=5", "=", "or", "maskM.iv", "--novalue"])


def test_pointwise_and_window_ops(data):
    assert_same(data, C + ["--novalue", "--precision=4", "=", "slidingsum", "--window=100", "--denom=W", "=", "clip",
                           "--min=2", "--max=6.5", "=", "invert", "=", "bestmax", "--window=31", "=", "localmin",
                           "--neighborhood=5", "--infinity=99"])
    assert_same(data, C + ["--novalue", "=", "cumulativesum", "=", "bestmin", "W=4", "=", "invert", "one"])
    assert_same(data, C + ["--window=50", "--novalue", "=", "sum", "--denom=actual", "--zero=-1", "=", "erase", "--max=2",
                           "--keep:inside"])
    assert_same(data, C + ["--novalue", "=", "sum", "--window=chromosome"])


def test_valued_input_and_output_options(data):
    assert_same(data, C + ["--precision=2"], stdin="vals.iv")
    assert_same(data, C + ["--value=5", "--precision=1", "--uncovered:show"], stdin="vals.iv")
    assert_same(data, C + ["--novalue", "--uncovered:NA", "--origin=one"])
    assert_same(data, C + ["--novalue", "--nocollapse", "--nooutputvalue", "=", "binarize", "9"])
    assert_same(data, ["chrA:300000", "chrB:1000:50000", "--novalue", "--cliptochromosome"])


def test_input_output_variables_operators(data):
    args = C + ["=", "input", "vals.iv", "--missing=-3", "=", "output", "mid.out", "--precision=3", "--uncovered:show", "=",
                "percentile", "10..90by20", "--quiet", "--preserve=scratch.tmp", "=", "variables", "=", "input", "reads.iv",
                "--novalue", "--overlap=max", "=", "percentile", "0..100"]
    (rc_r, out_r, err_r), _ = both(data, args)
    mid_ref = open(data / "mid.out", "rb").read()
    (rc_o, out_o, err_o) = run(OURS, args, data / "reads.iv", data)
    mid_ours = open(data / "mid.out", "rb").read()
    assert rc_r == 0 and rc_o == 0, (err_r[-300:], err_o[-1000:])
    assert out_r == out_o
    assert mid_ref == mid_ours
    assert err_r == err_o
    assert_same(data, C + ["=", "input", "vals.iv", "--overlap=min", "--missing=2"])


def test_progress_protocol(data):
    assert_same(data, C + ["--novalue", "--progress=operations", "=", "addconst", "1", "=", "abs", "=", "percentile", "50",
                           "--quiet", "=", "binarize", "--threshold=percentile50"])


def test_percentile_then_binarize_variants(data):
    """the executor hands the sorted post-percentile state to `binarize` without sorting it
    (gd_ops_percentile.c / gdsp_sorted_binarize); every variant must still print the reference's bytes"""
    # ties above, custom one/zero, a pointwise operator behind the binarize in the same fused chain
    assert_same(data, C + ["--novalue", "=", "percentile", "90", "=", "binarize", "--threshold=percentile90", "--ties:above",
                           "--one=3", "--zero=-1", "=", "addconst", "2"])
    # a literal threshold that is not the percentile, and a windowed operator behind it
    assert_same(data, C + ["--novalue", "=", "percentile", "95", "--quiet", "=", "binarize", "4", "=", "dilate", "5"])
    # last rank in the FIRST chromosome: the reference's bubble passes, not a global sort -- the hand-over must not fire
    assert_same(data, C + ["--novalue", "=", "percentile", "40", "=", "binarize", "--threshold=percentile40"])
    # several percentiles, the threshold is the lowest one; and the same under the operations trace
    assert_same(data, C + ["--novalue", "=", "percentile", "90..99", "=", "binarize", "--threshold=percentile90"])
    assert_same(data, C + ["--novalue", "--progress=operations", "=", "percentile", "97", "=", "binarize", "--threshold=percentile97"])


def test_percentile_leaves_the_reference_shuffle(data):
    """the state a percentile leaves in the vectors is visible to the next operator -- with
    --window / --min / --max that is the collect permutation (percentile.c:547-580) followed by the sorts and
    bubble passes over the front part; `addconst 0` is a no-op that makes the output show it"""
    assert_same(data, C + ["--novalue", "=", "percentile", "50", "--window=7", "--min=1", "=", "addconst", "0"])
    assert_same(data, C + ["--novalue", "--precision=3", "=", "slidingsum", "--window=25", "--denom=W", "=", "percentile",
                           "20..80by30", "--max=6", "--quiet", "=", "addconst", "0.5"])
    assert_same(data, C + ["--novalue", "=", "percentile", "99", "--min=2", "=", "binarize", "--threshold=percentile99"])
    assert_same(data, C + ["--novalue", "--precision=6", "=", "smooth", "--window=11", "=", "percentile", "50", "--window=1000",
                           "=", "clip", "--min=percentile50"])
    # all cells qualify, last rank in the first chromosome: the bubble passes (general K)
    assert_same(data, C + ["--novalue", "=", "percentile", "30", "=", "addconst", "0"])
    assert_same(data, C + ["--novalue", "=", "percentile", "10..75by65", "--window=2", "=", "abs"])


def test_errors_match(data):
    for args in (C + ["=", "nosuchop"], C + ["--novalue", "=", "smooth", "--window=0"], C + ["--bogus"],
                 C + ["--novalue", "=", "binarize", "--threshold=nosuchvar"]):
        (rc_r, out_r, err_r), (rc_o, out_o, err_o) = both(data, args)
        assert rc_r != 0 and rc_o != 0, args
        assert err_r.split(b"\n")[0] == err_o.split(b"\n")[0], (args, err_r[:200], err_o[:200])
    with open(data / "bad.iv", "w") as f:
        f.write("chrD\t1000\t2000\n")
    (rc_r, _, err_r), (rc_o, _, err_o) = both(data, C + ["--novalue"], stdin="bad.iv")
    assert rc_r != 0 and rc_o != 0 and err_r == err_o


def test_percentile_preserve_roundtrip(data):
    """--preserve: the reference writes the vectors with 10 decimals and reads them back, so the signal
    that continues down the pipeline is the 10-decimal rounding of the smoothed track (and the trace
    shows write_all/read_all); window/min/max percentiles are fine with --preserve"""
    assert_same(data, C + ["--novalue", "--precision=17", "--progress=operations",
                           "=", "smooth", "--window=31",
                           "=", "percentile", "90", "--preserve=scratch.pres", "--precision=12",
                           "=", "multiply", "trackB.iv",
                           "=", "percentile", "10..90by20", "--window=7", "--min=0.25", "--preserve=scratch2.pres",
                           "=", "clip", "--max=percentile90"])
    assert os.path.getsize(data / "scratch.pres") == 0


def test_file_extrema_and_map_operators(data):
    """minover / maxover / minwith / maxwith / map through the CLI"""
    rng = np.random.default_rng(77)
    with open(data / "curve.map", "w") as f:
        f.write("# depth -> score\n")
        for x, y in [(0, 0), (6, 1.5), (2, 0.25), (12, 1.75), (40, -3)]:
            f.write("%d %r\n" % (x, y))
    with open(data / "loose.iv", "w") as f:          # unsorted, overlapping, valued
        for _ in range(600):
            n, l = CHROMS[int(rng.integers(0, 4))]
            a = int(rng.integers(0, max(1, l - 900)))
            f.write("%s\t%d\t%d\t%s\n" % (n, a, min(l, a + int(rng.integers(1, 900))), repr(float(rng.integers(0, 13)) / 2)))
    assert_same(data, C + ["--novalue", "--progress=operations", "=", "maxover", "trackB.iv"])
    assert_same(data, C + ["--novalue", "=", "slidingsum", "--window=25", "=", "minover", "trackB.iv", "--infinity=99.5"])
    assert_same(data, C + ["--novalue", "--progress=operations", "=", "minwith", "loose.iv", "=", "maxwith", "trackB.iv"])
    assert_same(data, C + ["--novalue", "--precision=12", "=", "map", "curve.map", "=", "maxwith", "loose.iv", "--value=4"])


def test_output_values_the_device_formatter_declines(data):
    """NaN and values of 2^63 and beyond are not formatted on the GPU: the chromosome that holds them is
    written by the host loop (glibc printf), the others by gdsp_format_runs -- byte-identical either way"""
    with open(data / "odd.iv", "w") as f:
        f.write("chrA 10 20 nan\nchrA 30 40 1e300\nchrA 50 60 -2.5\nchrB 5 9 0.125\nchrC 1 2 9223372036854775808\n"
                "chrC 7 9 9223372036854774784\nchrD 0 3 -0.0004\n")
    # a precision beyond what the device formatter takes, long enough to overrun a fixed line buffer
    assert_same(data, C + ["--precision=600"], stdin="odd.iv")
    for prec in ("0", "3", "17"):
        assert_same(data, C + ["--precision=" + prec], stdin="odd.iv")
        assert_same(data, C + ["--precision=" + prec, "--uncovered:show", "--origin=one", "--nocollapse"], stdin="odd.iv")


def test_real_valued_intervals_are_applied_in_file_order(data):
    """decimal (non-dyadic) values: abutting bedGraph-like intervals, overlapping ones, NaN and huge
    values next to small ones -- the reference adds interval after interval per cell, and so must we
    (a difference array would be off in the last bits, or poisoned by NaN / 1e300)"""
    rng = np.random.default_rng(123)
    with open(data / "bedgraph.iv", "w") as f:                 # abutting, disjoint, decimals
        for n, l in CHROMS:
            pos = 0
            while pos < l:
                e = min(l, pos + int(rng.integers(1, 300)))
                f.write("%s\t%d\t%d\t%s\n" % (n, pos, e, repr(round(float(rng.normal(0, 3)), 3))))
                pos = e
    with open(data / "overlap.iv", "w") as f:                  # unsorted, overlapping, decimals and extremes
        for _ in range(3000):
            n, l = CHROMS[int(rng.integers(0, 4))]
            a = int(rng.integers(0, max(1, l - 700)))
            v = [repr(round(float(rng.normal(0, 1)), 4)), "0.1", "1e300", "-1e300", "nan", "1e-300"][int(rng.integers(0, 12)) % 6 if rng.random() < 0.2 else 0]
            f.write("%s\t%d\t%d\t%s\n" % (n, a, min(l, a + int(rng.integers(1, 700))), v))
    P = ["--precision=17", "--uncovered:show"]
    assert_same(data, C + P, stdin="bedgraph.iv")
    assert_same(data, C + P, stdin="overlap.iv")
    assert_same(data, C + P + ["=", "add", "bedgraph.iv", "=", "subtract", "overlap.iv"], stdin="bedgraph.iv")
    assert_same(data, C + P + ["=", "input", "overlap.iv", "--missing=-1.5", "=", "add", "overlap.iv"], stdin="bedgraph.iv")


def test_values_that_return_to_missing_and_integer_adds_on_real_signals(data):
    """read_intervals' per-cell rule (genodsp.c:1307-1330): a running value equal to missingVal reads as "not yet
    covered", so a zero-valued row under --overlap=min|max, a depth that reaches --missing or a sum that passes through
    it is overwritten by the next interval; NaN sticks once assigned.  And add / subtract apply overlapping intervals
    one after the other (add.c:280-281): integer values on a signal that holds non-integers are ((v+a)+b)."""
    rng = np.random.default_rng(321)
    with open(data / "quirks.iv", "w") as f:                   # unsorted, deeply overlapping, small integers + NaN
        for _ in range(2000):
            n, l = CHROMS[int(rng.integers(0, 4))]
            a = int(rng.integers(0, max(1, l - 1500)))
            v = ["0", "0", "2", "5", "-1", "3", "7", "-4", "nan", "2.5"][int(rng.integers(0, 10))]
            f.write("%s\t%d\t%d\t%s\n" % (n, a, min(l, a + int(rng.integers(1, 1500))), v))
    with open(data / "ints.iv", "w") as f:                     # the same without NaN / fractions
        for _ in range(2000):
            n, l = CHROMS[int(rng.integers(0, 4))]
            a = int(rng.integers(0, max(1, l - 1500)))
            f.write("%s\t%d\t%d\t%d\n" % (n, a, min(l, a + int(rng.integers(1, 1500))), int(rng.integers(-6, 9))))
    P = ["--precision=17", "--uncovered:show"]
    for ov in ("min", "max"):
        for missing in ("0", "2"):
            assert_same(data, C + P + ["=", "input", "quirks.iv", "--overlap=" + ov, "--missing=" + missing], stdin="vals.iv")
    assert_same(data, C + P + ["=", "input", "ints.iv", "--overlap=min", "--missing=-1"], stdin="vals.iv")
    assert_same(data, C + P + ["=", "input", "ints.iv", "--missing=3"], stdin="vals.iv")
    assert_same(data, C + P + ["=", "input", "reads.iv", "--novalue", "--missing=7"], stdin="vals.iv")
    assert_same(data, C + ["--novalue", "--precision=17", "=", "smooth", "--window=31", "=", "add", "ints.iv", "--value=4",
                           "=", "subtract", "ints.iv", "--value=4"])
    assert_same(data, C + ["--novalue", "=", "add", "ints.iv", "--value=4", "=", "subtract", "ints.iv", "--value=4"])


def test_percentile_bubble_passes_on_real_values(data):
    """general rank, every position qualifying: the reference sorts every chromosome on its own and then runs
    combine_sorted_vectors steps (percentile.c:611-651, :820-864); here each step is a split search and two
    merges (gdsp_merge_exchange) and the per-chromosome sorts must not disturb their neighbours.  Real values
    (many radix passes), K = first and K = second chromosome, the state read back by the next operator."""
    P = ["--novalue", "--precision=9", "=", "smooth", "--window=11"]
    assert_same(data, C + P + ["=", "percentile", "30", "--precision=9", "=", "addconst", "0"])
    assert_same(data, C + P + ["=", "percentile", "70", "--precision=9", "=", "addconst", "0"])
    assert_same(data, C + P + ["=", "percentile", "10..70by30", "--quiet", "=", "abs"])
    assert_same(data, C + ["=", "percentile", "45", "--precision=3", "=", "addconst", "0"], stdin="vals.iv")
