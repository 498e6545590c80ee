"""Pins the oracle to the UNMODIFIED reference on signals holding NaN, +-inf and signed zeros (CPU only).

Where the two agree bit-for-bit the GPU tests (tests/test_gpu_nonfinite.py) hold the kernels to the
oracle.  The operators whose reference answer depends on its scan history are pinned to the RULE written
in DESIGN section 12 instead, and this file shows that the reference itself does not follow a rule there:
  * bestmax/bestmin with NaN (minmax.c:1672-1706: a NaN maximum forces a rescan every step, and the rescan
    returns NaN only when the window happens to start on a NaN);
  * bestmax with +0.0 and -0.0 in one window (which zero is returned depends on which compare kept it);
  * clump with NaN/inf (the running sum, clump.c:600, stays NaN for the rest of the chromosome).
"""
import numpy as np
import pytest

from checkers import Oracle, RefGenome, have_ref
from nonfinite import NONFINITE_KINDS, nonfinite_signal, rule_best_extrema, same_bits_or_both_nan

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CHROMS = [("chrA", 5000), ("chrB", 12345), ("chrC", 777), ("chrD", 64)]
DBL_MAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def run_both(kind, ref_words, oracle_fn, seed):
    rng = np.random.default_rng(seed)
    g = RefGenome(CHROMS)
    out = []
    try:
        inputs = {}
        for name, n in CHROMS:
            inputs[name] = nonfinite_signal(rng, n, kind)
            g.vec[name][:] = inputs[name]
        g.apply(*ref_words)
        for name, n in CHROMS:
            out.append((name, inputs[name], oracle_fn(inputs[name].copy()), np.array(g.vec[name])))
    finally:
        g.close()
    return out


CASES = [
    (["localmax", "--neighborhood=11", "--zero=-7"], lambda o: lambda v: o.local_extrema(v, 11, True, -7.0)),
    (["localmax", "--neighborhood=101"], lambda o: lambda v: o.local_extrema(v, 101, True, 0.0)),
    (["localmin", "--neighborhood=11"], lambda o: lambda v: o.local_extrema(v, 11, False, DBL_MAX)),
    (["close", "30", "--threshold=5"], lambda o: lambda v: o.close(v, 30, 5.0)),
    (["open", "3", "--threshold=2"], lambda o: lambda v: o.open(v, 3, 2.0)),
    (["dilate", "30", "--threshold=6"], lambda o: lambda v: o.dilate(v, 15, 15, 6.0)),
    (["smooth", "--window=11"], lambda o: lambda v: o.smooth(v, 11)),
    (["slidingsum", "--window=11"], lambda o: lambda v: o.sliding_sum(v, 11)),
    (["sum", "--window=10"], lambda o: lambda v: o.block_sum(v, 10)),
    (["cumulativesum"], lambda o: o.cumulative),
    (["binarize", "5"], lambda o: lambda v: o.binarize(v, 5.0)),
    (["binarize", "0", "--ties:above"], lambda o: lambda v: o.binarize(v, 0.0, True)),
    (["abs"], lambda o: o.abs),
    (["clip", "--min=2", "--max=6"], lambda o: lambda v: o.clip(v, 2.0, 6.0)),
    (["addconst", "1.5"], lambda o: lambda v: o.addconst(v, 1.5)),
]


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("case", range(len(CASES)))
def test_oracle_equals_reference_on_nonfinite(orc, kind, case):
    words, mk = CASES[case]
    for name, vin, got, want in run_both(kind, words, mk(orc), seed=case + 1):
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (words, kind, name)


@pytest.mark.parametrize("kind", ["inf", "negzero"])
@pytest.mark.parametrize("W", [3, 10, 101])
def test_best_extrema_inf_and_zeros(orc, kind, W):
    for want_max, word in ((True, "bestmax"), (False, "bestmin")):
        for name, vin, got, want in run_both(kind, [word, "--window=%d" % W], lambda v: orc.best_extrema(v, W, want_max), seed=W):
            # values always agree; with +0.0 and -0.0 in one window the reference's choice of zero is
            # history dependent, so only `==` is pinned for that kind
            assert np.array_equal(got, want), (word, kind, W, name)
            if kind == "inf":
                assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (word, kind, W, name)


def test_best_extrema_nan_reference_is_history_dependent(orc):
    """documents WHY the NaN rule of DESIGN section 12 exists: the reference agrees with the rule (NaN = the
    identity) wherever it yields a number, and yields NaN on a set of cells that depends on its scan history"""
    W = 10
    differs = 0
    for name, vin, got, want in run_both("nan", ["bestmax", "--window=%d" % W], lambda v: orc.best_extrema(v, W, True), seed=3):
        rule = rule_best_extrema(vin, W, True)
        num = ~np.isnan(want)
        assert np.array_equal(rule[num], want[num]), name
        differs += int(np.isnan(want).sum())
    assert differs > 0


def test_clump_inf_nan_reference_poisons_the_chromosome(orc):
    """clump's running sum (clump.c:600) turns NaN at the first NaN (or inf followed by -inf) and stays NaN:
    everything after it compares false.  Not reproduced; finite signals are."""
    for name, vin, got, want in run_both("negzero", ["clump", "0.5", "--length=20"], lambda v: orc.clump(v, 0.5, 20, True), seed=9):
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), name


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("collapse", [True, False])
def test_runs_nonfinite_match_reference_text(orc, tmp_path, kind, collapse):
    rng = np.random.default_rng(40)
    g = RefGenome(CHROMS)
    try:
        inputs = {}
        for name, n in CHROMS:
            inputs[name] = nonfinite_signal(rng, n, kind); g.vec[name][:] = inputs[name]
        path = str(tmp_path / "out.txt")
        g.report(path, precision=2, collapse=collapse, show_uncovered=0)
        ref = {}
        for line in open(path):
            c, s, e, val = line.rstrip("\n").split("\t")
            ref.setdefault(c, []).append((int(s), int(e), val))
        for name, n in CHROMS:
            rs, re, rv = orc.runs(inputs[name], collapse, 0)
            lines = [(int(s), int(e), "%.2f" % x) for s, e, x in zip(rs, re, rv)]
            assert lines == ref.get(name, []), (name, kind, collapse)
    finally:
        g.close()
