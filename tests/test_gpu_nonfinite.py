"""GPU parity on signals holding NaN, +-inf, +0.0/-0.0.

tests/test_oracle_nonfinite.py pins the oracle to the reference on the same signal kinds; here the
kernels are held to the oracle: bit-for-bit on every non-NaN cell and NaN exactly where the reference
has NaN.  bestmax/bestmin follow the rule of DESIGN section 12 for NaN (the reference is history
dependent there) and are compared with `==` when a window mixes +0.0 and -0.0.  Windowed running sums
on non-finite input are covered by the exact-order mode (test_gpu_exact_order.py)."""
import numpy as np
import pytest

from checkers import Oracle
from nonfinite import NONFINITE_KINDS, nonfinite_signal, rule_best_extrema, same_bits_or_both_nan

pytestmark = pytest.mark.gpu

CHROMS = [("chr1", 70001), ("chr2", 8192), ("chr3", 4097), ("chr5", 1), ("chr6", 63), ("chr7", 33000)]
DBL_MAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture()
def genome():
    from genodsp_b200.genome import Genome
    g = Genome(CHROMS)
    yield g
    g.close()


def load(genome, seed, kind):
    rng = np.random.default_rng(seed)
    inputs = {}
    for name, n in CHROMS:
        inputs[name] = nonfinite_signal(rng, n, kind)
        genome.set_chrom(name, inputs[name])
    return inputs


def check(genome, inputs, fn, what, by_value=False):
    for name, n in CHROMS:
        want = fn(inputs[name].copy())
        got = genome.get_chrom(name)
        if by_value:
            ok = np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
        else:
            ok = same_bits_or_both_nan(got, want)
        if not ok:
            bad = np.nonzero(~((got == want) | (np.isnan(got) & np.isnan(want))))[0]
            at = int(bad[0]) if bad.size else -1
            raise AssertionError("%s %s: %d cells differ, first at %d: got %r want %r (input around: %r)" % (
                what, name, bad.size, at, got[at] if at >= 0 else None, want[at] if at >= 0 else None,
                inputs[name][max(0, at - 3):at + 4] if at >= 0 else None))


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("N", [3, 11, 101, 2049])
def test_local_extrema_nonfinite(genome, orc, kind, N):
    inputs = load(genome, N, kind)
    genome.localmax(N, zero=-2.0)
    check(genome, inputs, lambda v: orc.local_extrema(v, N, True, -2.0), "localmax N=%d %s" % (N, kind))
    inputs = load(genome, N + 1, kind)
    genome.localmin(N)
    check(genome, inputs, lambda v: orc.local_extrema(v, N, False, DBL_MAX), "localmin N=%d %s" % (N, kind))


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("W", [3, 4, 100, 1000, 2050])
def test_best_extrema_nonfinite(genome, orc, kind, W):
    inputs = load(genome, W, kind)
    genome.bestmax(W)
    if kind == "nan":
        check(genome, inputs, lambda v: rule_best_extrema(v, W, True), "bestmax W=%d nan (rule)" % W)
    else:
        check(genome, inputs, lambda v: orc.best_extrema(v, W, True), "bestmax W=%d %s" % (W, kind), by_value=(kind == "negzero"))
    inputs = load(genome, W + 1, kind)
    genome.bestmin(W)
    if kind == "nan":
        check(genome, inputs, lambda v: rule_best_extrema(v, W, False), "bestmin W=%d nan (rule)" % W)
    else:
        check(genome, inputs, lambda v: orc.best_extrema(v, W, False), "bestmin W=%d %s" % (W, kind), by_value=(kind == "negzero"))


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("L", [1, 5, 33, 1001])
def test_morphology_nonfinite(genome, orc, kind, L):
    T = 0.0 if kind == "negzero" else 5.0
    left, right = L // 2, L - L // 2
    inputs = load(genome, L, kind)
    genome.close_(L, T)
    check(genome, inputs, lambda v: orc.close(v, L, T), "close %d %s" % (L, kind))
    inputs = load(genome, L + 1, kind)
    genome.open_(L, T, one=2.0, zero=-1.0)
    check(genome, inputs, lambda v: orc.open(v, L, T, 2.0, -1.0), "open %d %s" % (L, kind))
    inputs = load(genome, L + 2, kind)
    genome.dilate(L, threshold=T)
    check(genome, inputs, lambda v: orc.dilate(v, left, right, T), "dilate %d %s" % (L, kind))
    inputs = load(genome, L + 4, kind)
    genome.erode(L, threshold=T)
    check(genome, inputs, lambda v: orc.erode(v, left, right, T), "erode %d %s" % (L, kind))


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("collapse", [True, False])
@pytest.mark.parametrize("show", [0, 1])
def test_runs_nonfinite(genome, orc, kind, collapse, show):
    """report_intervals compares raw values with `==` (genodsp.c:1640): a NaN cell is always its own run,
    -0.0 continues a run of +0.0 and is 'zero' for --uncovered:hide"""
    inputs = load(genome, 40, kind)
    got = genome.runs(collapse, show)
    for name, n in CHROMS:
        rs, re, rv = orc.runs(inputs[name], collapse, show)
        gs, ge, gv = got[name]
        assert np.array_equal(gs, rs) and np.array_equal(ge, re), (name, kind, collapse, show)
        assert same_bits_or_both_nan(gv, rv), (name, kind, collapse, show)


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
@pytest.mark.parametrize("W", [3, 55, 101])
def test_smooth_and_block_sum_nonfinite(genome, orc, kind, W):
    """(W = 55 and 101 run the shared-product kernel, 3 the direct FIR)"""
    inputs = load(genome, W, kind)
    genome.smooth(W)
    check(genome, inputs, lambda v: orc.smooth(v, W), "smooth %d %s" % (W, kind))
    inputs = load(genome, W + 1, kind)
    genome.sum(W + 1)
    check(genome, inputs, lambda v: orc.block_sum(v, W + 1), "sum %d %s" % (W + 1, kind))


@pytest.mark.parametrize("kind", NONFINITE_KINDS)
def test_pointwise_nonfinite(genome, orc, kind):
    G = type(genome)
    inputs = load(genome, 7, kind)
    genome.binarize(5.0)
    check(genome, inputs, lambda v: orc.binarize(v, 5.0), "binarize " + kind)
    inputs = load(genome, 8, kind)
    genome.binarize(0.0, ties_above=True)
    check(genome, inputs, lambda v: orc.binarize(v, 0.0, True), "binarize ties " + kind)
    inputs = load(genome, 9, kind)
    genome.pointwise([G.op_abs(), G.op_clip(2.0, 6.0), G.op_addconst(1.5)])
    check(genome, inputs, lambda v: orc.addconst(orc.clip(orc.abs(v), 2.0, 6.0), 1.5), "abs/clip/addconst " + kind)
    inputs = load(genome, 10, kind)
    genome.erase(1.0, 5.0, zero=-3.0)
    check(genome, inputs, lambda v: orc.erase(v, 1.0, 5.0, False, -3.0), "erase " + kind)


@pytest.mark.parametrize("kind", ["negzero"])
def test_sliding_sums_signed_zero(genome, orc, kind):
    inputs = load(genome, 11, kind)
    genome.slidingsum(11)
    check(genome, inputs, lambda v: orc.sliding_sum(v, 11), "slidingsum " + kind, by_value=True)
    inputs = load(genome, 12, kind)
    genome.cumulativesum()
    check(genome, inputs, orc.cumulative, "cumulativesum " + kind, by_value=True)
    inputs = load(genome, 13, kind)
    genome.clump(0.5, 20)
    check(genome, inputs, lambda v: orc.clump(v, 0.5, 20, True), "clump " + kind)
