"""Exact-order mode (gdsp_ctx_set_exact_order / `--exact-order`): slidingsum, cumulativesum and
clump evaluate their running sums in the reference's sequential order, so general real-valued signals --
and inf / NaN, which poison the reference's running sum to the end of the chromosome -- come out bit for bit.
The oracle follows the reference's order and is pinned to it on these kinds (tests/test_oracle_vs_ref.py,
tests/test_oracle_nonfinite.py)."""
import os
import subprocess

import numpy as np
import pytest

from checkers import Oracle, REF_BIN, ROOT, have_ref
from nonfinite import nonfinite_signal, same_bits_or_both_nan

pytestmark = pytest.mark.gpu

CHROMS = [("chr1", 70001), ("chr2", 8192), ("chr3", 1025), ("chr5", 1), ("chr6", 63), ("chr7", 33000)]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.fixture()
def genome():
    from genodsp_b200.genome import Genome
    g = Genome(CHROMS)
    yield g
    g.close()


def signal(rng, n, kind):
    if kind == "real":
        return rng.normal(0, 3, n)
    if kind == "decimal":
        return np.round(rng.normal(5, 2, n), 3)
    if kind == "wide":                             # magnitudes from 1e-8 to 1e8: heavy cancellation
        return rng.normal(0, 1, n) * 10.0 ** rng.integers(-8, 9, n)
    return nonfinite_signal(rng, n, kind)


KINDS = ["real", "decimal", "wide", "nan", "inf", "negzero"]


def load(genome, seed, kind):
    rng = np.random.default_rng(seed)
    inputs = {}
    for name, n in CHROMS:
        inputs[name] = signal(rng, n, kind)
        genome.set_chrom(name, inputs[name])
    return inputs


def check(genome, inputs, fn, what):
    for name, n in CHROMS:
        want = fn(inputs[name].copy())
        got = genome.get_chrom(name)
        if not same_bits_or_both_nan(got, want):
            bad = np.nonzero(~((got.view(np.uint64) == want.view(np.uint64)) | (np.isnan(got) & np.isnan(want))))[0]
            raise AssertionError("%s %s: %d cells differ, first at %d: got %r want %r" % (what, name, bad.size, bad[0], got[bad[0]], want[bad[0]]))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("W", [3, 100, 101, 5000])
def test_sliding_sum_exact_order(genome, orc, kind, W):
    inputs = load(genome, W, kind)
    d = float(W) if W == 100 else 1.0
    with genome.exact_order():
        genome.slidingsum(W, denom=d)
    check(genome, inputs, lambda v: orc.sliding_sum(v, W, d), "slidingsum W=%d %s" % (W, kind))


@pytest.mark.parametrize("kind", KINDS)
def test_cumulative_sum_exact_order(genome, orc, kind):
    inputs = load(genome, 5, kind)
    with genome.exact_order():
        genome.cumulativesum()
    check(genome, inputs, orc.cumulative, "cumulativesum " + kind)


@pytest.mark.parametrize("kind", ["real", "decimal", "wide"])
@pytest.mark.parametrize("L", [1, 10, 100, 1000, 5000])
def test_clump_exact_order(genome, orc, kind, L):
    """clump on real-valued tracks: no tolerance is possible for a discrete output, so this mode is the
    parity evidence (VERDICT r1 missing #2)"""
    T = {"real": 0.25, "decimal": 5.1, "wide": 0.0}[kind]
    inputs = load(genome, L, kind)
    with genome.exact_order():
        genome.clump(T, L)
    check(genome, inputs, lambda v: orc.clump(v, T, L, True), "clump L=%d %s" % (L, kind))
    inputs = load(genome, L + 1, kind)
    with genome.exact_order():
        genome.anticlump(T, L, one=3.0, zero=-1.0)
    check(genome, inputs, lambda v: orc.clump(v, T, L, False, 3.0, -1.0), "anticlump L=%d %s" % (L, kind))


def test_default_mode_unchanged_after_exact(genome, orc):
    inputs = load(genome, 9, "real")
    with genome.exact_order():
        genome.cumulativesum()
    assert genome.lib.gdsp_ctx_get_exact_order(genome.ctx) == 0
    rng = np.random.default_rng(1)
    ints = {}
    for name, n in CHROMS:
        ints[name] = rng.poisson(5, n).astype(np.float64); genome.set_chrom(name, ints[name])
    genome.cumulativesum()
    check(genome, ints, orc.cumulative, "cumulativesum default")


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/genodsp not built")
def test_cli_exact_order_matches_reference_on_decimals(tmp_path):
    """the CLI flag: our `slidingsum --exact-order` / `cumulativesum --exact-order` / `clump --exact-order` print
    the bytes the reference prints without the flag, at full precision, on decimal bedGraph values"""
    ours = os.path.join(ROOT, "genodsp_b200", "bin", "genodsp")
    rng = np.random.default_rng(4)
    chroms = [("chrA", 30000), ("chrB", 9000)]
    (tmp_path / "g.chroms").write_text("".join("%s %d\n" % c for c in chroms))
    with open(tmp_path / "bg.iv", "w") as f:
        for n, l in chroms:
            pos = 0
            while pos < l:
                e = min(l, pos + int(rng.integers(1, 40)))
                f.write("%s\t%d\t%d\t%s\n" % (n, pos, e, repr(round(float(rng.normal(0.3, 2)), 3))))
                pos = e
    C = ["--chromosomes=g.chroms", "--precision=17", "--uncovered:show"]
    for ref_ops, our_ops in (
            (["=", "slidingsum", "--window=101"], ["=", "slidingsum", "--window=101", "--exact-order"]),
            (["=", "cumulativesum"], ["=", "cumulativesum", "--exact-order"]),
            (["=", "clump", "0.35", "--length=50"], ["=", "clump", "0.35", "--length=50", "--exact-order"]),
            (["=", "anticlump", "0.1", "--length=200"], ["=", "anticlump", "0.1", "--length=200", "--exact-order"])):
        outs = []
        for binary, ops in ((REF_BIN, ref_ops), (ours, our_ops)):
            with open(tmp_path / "bg.iv", "rb") as fin:
                p = subprocess.run([binary] + C + ops, stdin=fin, capture_output=True, cwd=tmp_path, timeout=600)
            assert p.returncode == 0, p.stderr.decode()[-1000:]
            outs.append(p.stdout)
        assert outs[0] == outs[1], our_ops
