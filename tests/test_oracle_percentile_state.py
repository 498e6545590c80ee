"""Pins tests/percentile_model.py (the state op_percentile_apply leaves behind, including the collect
permutation of --window/--min/--max) to the UNMODIFIED reference, in process.  CPU only."""
import numpy as np
import pytest

from checkers import RefGenome, have_ref
import percentile_model as pm

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CHROMS = [("chrA", 500), ("chrB", 1234), ("chrC", 77), ("chrD", 64), ("chrE", 900)]


def run_ref(sig, words):
    g = RefGenome(CHROMS)
    try:
        for name, n in CHROMS:
            g.vec[name][:] = sig[name]
        names = g.sorted_names()
        g.apply(*words)
        return names, {name: np.array(g.vec[name]) for name, _ in CHROMS}
    finally:
        g.close()


def signal(rng, kind):
    if kind == "int":
        return {name: rng.poisson(3, n).astype(np.float64) + 1 for name, n in CHROMS}
    return {name: rng.normal(0, 3, n) for name, n in CHROMS}


@pytest.mark.parametrize("kind", ["int", "real"])
@pytest.mark.parametrize("args", [["--window=7"], ["--min=2"], ["--max=4"], ["--window=3", "--min=1.5", "--max=5"],
                                  ["--window=1000"], []])
def test_collect_closed_form(kind, args):
    rng = np.random.default_rng(len(args) + (5 if kind == "int" else 0))
    sig = signal(rng, kind)
    names, ref = run_ref(sig, ["percentile", "50", "--quiet", "--debug=collect"] + args)
    lengths = [dict(CHROMS)[n] for n in names]
    cat = np.concatenate([sig[n] for n in names])
    stride, mn, mx = 1, -np.inf, np.inf
    for a in args:
        if a.startswith("--window="): stride = int(a.split("=")[1])
        if a.startswith("--min="): mn = float(a.split("=")[1])
        if a.startswith("--max="): mx = float(a.split("=")[1])
    got, n = pm.collect(lengths, cat, stride, mn, mx)
    want = np.concatenate([ref[nm] for nm in names])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (kind, args, n, np.flatnonzero(got != want)[:8])


@pytest.mark.parametrize("kind", ["int", "real"])
@pytest.mark.parametrize("p", [5, 50, 90, 99])
@pytest.mark.parametrize("args", [["--window=7"], ["--min=2"], ["--window=3", "--min=1.5", "--max=5"], []])
def test_post_state(kind, p, args):
    rng = np.random.default_rng(p + len(args))
    sig = signal(rng, kind)
    names, ref = run_ref(sig, ["percentile", str(p), "--quiet"] + args)
    lengths = [dict(CHROMS)[n] for n in names]
    cat = np.concatenate([sig[n] for n in names])
    stride, mn, mx = 1, -np.inf, np.inf
    for a in args:
        if a.startswith("--window="): stride = int(a.split("=")[1])
        if a.startswith("--min="): mn = float(a.split("=")[1])
        if a.startswith("--max="): mx = float(a.split("=")[1])
    _, n = pm.collect(lengths, cat, stride, mn, mx)
    rank = int(np.uint32(np.float64(np.uint64(n) * np.uint64(p * 1000)) / np.float64(100000.0)))
    got, _ = pm.post_state(lengths, cat, rank, stride, mn, mx)
    want = np.concatenate([ref[nm] for nm in names])
    assert np.array_equal(got, want), (kind, p, args, n, np.flatnonzero(got != want)[:8])
