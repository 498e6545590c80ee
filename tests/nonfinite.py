"""Signals holding NaN, +-inf and signed zeros, shared by the CPU test that pins the oracle to the
reference on them and by the GPU parity tests."""
import numpy as np

NONFINITE_KINDS = ["nan", "inf", "negzero"]


def nonfinite_signal(rng, n, kind):
    if kind == "negzero":
        v = rng.integers(-1, 2, n).astype(np.float64)
    else:
        v = rng.poisson(5, n).astype(np.float64)
    k = max(1, n // 50)
    idx = rng.integers(0, n, k)
    if kind == "nan":
        v[idx] = np.nan
        if n > 40:
            v[20:23] = np.nan                       # consecutive NaNs
    elif kind == "inf":
        v[idx[::2]] = np.inf
        v[idx[1::2]] = -np.inf
    elif kind == "negzero":
        v[idx] = -0.0
        v = v * rng.choice([1.0, -1.0], n)          # +0.0 and -0.0 both present
    return v


def same_bits_or_both_nan(got, want):
    """bit-for-bit on every cell that is not NaN, NaN exactly where the reference has NaN (the sign and
    payload of a NaN are not compared: x86 propagates the operand's, the GPU returns the canonical one)"""
    got = np.ascontiguousarray(got, np.float64); want = np.ascontiguousarray(want, np.float64)
    gn, wn = np.isnan(got), np.isnan(want)
    if not np.array_equal(gn, wn):
        return False
    return bool(np.array_equal(got.view(np.uint64)[~wn], want.view(np.uint64)[~wn]))


def rule_best_extrema(v, W, want_max):
    """The rule this implementation follows for bestmax/bestmin on NaN: a NaN cell never
    wins -- it is the identity of the extremum (-inf for max, +inf for min) -- so a window holding only NaNs
    yields that identity.  (The reference's answer there depends on its scan history, minmax.c:1672-1706.)"""
    n = v.size
    l = (W - 1) // 2; r = (W - 1) - l
    neutral = -np.inf if want_max else np.inf
    a = np.where(np.isnan(v), neutral, v)
    out = np.empty(n)
    for i in range(n):
        lo = max(0, i - l); hi = min(n - 1, i + r)
        out[i] = a[lo:hi + 1].max() if want_max else a[lo:hi + 1].min()
    return out
