"""CPU check of the algebra behind the clump kernels (genodsp_b200/csrc/gdsp_clump.cu), against the oracle:

* the closed form of clump_search (clump.c:494-736): with P the prefix sums of d = v-T and
  M[j] = min(0, P[0..j]), end i is valid iff i+1 >= Lmin and M[i-Lmin] <= P[i]; cell p is marked iff
  max{P[i] : i >= p, i valid} >= M[p-1];
* the run trimming as carry propagation on 32-cell bit words: "fill upwards from the seeds S through the
  mask M" is (M & ~(M + S)) | S, a word hands a carry on like a full adder's generate/propagate pair, and
  the pairs fold per tile and per chromosome in both directions (k_clump_tilesum / _tilecarry / _emit).

Pure Python on small cases; the GPU tests (tests/test_gpu_ops.py::test_clump*) check the kernels themselves."""
import numpy as np
import pytest

from checkers import Oracle

M32 = 0xFFFFFFFF


def brev(x):
    return int("{:032b}".format(x)[::-1], 2)


def fill_up(mask, seeds, cin):
    seeds |= cin & mask & 1
    return ((mask & ~((mask + seeds) & M32)) | seeds) & M32


def gp(mask, seeds):
    """(generate, propagate) of one word"""
    return (fill_up(mask, seeds, 0) >> 31) & 1, 1 if mask == M32 else 0


def comb(a, b):
    """a then b (b later in the direction of travel)"""
    return b[0] | (b[1] & a[0]), a[1] & b[1]


def trim(marked, qual, tile_words):
    n = marked.size
    nt = ((n + 31) // 32 + tile_words - 1) // tile_words
    nw = nt * tile_words
    mk, qm = [0] * nw, [0] * nw
    for i in np.flatnonzero(marked):
        mk[i >> 5] |= 1 << (i & 31)
        if qual[i]:
            qm[i >> 5] |= 1 << (i & 31)
    up, dn = [], []
    for t in range(nt):
        a = (0, 1)
        for w in range(t * tile_words, (t + 1) * tile_words):
            a = comb(a, gp(mk[w], qm[w]))
        up.append(a)
        b = (0, 1)
        for w in range((t + 1) * tile_words - 1, t * tile_words - 1, -1):
            b = comb(b, gp(brev(mk[w]), brev(qm[w])))
        dn.append(b)
    cin_up, c = [0] * nt, 0
    for t in range(nt):
        cin_up[t] = c
        c = up[t][0] | (up[t][1] & c)
    cin_dn, c = [0] * nt, 0
    for t in range(nt - 1, -1, -1):
        cin_dn[t] = c
        c = dn[t][0] | (dn[t][1] & c)
    out = np.zeros(n, bool)
    for t in range(nt):
        words = range(t * tile_words, (t + 1) * tile_words)
        fu, fd = {}, {}
        c = cin_up[t]
        for w in words:
            fu[w] = fill_up(mk[w], qm[w], c)
            g, p = gp(mk[w], qm[w])
            assert (g | (p & c)) == (fu[w] >> 31) & 1          # the pair predicts the word's carry out
            c = g | (p & c)
        c = cin_dn[t]
        for w in reversed(words):
            fd[w] = brev(fill_up(brev(mk[w]), brev(qm[w]), c))
            g, p = gp(brev(mk[w]), brev(qm[w]))
            c = g | (p & c)
        for w in words:
            o = fu[w] & fd[w]
            for b in range(32):
                if (o >> b) & 1 and w * 32 + b < n:
                    out[w * 32 + b] = True
    return out


def clump_closed_form(v, T, lmin, above, tile_words):
    n = v.size
    d = (v - T) if above else (T - v)
    if (d < 0).all():                                    # clump.c:545-565
        return np.zeros(n, bool)
    P = np.cumsum(d)
    M = np.minimum(np.minimum.accumulate(P), 0.0)
    i = np.arange(n)
    mshift = np.where(i >= lmin, M[np.maximum(i - lmin, 0)], 0.0)
    q = np.where((i + 1 >= lmin) & (mshift <= P), P, -np.inf)
    sufmax = np.maximum.accumulate(q[::-1])[::-1]
    marked = sufmax >= np.concatenate(([0.0], M[:-1]))
    return trim(marked, (d >= 0) & marked, tile_words)


@pytest.mark.parametrize("tile_words", [1, 4])
@pytest.mark.parametrize("kind", ["int", "binary", "dyadic"])
def test_closed_form_and_bit_trimming_match_the_oracle(kind, tile_words):
    orc = Oracle()
    rng = np.random.default_rng({"int": 1, "binary": 2, "dyadic": 3}[kind] + tile_words)
    for trial in range(40):
        n = int(rng.integers(1, 1500))
        if kind == "int":
            v, T = rng.poisson(5, n).astype(float), 5.5
        elif kind == "binary":
            v, T = (rng.random(n) < 0.3).astype(float), 0.5
        else:
            v, T = rng.integers(-8, 9, n) / 8.0, 0.25
        if trial % 4 == 0 and n > 700:
            v[100:600] = T + 1                           # a run across many words and tiles
        lmin = int(rng.choice([1, 2, 10, 50, 100, 700]))
        for above in (True, False):
            want = orc.clump(v.copy(), T, lmin, above, 1.0, 0.0) != 0
            got = clump_closed_form(v, T, lmin, above, tile_words)
            assert np.array_equal(want, got), (kind, trial, n, lmin, above, np.flatnonzero(want != got)[:5])
