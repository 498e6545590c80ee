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


# ----------------------------------------------------------------------------------------------
# The slab-sharded variant (gdsp_clump_slab_* + slab.slab_clump_carries): a chromosome cut into pieces at
# multiples of TILE; only per-piece carries cross the cuts.  Emulated piece by piece exactly as the phases
# of the C-ABI see the data: a piece reads its own cells plus one halo tile to the left.
# ----------------------------------------------------------------------------------------------
def slab_clump_emulation(v, T, lmin, above, tile, cuts):
    n = v.size
    d_all = (v - T) if above else (T - v)
    bounds = [0] + list(cuts) + [n]
    pieces = [(bounds[k], bounds[k + 1]) for k in range(len(bounds) - 1)]
    # phase 1: head/tail aggregates {S, m}: m = minimum inclusive prefix sum relative to the range start
    def agg(a, b):
        if b <= a:
            return 0.0, np.inf
        c = np.cumsum(d_all[a:b])
        return float(c[-1]), float(c.min())
    aggs = []
    for k, (a, b) in enumerate(pieces):
        cont_r = b < n
        t = b - tile if cont_r else b
        aggs.append((agg(a, t), agg(t, b), bool((d_all[max(0, a - (tile if a else 0)):b] < 0).all())))
    allneg = all(x[2] for x in aggs)
    # phase 2 per piece: carry-in at the halo tile start, P/M over halo+owned, marks without the right carry
    marked_parts, sufmax_own, state = [], [], []
    for k, (a, b) in enumerate(pieces):
        P0, M0 = 0.0, 0.0
        for q in range(k):
            (Sh, mh), (St, mt), _ = aggs[q]
            M0 = min(M0, P0 + mh); P0 += Sh
            if q < k - 1:
                M0 = min(M0, P0 + mt); P0 += St
        e0 = a - tile if a else 0                         # first cell the piece can read
        d = d_all[e0:b]
        P = P0 + np.cumsum(d)
        M = np.minimum(np.minimum.accumulate(P), M0)      # includes everything before e0 (and P[-1] = 0)
        Mprev = np.concatenate(([M0], M[:-1]))            # M[p-1]
        idx = np.arange(e0, b)
        # M[i-lmin]: inside the readable range by construction (lmin <= tile)
        sh = np.where(idx >= lmin, np.concatenate((np.full(lmin, M0), M))[:idx.size] if lmin else M, 0.0)
        if lmin:
            shifted = np.concatenate((np.full(lmin, np.nan), M))[:idx.size]
            # cells whose i-lmin falls before e0 only occur in the halo tile (never used) or when e0 == 0
            sh = np.where(idx >= lmin, shifted, 0.0)
        q = np.where((idx + 1 >= lmin) & (sh <= P), P, -np.inf)
        own = idx >= a
        q_own = np.where(own, q, -np.inf)
        suf = np.maximum.accumulate(q_own[::-1])[::-1]
        marked = (suf >= Mprev) & own
        marked_parts.append(marked[own])
        sufmax_own.append(float(q_own.max()) if own.any() else -np.inf)
        state.append((Mprev[own], own))
    # phase 3: suffix maximum from the pieces to the right: everything with M[p-1] <= S_in is marked
    for k in range(len(pieces)):
        s_in = max(sufmax_own[k + 1:], default=-np.inf)
        if s_in > -np.inf:
            marked_parts[k] = marked_parts[k] | (state[k][0] <= s_in)
    marked = np.concatenate(marked_parts)
    qual = d_all >= 0
    # trimming with piece-level generate/propagate pairs (cells instead of words: same algebra)
    def piece_gp(mk, ql):
        g, p = 0, 1
        for m_, q_ in zip(mk, ql):
            g, p = ((1 if (m_ and q_) else 0) | ((1 if m_ else 0) & g)), p & (1 if m_ else 0)
        return g, p
    ups = [piece_gp(marked[a:b], qual[a:b]) for a, b in pieces]
    dns = [piece_gp(marked[a:b][::-1], qual[a:b][::-1]) for a, b in pieces]
    out = np.zeros(n, bool)
    for k, (a, b) in enumerate(pieces):
        cu = 0
        for q in range(k):
            cu = ups[q][0] | (ups[q][1] & cu)
        cd = 0
        for q in range(len(pieces) - 1, k, -1):
            cd = dns[q][0] | (dns[q][1] & cd)
        mk, ql = marked[a:b], qual[a:b]
        fu = np.zeros(b - a, bool); c = cu
        for i in range(b - a):
            c = 1 if (mk[i] and (ql[i] or c)) else 0
            fu[i] = bool(c)
        fd = np.zeros(b - a, bool); c = cd
        for i in range(b - a - 1, -1, -1):
            c = 1 if (mk[i] and (ql[i] or c)) else 0
            fd[i] = bool(c)
        out[a:b] = fu & fd
    if allneg:
        out[:] = False
    return out


@pytest.mark.parametrize("kind", ["int", "binary", "plateau"])
def test_slab_carries_match_the_oracle(kind):
    orc = Oracle()
    rng = np.random.default_rng({"int": 11, "binary": 12, "plateau": 13}[kind])
    tile = 64
    for trial in range(30):
        n = int(rng.integers(4 * tile, 30 * tile))
        if kind == "int":
            v, T = rng.poisson(5, n).astype(float), 5.5
        elif kind == "binary":
            v, T = (rng.random(n) < 0.3).astype(float), 0.5
        else:
            v, T = np.full(n, 4.0), 4.5
            v[n // 5: n // 5 + n // 2] = 5.0; v[::17] = 9.0; v[(n // 5 + n // 2):] = 3.0
        ncut = int(rng.integers(1, 4))
        cuts = sorted(set(int(c) * tile for c in rng.integers(1, n // tile, ncut)))
        lmin = int(rng.choice([1, 2, 10, 40, 64]))
        for above in (True, False):
            want = orc.clump(v.copy(), T, lmin, above, 1.0, 0.0) != 0
            got = slab_clump_emulation(v, T, lmin, above, tile, cuts)
            assert np.array_equal(want, got), (kind, trial, n, cuts, lmin, above, np.flatnonzero(want != got)[:8])


# ---------------------------------------------------------------------------------------------------------
# Quiet tiles (k_clump_classify + the early-out of k_clump_mark): from the group records alone
# -- {P, M} before every 512-cell group, P after 256 cells of a group, "the group holds no qualifying cell" --
# a 4096-cell tile is declared free of valid ends, and with e (the maximum valid P to its right) below M at the
# tile's end, free of marks.  Restated on the CPU and checked against the definitions cell by cell.
# ---------------------------------------------------------------------------------------------------------

def _quiet_tiles(d, L, tile=4096, group=512):
    """the classification rule of k_clump_classify for one whole chromosome (d = v - T per cell)"""
    n = d.size
    P = np.cumsum(d)
    Pm1 = np.concatenate([[0.0], P])                         # Pm1[j] = P[j-1], P[-1] = 0
    M = np.minimum.accumulate(Pm1)                           # M[j] (index j+1) = min(0, P[0..j]); M[-1] = 0 at index 0
    ngroups = n // group
    pbefore = Pm1[np.arange(ngroups + 1) * group]            # P before group g
    mbefore = M[np.arange(ngroups + 1) * group]              # M before group g
    falling = np.array([not np.any(d[g * group:(g + 1) * group] >= 0) for g in range(ngroups)])
    def half_p(H):                                           # P before 256-cell half H
        return Pm1[H * 256]
    quiet = np.zeros(n // tile, bool)
    for t in range(n // tile):
        c0 = t * tile
        ok = (L >= group) and (c0 >= L) and ((t + 1) * tile + group <= n)        # whole tile, not the last one
        for w in range(tile // group):
            if not ok:
                break
            G = c0 // group + w
            a = c0 + w * group
            ga = (a - L) // group
            ok = mbefore[ga] > pbefore[G] and bool(np.all(falling[ga:G + 1]))
            for h in range(2):
                if not ok:
                    break
                H = 2 * G + h
                bh = a + 256 * h + 255
                Hl = (bh - L) // 256
                ok = half_p(Hl + 1) > half_p(H)
        quiet[t] = ok
    return quiet, P, M


@pytest.mark.parametrize("L", [512, 1000, 2500, 4096])
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_quiet_tile_rule_is_sound(L, seed):
    rng = np.random.default_rng(10 * L + seed)
    n = 40 * 4096
    d = np.full(n, -1.0)
    pos = int(rng.integers(0, 3000))
    while pos < n:
        pos += int(rng.choice([700, 5000, 9000, 20000, 45000]))
        hump = int(rng.choice([40, 600, 1500, 6000, 12000]))
        d[pos:pos + hump] = float(rng.choice([0.5, 1.0, 4.0, 16.0]))
        pos += hump
    if seed >= 2:
        d[: n // 3] -= 3.0                                   # a deep early minimum: later falling stretches stay above it
        d[n // 3: n // 3 + 8000] = 12.0
    quiet, P, M = _quiet_tiles(d, L)
    # the definitions: end i is valid iff i+1 >= L and M[i-L] <= P[i]; S[p] = max valid P at or after p;
    # cell p is marked iff S[p] >= M[p-1]
    idx = np.arange(n)
    Mlag = np.where(idx - L >= -1, M[np.clip(idx - L + 1, 0, n)], np.inf)
    valid = (idx + 1 >= L) & (Mlag <= P)
    S = np.maximum.accumulate(np.where(valid, P, -np.inf)[::-1])[::-1]
    marked = S >= M[:n]                                      # M[p-1] sits at index p
    assert quiet.any() and not quiet.all()
    seen_fallthrough = False
    for t in np.nonzero(quiet)[0]:
        lo, hi = t * 4096, (t + 1) * 4096
        assert not valid[lo:hi].any(), (L, seed, t)
        e = S[hi] if hi < n else -np.inf                     # maximum valid P to the right of the tile
        m_end = M[hi]                                        # M at the tile's last cell
        if e < m_end:
            assert not marked[lo:hi].any(), (L, seed, t)
        else:
            seen_fallthrough = True                           # the kernel runs the full path with e in hand
    if seed < 2:
        assert seen_fallthrough or L > 1000
