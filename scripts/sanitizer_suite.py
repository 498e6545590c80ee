#!/usr/bin/env python
"""One pass over every kernel family of libgdsp_b200.so on small genomes, for compute-sanitizer
(memcheck / initcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python scripts/sanitizer_suite.py
    compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitizer_suite.py
    compute-sanitizer --tool initcheck python scripts/sanitizer_suite.py
    compute-sanitizer --tool synccheck python scripts/sanitizer_suite.py

Results are also compared with the oracle, so a pass means "clean AND correct".  The genomes are small on
purpose (the tools slow kernels down 10-100x) but shaped to reach the interesting code: chromosomes that
are a tile multiple, shorter than a tile, one cell long; window widths on both sides of every kernel
switch; slab layouts with halos; the look-back scans (cumulativesum, valued accumulate, clump) and the
2-bit clump words."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def main():
    import torch
    from checkers import Oracle
    from genodsp_b200 import capi, slab
    from genodsp_b200.genome import Genome
    orc = Oracle()
    rng = np.random.default_rng(3)
    chroms = [("chr1", 20001), ("chr2", 8192), ("chr3", 4097), ("chr5", 1), ("chr6", 63)]
    g = Genome(chroms)
    bits = lambda a: np.ascontiguousarray(a, np.float64).view(np.uint64)
    checked = [0]

    # guard cells: every cell of the two signal buffers that belongs to no chromosome (the alignment pads
    # between segments and the spare cells at the end) holds a canary; a kernel that writes one cell past a
    # segment end -- the typical tile-edge bug -- destroys it.  (compute-sanitizer is closed on this GPU pool,
    # so this and the oracle comparison are the memory-safety evidence we can produce; profiles/r2_sanitizer.md)
    CANARY = -6.02214076e23
    pad = torch.ones(g.buffer_cells, dtype=torch.bool, device=g.device)
    for (lo, hi, *_r) in g.segs:
        pad[lo:hi] = False
    g._sig[pad] = CANARY
    g.tmp[pad] = CANARY
    n_guard = int(pad.sum())

    def rearm():
        """percentile / sort use the partner buffer as dense scratch (pads included, by design): set the canaries again"""
        g._sig[pad] = CANARY
        g.tmp[pad] = CANARY

    def guards_intact(what):
        for buf in (g._sig, g.tmp):
            assert bool((buf[pad] == CANARY).all()), "a guard cell was overwritten during: " + what

    def load(kind="int"):
        ins = {}
        for name, n in chroms:
            v = rng.poisson(5, n).astype(np.float64) if kind == "int" else rng.normal(0, 3, n)
            ins[name] = v; g.set_chrom(name, v)
        return ins

    def check(ins, fn, what):
        for name, n in chroms:
            want = fn(ins[name].copy()); got = g.get_chrom(name)
            assert np.array_equal(bits(got), bits(want)), (what, name)
        guards_intact(what)
        checked[0] += 1

    # accumulate: binned unit path, valued int path, valued fp64 path
    m = 4000
    seg = rng.integers(0, g.nseg, m).astype(np.uint32)
    lens = np.array([g.segs[k][5] for k in seg])
    start = (rng.random(m) * lens).astype(np.uint32)
    end = np.minimum(lens, start + rng.integers(0, 300, m)).astype(np.uint32)
    for val in (None, rng.integers(-8, 9, m).astype(np.float64), rng.integers(-8, 9, m) / 4.0):
        g.fill(0.0)
        g.accumulate(seg, start, end, val)
        for k in range(g.nseg):
            name, n = g.chroms[g.seg_chrom[k]]
            sel = seg == k
            want = orc.accumulate(np.zeros(n), start[sel], end[sel], None if val is None else val[sel])
            assert np.array_equal(bits(g.get_chrom(name)), bits(want)), ("accumulate", name)
        checked[0] += 1
    # windowed sums
    for W in (3, 101, 5000):
        ins = load(); g.slidingsum(W); check(ins, lambda v: orc.sliding_sum(v, W), "slidingsum %d" % W)
        ins = load(); g.sum(W); check(ins, lambda v: orc.block_sum(v, W), "sum %d" % W)
    for W in (3, 101, 1001):
        ins = load("real"); g.smooth(W); check(ins, lambda v: orc.smooth(v, W), "smooth %d" % W)
    ins = load(); g.cumulativesum(); check(ins, orc.cumulative, "cumulativesum")
    with g.exact_order():
        ins = load("real"); g.cumulativesum(); check(ins, orc.cumulative, "cumulativesum exact")
        ins = load("real"); g.slidingsum(101); check(ins, lambda v: orc.sliding_sum(v, 101), "slidingsum exact")
        ins = load("real"); g.clump(0.25, 100); check(ins, lambda v: orc.clump(v, 0.25, 100, True), "clump exact")
    # extrema: small-window kernel, van Herk kernel (two block sizes), wide fallback
    for N in (3, 11, 101, 2049, 3001):
        ins = load(); g.localmax(N, zero=-2.0); check(ins, lambda v: orc.local_extrema(v, N, True, -2.0), "localmax %d" % N)
    for W in (4, 100, 2050, 6145):
        ins = load("real"); g.bestmin(W); check(ins, lambda v: orc.best_extrema(v, W, False), "bestmin %d" % W)
    # morphology
    for L in (1, 33, 1001):
        ins = load(); g.close_(L, 5.0); check(ins, lambda v: orc.close(v, L, 5.0), "close %d" % L)
        ins = load(); g.open_(L, 5.0); check(ins, lambda v: orc.open(v, L, 5.0), "open %d" % L)
        ins = load(); g.dilate(L, threshold=5.0); check(ins, lambda v: orc.dilate(v, L // 2, L - L // 2, 5.0), "dilate %d" % L)
        ins = load(); g.erode(L, threshold=5.0); check(ins, lambda v: orc.erode(v, L // 2, L - L // 2, 5.0), "erode %d" % L)
    # pointwise, interval tables
    G = type(g)
    ins = load(); g.pointwise([G.op_addconst(-1.5), G.op_abs(), G.op_clip(0.5, 6.0), G.op_invert(2.0), G.op_binarize(-1.0)])
    check(ins, lambda v: orc.binarize(orc.invert(orc.clip(orc.abs(orc.addconst(v, -1.5)), 0.5, 6.0), 2.0), -1.0), "pointwise chain")
    ts, te, tv, tseg = [], [], [], []
    for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(g.segs):
        cuts = np.unique(rng.integers(0, clen, 40))
        a, b = cuts[0::2], cuts[1::2]; mm = min(a.size, b.size)
        tseg.append(np.full(mm, k, np.uint32)); ts.append(a[:mm].astype(np.uint32)); te.append(b[:mm].astype(np.uint32)); tv.append(rng.integers(1, 9, mm) / 4.0)
    tseg, ts, te, tv = map(np.concatenate, (tseg, ts, te, tv))
    table = g.interval_table(tseg, ts, te, tv)
    ins = load()
    g.pointwise([(capi.PW_IVL_ADD, 0.0, 0, 0, 0, table), (capi.PW_IVL_MUL, 0.0, 0, 0, 0, table)])
    for k in range(g.nseg):
        name, n = g.chroms[g.seg_chrom[k]]; sel = tseg == k
        v = ins[name].copy(); orc.add_intervals(v, ts[sel], te[sel], tv[sel], 1.0); orc.sorted_intervals(v, ts[sel], te[sel], tv[sel], 0, 0.0)
        assert np.array_equal(bits(g.get_chrom(name)), bits(v)), ("ivl", name)
    guards_intact("interval-table chain")
    g.maxover(table); g.minover(table)
    guards_intact("maxover / minover")
    table.close(); checked[0] += 1
    # percentile (selection, ranked counts, fill step), sort, collect permutation, runs, text formatter
    ins = load()
    got = g.percentile(10.0, 90.0, step=40.0)
    g.binarize(got["percentile50"])
    ins = load("real"); g.percentile(99.0); _ = g.sig; g.percentile_collect(7, 1.0, 1e9)
    rearm()                                       # (selection, sort and collect used the partner buffer as dense scratch)
    ins = load(); g.binarize(6.0); r = g.runs()
    guards_intact("binarize + runs")
    for name, n in chroms:
        rs, re, rv = orc.runs(orc.binarize(ins[name].copy(), 6.0))
        assert np.array_equal(r[name][0], rs) and np.array_equal(r[name][1], re), ("runs", name)
    from genodsp_b200.genome import format_runs
    format_runs(g, "chr1", r["chr1"][0], r["chr1"][1], r["chr1"][2], 3)
    g.text_roundtrip(); g.map_values([0.0, 1.0, 5.0], [1.0, 0.0, 2.0]); g.minmax(); checked[0] += 1
    guards_intact("format / text round trip / map / minmax")
    # clump: fast path, stored-prefix path
    for L in (10, 1000, 5000):
        ins = load(); g.clump(5.5, L); check(ins, lambda v: orc.clump(v, 5.5, L, True), "clump %d" % L)
    guards_intact("the whole single-GPU suite")
    g.close()
    # slab pieces: halos, slab clump carries, slab percentile
    schroms = [("chr1", 40001), ("chr2", 16384), ("chr3", 777)]
    order = sorted(range(len(schroms)), key=lambda i: -schroms[i][1])
    lengths = [schroms[i][1] for i in order]
    ranks = []
    for r in range(2):
        segs_s, cells = slab.partition(lengths, 2, r, 4096, 4096)
        gg = Genome(schroms, segs=[(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s], buffer_cells=cells)
        gg.plan = slab.halo_plan(lengths, 2, r, 4096, 4096)
        ranks.append(gg)
    sig = {name: rng.poisson(4, n).astype(np.float64) for name, n in schroms}

    def scatter():
        for gg in ranks:
            for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(gg.segs):
                gg.sig[lo:hi].copy_(torch.from_numpy(np.ascontiguousarray(sig[gg.chroms[gg.seg_chrom[k]][0]][pos0:pos0 + hi - lo])))
        for r, gg in enumerate(ranks):
            for peer, s_lo, s_hi, r_lo, r_hi in gg.plan:
                theirs = [p for p in ranks[peer].plan if p[0] == r][0]
                gg.sig[r_lo:r_hi].copy_(ranks[peer].sig[theirs[1]:theirs[2]])

    def gathered():
        out = {n: np.zeros(l) for n, l in schroms}
        for gg in ranks:
            s = gg.sig.cpu().numpy()
            for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(gg.segs):
                out[gg.chroms[gg.seg_chrom[k]][0]][pos0:pos0 + hi - lo] = s[lo:hi]
        return out

    scatter()
    for gg in ranks:
        gg.smooth(101)
    got = gathered()
    for name, n in schroms:
        assert np.array_equal(bits(got[name]), bits(orc.smooth(sig[name].copy(), 101))), ("slab smooth", name)
    scatter()
    for gg in ranks:
        gg.close_(150, 4.5)
    scatter()
    slab.slab_clump_carries(ranks, slab.virtual_gather, average=4.5, length=40)
    got = gathered()
    for name, n in schroms:
        assert np.array_equal(bits(got[name]), bits(orc.clump(sig[name].copy(), 4.5, 40, True))), ("slab clump", name)
    scatter()
    slab.slab_percentile_then_binarize(ranks, slab.virtual_gather, 90000)
    scatter()
    slab.slab_cumulativesum(ranks, slab.virtual_gather)
    for gg in ranks:
        gg.close()
    torch.cuda.synchronize()
    print("sanitizer suite: %d operator groups ran and matched the oracle; %d guard cells around the chromosomes intact in both buffers"
          % (checked[0] + 5, n_guard))


if __name__ == "__main__":
    main()
