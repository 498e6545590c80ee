"""Slab-sharded execution must give the same per-base results as whole chromosomes.

The ranks of a 2/3/4-way slab partition are emulated on ONE GPU: one Genome per virtual rank with that
rank's segment table (genodsp_b200/slab.py), halos moved by device-to-device copies that follow the
same plan the NCCL exchange uses.  This exercises every kernel's lo/hi/dlo/dhi/pos0 handling
(interval clipping, windows reading halo cells, blocks aligned to chromosome coordinate 0, strided
sampling) without needing several GPUs."""
import numpy as np
import pytest

from checkers import Oracle

pytestmark = pytest.mark.gpu

CHROMS = [("chr1", 90001), ("chr2", 50000), ("chr3", 30011), ("chr4", 777)]
HALO = 600


def make_ranks(world, halo=HALO, align=1, chroms=None):
    from genodsp_b200 import slab
    from genodsp_b200.genome import Genome
    chroms = chroms or CHROMS
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    lengths = [chroms[i][1] for i in order]
    ranks = []
    for r in range(world):
        segs_s, cells = slab.partition(lengths, world, r, halo, align)
        segs = [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s]
        g = Genome(chroms, segs=segs, buffer_cells=cells)
        g.plan = slab.halo_plan(lengths, world, r, halo, align)
        ranks.append(g)
    return ranks, order


def exchange(ranks):
    for r, g in enumerate(ranks):
        for peer, s_lo, s_hi, r_lo, r_hi in g.plan:
            theirs = [p for p in ranks[peer].plan if p[0] == r][0]
            g.sig[r_lo:r_hi].copy_(ranks[peer].sig[theirs[1]:theirs[2]])


def scatter_signal(ranks, inputs):
    for g in ranks:
        for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(g.segs):
            name = g.chroms[g.seg_chrom[k]][0]
            g.sig[lo:hi].copy_(g.torch.from_numpy(np.ascontiguousarray(inputs[name][pos0:pos0 + hi - lo])))


def gather_signal(ranks):
    out = {n: np.zeros(l) for n, l in ranks[0].chroms}
    for g in ranks:
        sig = g.sig.cpu().numpy()
        for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(g.segs):
            out[g.chroms[g.seg_chrom[k]][0]][pos0:pos0 + hi - lo] = sig[lo:hi]
    return out


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_pipeline_matches_whole_chromosomes(world):
    orc = Oracle()
    rng = np.random.default_rng(world)
    ranks, order = make_ranks(world)
    try:
        # intervals over the whole genome; every rank gets those touching its pieces (clipped by the kernel)
        m = 30000
        ci = rng.integers(0, len(CHROMS), m)
        lens = np.array([CHROMS[c][1] for c in ci])
        start = (rng.random(m) * lens).astype(np.uint32)
        end = np.minimum(lens, start + rng.integers(1, 400, m)).astype(np.uint32)
        want = {}
        for c, (name, n) in enumerate(CHROMS):
            want[name] = orc.accumulate(np.zeros(n), start[ci == c], end[ci == c])
        for g in ranks:
            segs, s, e = [], [], []
            for k in range(g.nseg):
                sel = ci == g.seg_chrom[k]
                segs.append(np.full(int(sel.sum()), k, np.uint32)); s.append(start[sel]); e.append(end[sel])
            g.accumulate(np.concatenate(segs), np.concatenate(s), np.concatenate(e))
        got = gather_signal(ranks)
        for name, _ in CHROMS:
            assert np.array_equal(bits(got[name]), bits(want[name])), ("accumulate", name)

        def stage(label, gpu_fn, cpu_fn, needs_halo=True):
            if needs_halo:
                exchange(ranks)
            for g in ranks:
                gpu_fn(g)
            for name in want:
                want[name] = cpu_fn(want[name])
            got = gather_signal(ranks)
            for name, _ in CHROMS:
                bad = np.nonzero(bits(got[name]) != bits(want[name]))[0]
                assert bad.size == 0, (label, world, name, bad[:5], got[name][bad[:5]], want[name][bad[:5]])

        stage("smooth", lambda g: g.smooth(101), lambda v: orc.smooth(v, 101))
        stage("localmax", lambda g: g.localmax(11), lambda v: orc.local_extrema(v, 11, True, 0.0))
        stage("addconst+binarize", lambda g: g.pointwise([type(g).op_addconst(0.5), type(g).op_binarize(0.75)]),
              lambda v: orc.binarize(orc.addconst(v, 0.5), 0.75), needs_halo=False)
        # restart from depth for the remaining windowed operators
        depth = {}
        for c, (name, n) in enumerate(CHROMS):
            depth[name] = orc.accumulate(np.zeros(n), start[ci == c], end[ci == c])
        for label, gpu_fn, cpu_fn in [
            ("slidingsum", lambda g: g.slidingsum(100), lambda v: orc.sliding_sum(v, 100, 1.0)),
            ("bestmax", lambda g: g.bestmax(1000), lambda v: orc.best_extrema(v, 1000, True)),
            ("bestmin", lambda g: g.bestmin(31), lambda v: orc.best_extrema(v, 31, False)),
            ("sum", lambda g: g.sum(101, denom_actual=True), lambda v: orc.block_sum(v, 101, 1.0, True, 0.0)),
        ]:
            scatter_signal(ranks, depth)
            for name in want:
                want[name] = depth[name].copy()
            stage(label, gpu_fn, cpu_fn)
        # strided min/max/count (percentile 0..100 path) combine across ranks by min/max/sum
        scatter_signal(ranks, depth)
        parts = [g.minmax(stride=7, mn=1.0, mx=1e9) for g in ranks]
        allv = np.concatenate([depth[n][::7] for n, _ in CHROMS])
        allv = allv[(allv >= 1.0)]
        assert min(p[0] for p in parts if p[2]) == allv.min() and max(p[1] for p in parts if p[2]) == allv.max()
        assert sum(p[2] for p in parts) == allv.size
    finally:
        for g in ranks:
            g.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["depth", "real"])
def test_slab_operators_on_virtual_ranks(world, kind):
    """percentile selection, cumulative sum, invert (global min/max) and run-length output over a
    slab-sharded genome: the per-rank CUDA calls (gdsp_pct_sample / gdsp_pct_count / gdsp_minmax /
    gdsp_cumulative_sum / gdsp_runs) combined by genodsp_b200.slab, against the whole-genome answers"""
    from genodsp_b200 import slab
    rng = np.random.default_rng(world + (7 if kind == "real" else 0))
    ranks, order = make_ranks(world, halo=0)
    try:
        sig = {}
        for name, n in CHROMS:
            if kind == "depth":
                sig[name] = np.repeat(rng.poisson(3, n // 5 + 1), 5)[:n].astype(np.float64)
            else:
                sig[name] = rng.normal(0, 3, n)
        scatter_signal(ranks, sig)
        allv = np.sort(np.concatenate([sig[CHROMS[i][0]] for i in order]))
        plist = [0, 500, 25000, 50000, 99000, 99990, 100000]
        got, n = slab.slab_percentiles(ranks, slab.virtual_gather, plist, sample_per_rank=4096)
        assert n == allv.size
        want = [float(allv[slab._pct_rank(allv.size, p)]) for p in plist]
        assert got == want, (got, want)
        # strided samples with limits (percentile --window / --min / --max)
        sel = np.concatenate([sig[CHROMS[i][0]][::3] for i in order]); sel = np.sort(sel[(sel >= 1.0) & (sel <= 6.0)])
        got, n = slab.slab_percentiles(ranks, slab.virtual_gather, [50000, 99000], stride=3, mn=1.0, mx=6.0, sample_per_rank=4096)
        assert n == sel.size and got == [float(sel[slab._pct_rank(sel.size, p)]) for p in (50000, 99000)]
        # run-length output: runs cut by a slab boundary are merged back
        runs = slab.slab_runs(ranks, slab.virtual_gather)
        for name, _ in CHROMS:
            v = sig[name]
            head = np.concatenate([[True], v[1:] != v[:-1]])
            st = np.nonzero(head)[0]; en = np.concatenate([st[1:], [v.size]]); val = v[st]
            keep = val != 0
            assert np.array_equal(runs[name][0], st[keep]) and np.array_equal(runs[name][1], en[keep]), name
            assert np.array_equal(bits(runs[name][2]), bits(val[keep])), name
        # invert about the global (min+max)/2
        mid = slab.slab_invert(ranks, slab.virtual_gather)
        assert mid == (allv[0] + allv[-1]) / 2.0
        got = gather_signal(ranks)
        for name, _ in CHROMS:
            assert np.array_equal(bits(got[name]), bits(2 * mid - sig[name])), name
        # cumulative sum with carries across the cuts
        scatter_signal(ranks, sig)
        slab.slab_cumulativesum(ranks, slab.virtual_gather)
        got = gather_signal(ranks)
        for name, _ in CHROMS:
            want = np.cumsum(sig[name])
            if kind == "depth":
                assert np.array_equal(bits(got[name]), bits(want)), name
            else:
                scale = np.cumsum(np.abs(sig[name]))
                assert np.max(np.abs(got[name] - want) / scale) <= 1e-12, name
    finally:
        for g in ranks:
            g.close()


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_morphology_with_halos(world):
    """open / close / dilate / erode on slab pieces: the marker search continues into the halo cells the
    neighbouring rank supplied (halo >= the operator's reach), so the owned cells equal the
    whole-chromosome result"""
    orc = Oracle()
    rng = np.random.default_rng(40 + world)
    ranks, order = make_ranks(world, halo=HALO)
    try:
        sig = {}
        for name, n in CHROMS:
            v = np.zeros(n)
            pos = 0
            while pos < n:                                   # runs and gaps of 1..400 cells
                L = int(rng.integers(1, 400)); v[pos:pos + L] = float(rng.integers(0, 2)) * float(rng.integers(1, 9)); pos += L
            sig[name] = v
        cases = [
            ("open 150", lambda g: g.open_(150, 0.5), lambda v: orc.open(v, 150, 0.5)),
            ("close 150", lambda g: g.close_(150, 0.5), lambda v: orc.close(v, 150, 0.5)),
            ("close 598", lambda g: g.close_(598, 0.5), lambda v: orc.close(v, 598, 0.5)),
            ("dilate 301", lambda g: g.dilate(301, threshold=0.5), lambda v: orc.dilate(v, 150, 151, 0.5)),
            ("erode 77", lambda g: g.erode(77, threshold=0.5), lambda v: orc.erode(v, 38, 39, 0.5)),
        ]
        for label, gpu_fn, cpu_fn in cases:
            scatter_signal(ranks, sig)
            exchange(ranks)
            for g in ranks:
                gpu_fn(g)
            got = gather_signal(ranks)
            for name, _ in CHROMS:
                want = cpu_fn(sig[name].copy())
                bad = np.nonzero(bits(got[name]) != bits(want))[0]
                assert bad.size == 0, (label, world, name, bad[:8], got[name][bad[:8]], want[bad[:8]])
        # a reach larger than the halo must be refused, not silently wrong
        from genodsp_b200 import capi
        scatter_signal(ranks, sig)
        with pytest.raises(capi.GdspError):
            for g in ranks:
                g.open_(5000, 0.5)
    finally:
        for g in ranks:
            g.close()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_clump_reassembles_cut_chromosomes(world):
    """clump / anticlump on slab pieces: chromosomes cut by a slab boundary are put back together on one
    rank (slab_clump), whole ones run in place; the result equals the whole-chromosome oracle"""
    from genodsp_b200 import slab
    from genodsp_b200.genome import Genome
    orc = Oracle()
    rng = np.random.default_rng(70 + world)
    ranks, order = make_ranks(world, halo=0)
    try:
        sig = {name: rng.poisson(4, n).astype(np.float64) for name, n in CHROMS}
        factory = lambda name, clen, rank: Genome([(name, clen)])
        for above in (True, False):
            scatter_signal(ranks, sig)
            slab.slab_clump(slab.VirtualTransport(ranks), slab.virtual_gather, factory, average=4.5, length=40,
                            above=above)
            got = gather_signal(ranks)
            for name, _ in CHROMS:
                want = orc.clump(sig[name].copy(), 4.5, 40, above)
                bad = np.nonzero(bits(got[name]) != bits(want))[0]
                assert bad.size == 0, (world, above, name, bad[:8])
    finally:
        for g in ranks:
            g.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_slab_percentile_then_binarize(world):
    """cfg3's `percentile 99 = binarize --threshold=percentile99` on slabs: the reference's sorted
    post-state thresholded = a step at K, computed from region counts instead of a distributed sort"""
    from genodsp_b200 import slab
    rng = np.random.default_rng(5 + world)
    ranks, order = make_ranks(world, halo=0)
    try:
        sig = {name: np.floor(rng.gamma(2.0, 2.0, n)) / 4.0 for name, n in CHROMS}
        scatter_signal(ranks, sig)
        allv = np.sort(np.concatenate([sig[CHROMS[i][0]] for i in order]))
        for p, ties in ((99000, False), (50000, True), (10, False)):
            scatter_signal(ranks, sig)
            thr, K = slab.slab_percentile_then_binarize(ranks, slab.virtual_gather, p, ties_above=ties, one=2.0, zero=-1.0)
            assert thr == allv[slab._pct_rank(allv.size, p)]
            want_sorted = np.where((allv >= thr) if ties else (allv > thr), 2.0, -1.0)
            got = gather_signal(ranks)
            cat = np.concatenate([got[CHROMS[i][0]] for i in order])
            assert np.array_equal(cat, want_sorted), (world, p, K)
    finally:
        for g in ranks:
            g.close()


CLUMP_CHROMS = [("chr1", 150001), ("chr2", 61440), ("chr3", 30011), ("chr4", 777)]


def _clump_signals(rng):
    """signals that make every carry of slab_clump_carries matter across the cuts"""
    sig = {}
    # (a) coverage-like noise: many short clumps, some straddling a cut
    sig["noise"] = {name: rng.poisson(4, n).astype(np.float64) for name, n in CLUMP_CHROMS}
    # (b) one long rise then a long decline: a clump that spans several pieces (suffix maximum from the right,
    #     trimming carries from both sides), qualifying cells sparse
    b = {}
    for name, n in CLUMP_CHROMS:
        v = np.full(n, 4.0)
        v[n // 5: n // 5 + n // 2] = 5.0            # long stretch above T = 4.5
        v[::97] = 9.0
        v[(n // 5 + n // 2):] = 3.0
        b[name] = v
    sig["plateau"] = b
    # (c) everything qualifies / nothing qualifies
    sig["all"] = {name: np.full(n, 7.0) for name, n in CLUMP_CHROMS}
    sig["none"] = {name: np.full(n, 1.0) for name, n in CLUMP_CHROMS}
    # (d) slow drift: the prefix-sum minimum is reached far to the left of every end
    d = {}
    for name, n in CLUMP_CHROMS:
        v = 4.0 + np.sign(np.sin(np.arange(n) / 9000.0)) * 1.0 + (rng.integers(0, 2, n) * 0.5)
        d[name] = v
    sig["drift"] = d
    # (e) long stretches below the threshold (whole tiles k_clump_mark settles from the group records alone) beside
    #     humps of every size: quiet tiles next to cuts, marks reaching into quiet stretches from a hump on their right
    e = {}
    for name, n in CLUMP_CHROMS:
        v = np.full(n, 3.0)
        pos = int(rng.integers(0, 2000))
        while pos < n:
            pos += int(rng.choice([900, 6000, 14000, 30000]))
            hump = int(rng.choice([50, 700, 2500, 9000]))
            v[pos:pos + hump] = float(rng.choice([5.0, 8.0, 20.0]))
            pos += hump
        e[name] = v
    sig["humps"] = e
    return sig


@pytest.mark.parametrize("world", [2, 3, 4])
def test_slab_clump_carries(world):
    """clump / anticlump on slab pieces with per-piece carries (gdsp_clump_slab_*): only a few numbers per
    piece cross the cuts, and the result equals the whole-chromosome oracle bit for bit"""
    from genodsp_b200 import slab
    orc = Oracle()
    rng = np.random.default_rng(170 + world)
    ranks, order = make_ranks(world, halo=4096, align=4096, chroms=CLUMP_CHROMS)
    try:
        cut = sum(1 for g in ranks for k in range(g.nseg) if g.segs[k][4] > 0)
        assert cut >= 1
        for label, sig in _clump_signals(rng).items():
            for above, L in ((True, 40), (False, 40), (True, 1), (True, 1000), (True, 4096)):
                if label in ("all", "none") and L not in (40, 4096):
                    continue
                scatter_signal(ranks, sig)
                exchange(ranks)
                slab.slab_clump_carries(ranks, slab.virtual_gather, average=4.5, length=L, above=above, one=2.0, zero=-1.0)
                got = gather_signal(ranks)
                for name, _ in CLUMP_CHROMS:
                    want = orc.clump(sig[name].copy(), 4.5, L, above, 2.0, -1.0)
                    bad = np.nonzero(bits(got[name]) != bits(want))[0]
                    assert bad.size == 0, (label, world, above, L, name, bad.size, bad[:8], got[name][bad[:8]], want[bad[:8]])
    finally:
        for g in ranks:
            g.close()


def test_gdsp_comm_single_rank_collectives():
    """the library's NCCL binding (gdsp_comm_*) with a one-rank communicator: every collective is the
    identity, which checks the staging, the stream ordering and the ABI.  (The N > 1 path is exercised by
    `bench.py --gpus N`, which compares every cell with the single-GPU run; NCCL refuses two ranks on one
    device, so it cannot run in this single-GPU suite.)"""
    import ctypes as C
    from genodsp_b200 import capi
    from genodsp_b200.genome import Genome
    g = Genome([("chr1", 5000)])
    try:
        lib = g.lib
        idbuf = (C.c_ubyte * 128)()
        capi.check(lib.gdsp_comm_unique_id(idbuf))
        h = C.c_void_p()
        capi.check(lib.gdsp_comm_create(g.ctx, idbuf, 1, 0, C.byref(h)))
        assert lib.gdsp_comm_rank(h) == 0 and lib.gdsp_comm_size(h) == 1
        counts = np.array([3, 0, 2 ** 40 + 5, 7], dtype=np.uint64)
        capi.check(lib.gdsp_comm_allreduce_sum_u64(h, counts.ctypes.data_as(C.POINTER(C.c_uint64)), 4))
        assert counts.tolist() == [3, 0, 2 ** 40 + 5, 7]
        vin = np.array([1.5, -2.0, np.pi]); vout = np.zeros(3)
        dp = C.POINTER(C.c_double)
        capi.check(lib.gdsp_comm_allgather_f64(h, vin.ctypes.data_as(dp), 3, vout.ctypes.data_as(dp)))
        assert np.array_equal(vin, vout)
        capi.check(lib.gdsp_comm_broadcast_f64(h, vin.ctypes.data_as(dp), 3, 0))
        t = g.torch
        a = t.arange(1000, dtype=t.float64, device=g.device); b = t.zeros_like(a)
        capi.check(lib.gdsp_comm_allgather_dev(h, C.c_void_p(a.data_ptr()), 1000, C.c_void_p(b.data_ptr())))
        g.sync()
        assert bool(t.equal(a, b))
        capi.check(lib.gdsp_comm_exchange_halos(h, C.c_void_p(g.sig.data_ptr()), None, 0))
        bad = (capi.Halo * 1)()
        bad[0].peer = 0                                   # a rank cannot exchange with itself
        assert lib.gdsp_comm_exchange_halos(h, C.c_void_p(g.sig.data_ptr()), bad, 1) != 0
        lib.gdsp_comm_destroy(h)
    finally:
        g.close()
