"""GPU parity tests: every operator runs through the C-ABI (libgdsp_b200.so) on a
multi-chromosome genome and is compared with the plain-C oracle on the same seeded
inputs.  Bit-exact unless a tolerance is written in the test."""
import numpy as np
import pytest

from checkers import Oracle

pytestmark = pytest.mark.gpu

# lengths chosen to hit tile edges: exact tile multiples, tiny, odd
CHROMS = [("chr1", 70001), ("chr2", 8192), ("chr3", 4096), ("chr4", 4097), ("chr5", 1), ("chr6", 63), ("chr7", 130000)]


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
    if kind == "int":
        return rng.poisson(5, n).astype(np.float64)
    if kind == "sparse":
        v = np.zeros(n)
        for _ in range(max(1, n // 200)):
            s = rng.integers(0, n); L = rng.integers(1, 150)
            v[s:s + L] += rng.integers(1, 4)
        return v
    if kind == "dyadic":
        return rng.integers(-4096, 4096, n).astype(np.float64) / 1024.0
    return rng.normal(0, 3, n)


KINDS = ["int", "sparse", "dyadic", "real"]


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def load(genome, rng, kind):
    inputs = {}
    for name, n in CHROMS:
        inputs[name] = signal(rng, n, kind)
        genome.set_chrom(name, inputs[name])
    return inputs


def compare(genome, inputs, fn, exact=True, rtol=0.0, what="", scale_fn=None, exact_fn=None):
    """exact: bit-for-bit.  Otherwise |got-want| <= rtol * scale, where scale is the magnitude of
    the summands feeding each output (the same operator applied to |v|): the reference's own
    running sums drift by ~eps*sqrt(steps)*|summands| from the true value, so an error relative to a
    cancelling result would measure the reference's drift, not ours."""
    for name, n in CHROMS:
        want = fn(inputs[name].copy())
        got = genome.get_chrom(name)
        if exact:
            bad = np.nonzero(bits(got) != bits(want))[0]
            assert bad.size == 0, "%s %s: %d mismatches, first at %d: got %r want %r" % (
                what, name, bad.size, bad[0], got[bad[0]], want[bad[0]])
        else:
            scale = np.maximum(np.abs(want), 1e-300)
            if scale_fn is not None:
                scale = np.maximum(scale, np.abs(scale_fn(np.abs(inputs[name]))))
            drift = 0.0
            if exact_fn is not None:
                # the reference's sequential running sums drift from the exact (extended precision) sum by
                # ~eps*sqrt(steps)*|summands|; we must be within rtol of the exact value, and within
                # rtol + that drift of the reference
                truth = exact_fn(inputs[name])
                err_true = np.max(np.abs(got - truth) / scale) if n else 0.0
                assert err_true <= rtol, "%s %s: rel err vs exact sum %g > %g" % (what, name, err_true, rtol)
                drift = np.abs(want - truth)
            err = np.max(np.maximum(np.abs(got - want) - drift, 0.0) / scale) if n else 0.0
            assert err <= rtol, "%s %s: rel err %g > %g" % (what, name, err, rtol)


def _exact_sliding_sum(v, W, d):
    """sliding sum in extended precision (np.longdouble prefix sums), rounded once"""
    n = v.size; h = (W - 1) // 2
    c = np.concatenate([[np.longdouble(0)], np.cumsum(v.astype(np.longdouble))])
    i = np.arange(n)
    hi = np.minimum(n, i + h + 1); lo = np.maximum(0, i + h - W + 1)
    return ((c[hi] - c[lo]) / np.longdouble(d)).astype(np.float64)


def test_fill_and_roundtrip(genome):
    genome.fill(3.5)
    for name, n in CHROMS:
        assert np.all(genome.get_chrom(name) == 3.5)


@pytest.mark.parametrize("novalue", [True, False])
def test_accumulate(genome, orc, novalue):
    rng = np.random.default_rng(1)
    m = 20000
    seg = rng.integers(0, genome.nseg, m).astype(np.uint32)
    lens = np.array([genome.segs[k][5] for k in seg])
    start = (rng.random(m) * lens).astype(np.uint32)
    end = np.minimum(lens, start + rng.integers(0, 300, m)).astype(np.uint32)
    val = None if novalue else rng.integers(-8, 9, m).astype(np.float64) / 4.0
    genome.fill(0.0)
    genome.accumulate(seg, start, end, val)
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        sel = seg == k
        v = np.zeros(n)
        orc.accumulate(v, start[sel], end[sel], None if novalue else val[sel])
        got = genome.get_chrom(name)
        assert np.array_equal(bits(got), bits(v)), name
    # adding on top of an existing signal (add operator path)
    before = {name: genome.get_chrom(name) for name, _ in CHROMS}
    genome.accumulate(seg, start, end, val, add=True)
    for name, n in CHROMS:
        assert np.array_equal(genome.get_chrom(name), 2 * before[name])


@pytest.mark.parametrize("path", ["host", "dev"])
@pytest.mark.parametrize("force_diff", [False, True])
@pytest.mark.parametrize("shape", ["uniform", "sorted_pileup", "long", "edges"])
def test_accumulate_unit_paths(genome, orc, monkeypatch, path, force_diff, shape):
    """unit-weight depth: the binned path (events bucketed per 16384-cell tile, counts in shared
    memory) and the difference-array path must both give the reference loop's result
    (genodsp.c:1325-1329), from host arrays and from device arrays"""
    import torch
    if force_diff:
        monkeypatch.setenv("GDSP_ACCUMULATE_DIFF", "1")
    rng = np.random.default_rng({"uniform": 11, "sorted_pileup": 12, "long": 13, "edges": 14}[shape])
    m = 30000
    seg = rng.integers(0, genome.nseg, m).astype(np.uint32)
    lens = np.array([genome.segs[k][5] for k in seg])
    if shape == "uniform":
        start = (rng.random(m) * lens).astype(np.uint32)
        end = np.minimum(lens, start + rng.integers(0, 300, m)).astype(np.uint32)
    elif shape == "sorted_pileup":
        # position-sorted reads with heavy duplicates: whole warps hit one bucket / one cell
        seg = np.sort(seg); lens = np.array([genome.segs[k][5] for k in seg])
        start = ((rng.integers(0, 40, m) / 40.0) * lens).astype(np.uint32)
        order = np.lexsort((start, seg)); seg, start, lens = seg[order], start[order], lens[order]
        end = np.minimum(lens, start + 100).astype(np.uint32)
    elif shape == "long":
        # intervals much longer than a tile, many reaching the chromosome end
        start = (rng.random(m) * lens * 0.5).astype(np.uint32)
        end = np.minimum(lens, start + rng.integers(0, 200000, m)).astype(np.uint32)
    else:
        # empty intervals, whole-chromosome intervals, single cells at both ends, unknown segment ids
        start = np.where(rng.random(m) < 0.5, 0, np.maximum(lens, 1) - 1).astype(np.uint32)
        end = np.where(rng.random(m) < 0.3, start, np.where(rng.random(m) < 0.5, lens, np.minimum(lens, start + 1))).astype(np.uint32)
        seg[::97] = genome.nseg + 3
    genome.fill(7.0)                        # must be overwritten, not added to
    if path == "host":
        genome.accumulate(seg, start, end)
    else:
        dev = torch.device("cuda", 0)
        t = [torch.from_numpy(a.astype(np.int64)).to(dev).to(torch.int32) for a in (seg, start, end)]
        genome.accumulate(t[0], t[1], t[2], host=False)
        torch.cuda.synchronize()
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        sel = seg == k
        v = np.zeros(n)
        orc.accumulate(v, start[sel], end[sel], None)
        got = genome.get_chrom(name)
        assert np.array_equal(bits(got), bits(v)), (name, np.nonzero(got != v)[0][:5])
    before = {name: genome.get_chrom(name) for name, _ in CHROMS}
    if path == "host":
        genome.accumulate(seg, start, end, add=True)
    else:
        genome.accumulate(t[0], t[1], t[2], host=False, add=True)
    for name, n in CHROMS:
        assert np.array_equal(genome.get_chrom(name), 2 * before[name])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("W", [3, 100, 101, 1000, 5000])
def test_sliding_sum(genome, orc, kind, W):
    inputs = load(genome, np.random.default_rng(W), kind)
    genome.slidingsum(W, denom=float(W) if W == 100 else 1.0)
    d = float(W) if W == 100 else 1.0
    # integer / dyadic signals: every partial sum is exact, so any association order gives the
    # reference's bits; general reals: 1e-12 relative (BASELINE.json north_star)
    compare(genome, inputs, lambda v: orc.sliding_sum(v, W, d), exact=(kind != "real"), rtol=1e-12,
            what="slidingsum W=%d" % W, scale_fn=lambda a: orc.sliding_sum(a, W, d),
            exact_fn=lambda v: _exact_sliding_sum(v, W, d))


def test_text_roundtrip_is_printf_strtod(genome, orc):
    """gdsp_text_roundtrip (percentile --preserve) == strtod(printf("%.10f")) per cell, computed with
    integer arithmetic on the device: random magnitudes, exact ties at the 11th decimal (odd/2048),
    values that round to +-0, integers, infinities, huge values"""
    rng = np.random.default_rng(10)
    inputs = {}
    for name, n in CHROMS:
        parts = [rng.normal(0, 3, n), rng.integers(-5, 9, n).astype(np.float64), (2 * rng.integers(0, 4096, n) + 1) / 2048.0,
                 rng.normal(0, 1e-9, n), rng.normal(0, 1e-11, n), rng.normal(0, 1e7, n), rng.normal(0, 1e-3, n)]
        v = np.choose(rng.integers(0, len(parts), n), parts)
        special = np.array([0.0, -0.0, np.inf, -np.inf, 1e300, 5e-11, -5e-11, 2.5e-11, 7.5e-11, 123456.78901234567,
                            4.9e-324, -2.2250738585072014e-308, 0.1, 1e-10, 1.5e-10, 0.99999999995, 4503599627370495.5])
        k = min(n, special.size)
        v[:k] = special[:k]
        inputs[name] = v
        genome.set_chrom(name, v)
    genome.text_roundtrip()
    for name, n in CHROMS:
        want = orc.text_roundtrip10(inputs[name].copy())
        got = genome.get_chrom(name)
        bad = np.nonzero(bits(got) != bits(want))[0]
        assert bad.size == 0, (name, inputs[name][bad[:5]], got[bad[:5]], want[bad[:5]])


@pytest.mark.parametrize("precision", [0, 1, 3, 10, 17])
def test_format_runs_matches_printf(genome, precision):
    """gdsp_format_runs: device-side "%s\\t%u\\t%u\\t%.*f\\n" of run lists == C printf (Python's % uses the same
    correctly rounded conversion): ties at the last printed digit, negatives rounding to zero, -0.0,
    subnormals, integers up to 2^62, many magnitudes"""
    from genodsp_b200.genome import format_runs
    rng = np.random.default_rng(precision)
    n = 20000
    parts = [rng.normal(0, 3, n), rng.integers(-50, 90, n).astype(np.float64), (2 * rng.integers(0, 4096, n) + 1) / 2048.0,
             rng.normal(0, 1e-9, n), rng.normal(0, 1e7, n), rng.normal(0, 1e-3, n), rng.integers(0, 2 ** 40, n) / 1024.0,
             rng.integers(-1000, 1000, n) / 8.0]
    val = np.choose(rng.integers(0, len(parts), n), parts)
    special = np.array([0.0, -0.0, 0.5, 1.5, 2.5, -0.5, 0.05, 0.15, 0.25, 1e-300, -1e-300, 4.9e-324, 2.0 ** 62, -(2.0 ** 62),
                        0.0005, -0.0004, 123456789.987654321, 9.999999999999999e15, 0.99999999999999989, 1e15 + 0.5])
    val[:special.size] = special
    start = np.sort(rng.integers(0, 2 ** 31, n)).astype(np.uint32)
    end = (start + rng.integers(1, 1000, n)).astype(np.uint32)
    for with_value in (True, False):
        got = format_runs(genome, "chr12_random", start, end, val, precision, add_start=1, add_end=0, with_value=with_value)
        assert got is not None
        want = "".join("chr12_random\t%d\t%d%s\n" % (int(start[k]) + 1, int(end[k]), ("\t%.*f" % (precision, val[k])) if with_value else "")
                       for k in range(n)).encode()
        if got != want:
            gl, wl = got.split(b"\n"), want.split(b"\n")
            for k, (a, b) in enumerate(zip(gl, wl)):
                assert a == b, (k, val[k], a, b)
        assert got == want
    # values the device declines: the caller must fall back
    val[7] = np.inf
    assert format_runs(genome, "c", start, end, val, precision) is None


def test_smooth_to_host_matches_smooth(genome, orc):
    """the pipelined smooth + device->host delivery used by bench.py's e2e leg: same bits as smooth()"""
    import torch
    inputs = load(genome, np.random.default_rng(77), "real")
    out_h = torch.empty(genome.buffer_cells, dtype=torch.float64, pin_memory=True)
    copied = genome.smooth_to_host(101, out_h)
    torch.cuda.synchronize()
    assert copied == 8 * sum(n for _, n in CHROMS)
    host = out_h.numpy()
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        lo, hi = genome.segs[k][0], genome.segs[k][1]
        want = orc.smooth(inputs[name].copy(), 101)
        assert np.array_equal(bits(host[lo:hi]), bits(want)), name
        assert np.array_equal(bits(genome.get_chrom(name)), bits(want)), name


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("W", [3, 100, 101, 4096, 5000])
def test_block_sum(genome, orc, kind, W):
    inputs = load(genome, np.random.default_rng(W + 7), kind)
    genome.sum(W, denom=1.0, denom_actual=(W == 101), zero=-1.0 if W == 3 else 0.0)
    exact = (kind != "real") or W <= 4096       # W<=4096 folds each block in the reference's order
    compare(genome, inputs, lambda v: orc.block_sum(v, W, 1.0, W == 101, -1.0 if W == 3 else 0.0),
            exact=exact, rtol=1e-12, what="sum W=%d" % W,
            scale_fn=lambda a: orc.block_sum(a, W, 1.0, W == 101, 0.0))


@pytest.mark.parametrize("W", [4, 8, 12, 16, 20, 36, 64, 128, 260, 1000, 4092])
def test_block_sum_four_lanes_per_block(genome, orc, W):
    """k_block_sum4 (widths that are a multiple of 4, from 16 up: a block's quads are dealt out to four lanes and the
    running sum is handed on between them) and its neighbours below 16; general reals, so that any change in the
    order of the additions shows; `--denom=actual` makes the short last block of every chromosome count"""
    inputs = load(genome, np.random.default_rng(W), "real")
    genome.sum(W, denom=3.0, denom_actual=(W % 8 == 0), zero=0.25)
    compare(genome, inputs, lambda v: orc.block_sum(v, W, 3.0, W % 8 == 0, 0.25), exact=True, what="sum W=%d" % W)


@pytest.mark.parametrize("kind", ["int", "real"])
def test_block_sum_whole_chromosome(genome, orc, kind):
    inputs = load(genome, np.random.default_rng(3), kind)
    genome.sum(window_is_chromosome=True)
    compare(genome, inputs, lambda v: orc.block_sum(v, v.size), exact=(kind == "int"), rtol=1e-12, what="sum chrom",
            scale_fn=lambda a: orc.block_sum(a, a.size))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("W", [3, 11, 101, 513, 1001])
@pytest.mark.parametrize("direct", [False, True])
def test_smooth_bit_exact(genome, orc, kind, W, direct):
    """both kernels behind gdsp_smooth: the shared-product one (k_smooth_sym, the default for the reference's
    symmetric windows) and the direct FIR (k_smooth_ct)"""
    inputs = load(genome, np.random.default_rng(W), kind)
    genome.smooth(W, direct=direct)
    # chromosomes shorter than the window hit the reference's u32 wrap (sum.c:657); the oracle defines them
    compare(genome, inputs, lambda v: orc.smooth(v, W), exact=True, what="smooth W=%d" % W)


@pytest.mark.parametrize("W", [5, 7, 9, 13, 15, 17, 19, 21, 23, 31, 51, 53, 55, 77, 91, 93, 95, 97, 99, 103, 151])
def test_smooth_shared_product_widths(genome, orc, W):
    """k_smooth_sym holds (W-1)/2 pair accumulators per side in registers: every alignment class of the
    stream lead ((W-1)/2 mod 4 decides how long a finished output waits for its aligned store), the widest
    window it takes (101, in test_smooth_bit_exact) and the first ones it leaves to the direct FIR (7, 103)"""
    inputs = load(genome, np.random.default_rng(W), "real")
    genome.smooth(W)
    compare(genome, inputs, lambda v: orc.smooth(v, W), exact=True, what="smooth W=%d" % W)


@pytest.mark.parametrize("kind", KINDS)
def test_cumulative_sum(genome, orc, kind):
    inputs = load(genome, np.random.default_rng(5), kind)
    genome.cumulativesum()
    compare(genome, inputs, orc.cumulative, exact=(kind != "real"), rtol=1e-12, what="cumulativesum",
            scale_fn=orc.cumulative)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("N", [3, 11, 101, 2049, 3001])
def test_local_extrema(genome, orc, kind, N):
    inputs = load(genome, np.random.default_rng(N), kind)
    genome.localmax(N, zero=-2.0)
    compare(genome, inputs, lambda v: orc.local_extrema(v, N, True, -2.0), what="localmax N=%d" % N)
    inputs = load(genome, np.random.default_rng(N + 1), kind)
    genome.localmin(N)
    compare(genome, inputs, lambda v: orc.local_extrema(v, N, False, np.finfo(np.float64).max), what="localmin N=%d" % N)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("W", [3, 4, 100, 101, 1000, 2050, 6145])
def test_best_extrema(genome, orc, kind, W):
    inputs = load(genome, np.random.default_rng(W), kind)
    genome.bestmax(W)
    compare(genome, inputs, lambda v: orc.best_extrema(v, W, True), what="bestmax W=%d" % W)
    inputs = load(genome, np.random.default_rng(W + 1), kind)
    genome.bestmin(W)
    compare(genome, inputs, lambda v: orc.best_extrema(v, W, False), what="bestmin W=%d" % W)


@pytest.mark.parametrize("W", [64, 65, 66, 79, 80, 81, 95, 96, 97, 127, 128, 129, 513, 1024, 2033, 2048, 2049])
def test_best_extrema_block_kernel_widths(genome, orc, W):
    """k_extrema_blocks (64 <= W <= 2049): windows that end exactly on a 16-cell block boundary (W-1 a multiple of
    16), just before and just after one, the narrowest (two whole blocks inside) and the widest it takes; also as
    a wide localmax (the same kernel in its keep-unless-beaten mode)"""
    inputs = load(genome, np.random.default_rng(W), "real")
    genome.bestmax(W)
    compare(genome, inputs, lambda v: orc.best_extrema(v, W, True), what="bestmax W=%d" % W)
    inputs = load(genome, np.random.default_rng(W + 3), "int")
    genome.bestmin(W)
    compare(genome, inputs, lambda v: orc.best_extrema(v, W, False), what="bestmin W=%d" % W)
    if W % 2 == 1:
        inputs = load(genome, np.random.default_rng(W + 5), "real")
        genome.localmax(W, zero=-2.0)
        compare(genome, inputs, lambda v: orc.local_extrema(v, W, True, -2.0), what="localmax N=%d" % W)


def test_best_extrema_very_wide(orc):
    from genodsp_b200.genome import Genome
    g = Genome([("a", 30000), ("b", 100)])
    rng = np.random.default_rng(0)
    ins = {"a": signal(rng, 30000, "real"), "b": signal(rng, 100, "real")}
    for k, v in ins.items():
        g.set_chrom(k, v)
    g.bestmax(8001)
    for k, v in ins.items():
        assert np.array_equal(g.get_chrom(k), orc.best_extrema(v.copy(), 8001, True))
    g.close()


@pytest.mark.parametrize("kind", ["int", "sparse", "real"])
@pytest.mark.parametrize("L", [1, 2, 5, 31, 32, 33, 101, 1001, 20000])
def test_morphology(genome, orc, kind, L):
    T = {"int": 5.0, "sparse": 0.0, "real": 0.5}[kind]
    left, right = L // 2, L - L // 2
    inputs = load(genome, np.random.default_rng(L), kind)
    genome.close_(L, T)
    compare(genome, inputs, lambda v: orc.close(v, L, T), what="close %d" % L)
    inputs = load(genome, np.random.default_rng(L + 1), kind)
    genome.open_(L, T, one=2.0, zero=-1.0)
    compare(genome, inputs, lambda v: orc.open(v, L, T, 2.0, -1.0), what="open %d" % L)
    inputs = load(genome, np.random.default_rng(L + 2), kind)
    genome.dilate(L, threshold=T)
    compare(genome, inputs, lambda v: orc.dilate(v, left, right, T), what="dilate %d" % L)
    inputs = load(genome, np.random.default_rng(L + 3), kind)
    genome.dilate(left=3, right=9, threshold=T)
    compare(genome, inputs, lambda v: orc.dilate(v, 3, 9, T), what="dilate 3/9")
    inputs = load(genome, np.random.default_rng(L + 4), kind)
    genome.erode(L, threshold=T)
    compare(genome, inputs, lambda v: orc.erode(v, left, right, T), what="erode %d" % L)


@pytest.mark.parametrize("kind", KINDS)
def test_pointwise_single_ops(genome, orc, kind):
    G = type(genome)
    cases = [
        ([G.op_binarize(3.0)], lambda v: orc.binarize(v, 3.0)),
        ([G.op_binarize(3.0, True, 5.0, -5.0)], lambda v: orc.binarize(v, 3.0, True, 5.0, -5.0)),
        ([G.op_addconst(2.5)], lambda v: orc.addconst(v, 2.5)),
        ([G.op_abs()], orc.abs),
        ([G.op_clip(1.0, 4.0)], lambda v: orc.clip(v, 1.0, 4.0)),
        ([G.op_clip(1.0, None)], lambda v: orc.clip(v, 1.0, None)),
        ([G.op_clip(None, 2.0)], lambda v: orc.clip(v, None, 2.0)),
        ([G.op_erase(1.0, 4.0)], lambda v: orc.erase(v, 1.0, 4.0)),
        ([G.op_erase(1.0, 4.0, True, -1.0)], lambda v: orc.erase(v, 1.0, 4.0, True, -1.0)),
        ([G.op_erase(2.0, None)], lambda v: orc.erase(v, 2.0, None)),
        ([G.op_erase(None, 2.0, True)], lambda v: orc.erase(v, None, 2.0, True)),
        ([G.op_invert(1.25)], lambda v: orc.invert(v, 1.25)),
    ]
    for i, (ops, fn) in enumerate(cases):
        inputs = load(genome, np.random.default_rng(100 + i), kind)
        genome.pointwise(ops)
        compare(genome, inputs, fn, what="pointwise case %d" % i)


@pytest.mark.parametrize("kind", KINDS)
def test_pointwise_fused_chain_equals_sequence(genome, orc, kind):
    G = type(genome)
    inputs = load(genome, np.random.default_rng(77), kind)
    genome.pointwise([G.op_addconst(-1.5), G.op_abs(), G.op_clip(0.5, 6.0), G.op_invert(2.0), G.op_binarize(-1.0)])
    compare(genome, inputs,
            lambda v: orc.binarize(orc.invert(orc.clip(orc.abs(orc.addconst(v, -1.5)), 0.5, 6.0), 2.0), -1.0),
            what="fused chain")


def test_invert_auto_mid(genome, orc):
    inputs = load(genome, np.random.default_rng(8), "real")
    allv = np.concatenate(list(inputs.values()))
    lo, hi, n = genome.minmax()
    assert (lo, hi, n) == (allv.min(), allv.max(), allv.size)
    genome.invert()
    mid = (allv.min() + allv.max()) / 2.0
    compare(genome, inputs, lambda v: orc.invert(v, mid), what="invert")


def test_count_non_integer(genome):
    """gdsp_count_non_integer: the check the host makes before it folds overlapping integer-valued
    intervals of add / subtract into one difference array (add.c:280-281 adds them one by one)"""
    rng = np.random.default_rng(81)
    inputs = load(genome, rng, "int")
    assert genome.count_non_integer() == 0
    want = 0
    for name, n in CHROMS:                           # a few of every kind of offender, some at tile edges
        v = inputs[name].copy()
        for pos, x in [(0, 0.5), (n - 1, np.nan), (n // 2, np.inf), (min(n - 1, 8191), -3.25), (min(n - 1, 8192), 2.0 ** 53),
                       (n // 3, -0.0), (n // 5, -7.0), (n // 7, 2.0 ** 52)]:
            v[pos] = x
        want += int(np.count_nonzero(~((np.abs(v) <= 2.0 ** 52) & (v == np.rint(v)))))
        genome.set_chrom(name, v)
    assert want > 0
    assert genome.count_non_integer() == want
    assert genome.count_non_integer(limit=6.0) == sum(
        int(np.count_nonzero(~((np.abs(genome.get_chrom(name)) <= 6.0) & (genome.get_chrom(name) == np.rint(genome.get_chrom(name))))))
        for name, _ in CHROMS)


def _sorted_disjoint(rng, genome, skip=()):
    seg, s, e, val = [], [], [], []
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        if name in skip:
            continue
        pos = int(rng.integers(0, 50))
        while pos < n:
            L = int(rng.integers(1, 300)); end = min(n, pos + L)
            seg.append(k); s.append(pos); e.append(end); val.append(float(rng.integers(1, 7)) / 2.0 * (-1) ** int(rng.integers(0, 2)))
            pos = end + int(rng.integers(0, 200))
    return np.array(seg, np.uint32), np.array(s, np.uint32), np.array(e, np.uint32), np.array(val)


@pytest.mark.parametrize("op", ["multiply", "divide", "masknot", "and", "mask", "or", "add", "subtract"])
def test_interval_table_ops(genome, orc, op):
    from genodsp_b200 import capi
    rng = np.random.default_rng(len(op))
    seg, s, e, val = _sorted_disjoint(rng, genome, skip=("chr3",))
    table = genome.interval_table(seg, s, e, val)
    inputs = load(genome, rng, "real")
    big = np.finfo(np.float64).max
    prog = {
        "multiply": [(capi.PW_IVL_MUL, 0.0, 0, 0, 0, table)],
        "divide": [(capi.PW_IVL_DIV, big, 0, 0, 0, table)],
        "masknot": [(capi.PW_IVL_SET_OUTSIDE, -2.0, 0, 0, 0, table)],
        "and": [(capi.PW_NONZERO_TO_ONE, 0.0), (capi.PW_IVL_SET_OUTSIDE, 0.0, 0, 0, 0, table)],
        "mask": [(capi.PW_IVL_SET, -3.0, 0, 0, 0, table)],
        "or": [(capi.PW_NONZERO_TO_ONE, 0.0), (capi.PW_IVL_SET, 1.0, 0, 0, 0, table)],
        "add": [(capi.PW_IVL_ADD, 0.0, 0, 0, 0, table)],
        "subtract": [(capi.PW_IVL_SUB, 0.0, 0, 0, 0, table)],
    }[op]
    genome.pointwise(prog)
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        sel = seg == k
        v = inputs[name].copy()
        if op == "multiply":
            orc.sorted_intervals(v, s[sel], e[sel], val[sel], 0, 0.0)
        elif op == "divide":
            orc.sorted_intervals(v, s[sel], e[sel], val[sel], 1, big)
        elif op == "masknot":
            orc.sorted_intervals(v, s[sel], e[sel], val[sel], 2, -2.0)
        elif op == "and":
            orc.sorted_intervals(orc.logical_prep(v), s[sel], e[sel], val[sel], 3, 0.0)
        elif op == "mask":
            orc.mask_intervals(v, s[sel], e[sel], -3.0)
        elif op == "or":
            orc.or_intervals(orc.logical_prep(v), s[sel], e[sel], val[sel])
        elif op == "add":
            orc.add_intervals(v, s[sel], e[sel], val[sel], 1.0)
        else:
            orc.add_intervals(v, s[sel], e[sel], val[sel], -1.0)
        got = genome.get_chrom(name)
        assert np.array_equal(bits(got), bits(v)), (op, name)
    table.close()


@pytest.mark.parametrize("op", ["minover", "maxover"])
@pytest.mark.parametrize("kind", ["int", "real"])
def test_over_intervals(genome, orc, op, kind):
    """minover / maxover: gdsp_ivl_arg_extrema + GDSP_PW_IVL_KEEP_AT vs the reference loop (ties by
    inset, then earliest), fill everywhere else incl. a chromosome the table never mentions"""
    rng = np.random.default_rng(5 + len(op))
    seg, s, e, val = _sorted_disjoint(rng, genome, skip=("chr3",))
    # a few long intervals (more than one warp stride) and single-cell ones
    table = genome.interval_table(seg, s, e, None)
    inputs = load(genome, rng, kind)
    if kind == "int":
        for name in inputs:
            inputs[name] = np.floor(inputs[name] / 3.0); genome.set_chrom(name, inputs[name])
    fill = -2.5
    (genome.minover if op == "minover" else genome.maxover)(table, fill)
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        sel = seg == k
        want = orc.over_intervals(inputs[name].copy(), s[sel], e[sel], op == "maxover", fill)
        got = genome.get_chrom(name)
        bad = np.nonzero(bits(got) != bits(want))[0]
        assert bad.size == 0, (op, name, bad[:5], got[bad[:5]], want[bad[:5]])
    table.close()


@pytest.mark.parametrize("op", ["minwith", "maxwith"])
def test_with_intervals(genome, orc, op):
    from genodsp_b200 import capi
    rng = np.random.default_rng(31)
    seg, s, e, val = _sorted_disjoint(rng, genome, skip=("chr2",))
    table = genome.interval_table(seg, s, e, val)
    inputs = load(genome, rng, "real")
    genome.pointwise([(capi.PW_IVL_MIN if op == "minwith" else capi.PW_IVL_MAX, 0.0, 0, 0, 0, table)])
    for k in range(genome.nseg):
        name, n = genome.chroms[genome.seg_chrom[k]]
        sel = seg == k
        want = orc.with_intervals(inputs[name].copy(), s[sel], e[sel], val[sel], op == "maxwith")
        assert np.array_equal(bits(genome.get_chrom(name)), bits(want)), (op, name)
    table.close()


@pytest.mark.parametrize("npoints", [1, 2, 23, 5000])
def test_map_values(genome, orc, npoints):
    rng = np.random.default_rng(npoints)
    vin = np.sort(rng.choice(np.arange(-40000, 40001), npoints, replace=False)) / 4096.0
    vout = rng.normal(0, 5, npoints)
    inputs = load(genome, rng, "real")
    for name, n in CHROMS:                 # exact hits and both clamped ends
        k = min(n, npoints)
        inputs[name][:k] = vin[:k]
        if n > k + 2:
            inputs[name][k] = vin[0] - 1; inputs[name][k + 1] = vin[-1] + 1
        genome.set_chrom(name, inputs[name])
    genome.map_values(vin, vout)
    compare(genome, inputs, lambda v: orc.map_values(v, vin, vout), what="map %d" % npoints)


@pytest.mark.parametrize("collapse", [True, False])
@pytest.mark.parametrize("show", [0, 1])
@pytest.mark.parametrize("kind", ["sparse", "int", "real"])
def test_runs(genome, orc, collapse, show, kind):
    inputs = load(genome, np.random.default_rng(40), kind)
    for name, n in CHROMS:
        if n > 100:
            v = inputs[name]; v[0:3] = 0.0; v[50:60] = 2.0; v[n - 4:] = 0.0
            genome.set_chrom(name, v)
    got = genome.runs(collapse, show, cap=1000)       # small cap: exercises the capacity retry
    for name, n in CHROMS:
        rs, re, rv = orc.runs(inputs[name], collapse, show)
        gs, ge, gv = got[name]
        assert np.array_equal(gs, rs) and np.array_equal(ge, re), (name, collapse, show)
        assert np.array_equal(bits(gv), bits(rv)), (name, collapse, show)


@pytest.mark.parametrize("kind", ["int", "sparse", "dyadic"])
@pytest.mark.parametrize("L", [1, 10, 100, 1000, 5000])
def test_clump(genome, orc, kind, L):
    T = {"int": 5.5, "sparse": 0.5, "dyadic": 0.25}[kind]
    inputs = load(genome, np.random.default_rng(L), kind)
    genome.clump(T, L)
    compare(genome, inputs, lambda v: orc.clump(v, T, L, True), what="clump L=%d" % L)
    inputs = load(genome, np.random.default_rng(L + 1), kind)
    genome.anticlump(T, L, one=3.0, zero=-1.0)
    compare(genome, inputs, lambda v: orc.clump(v, T, L, False, 3.0, -1.0), what="anticlump L=%d" % L)


@pytest.mark.parametrize("L", [2, 127, 128, 129, 1000, 4095, 4096, 4097])
def test_clump_long_runs_and_halo_edges(genome, orc, L):
    """minimum lengths either side of the 128-cell carry groups and of the 4096-cell limit of the
    path that rebuilds the prefix sums from group carries; plateaus that span many tiles (the run
    trimming carries cross tile and word boundaries), runs touching both chromosome ends"""
    rng = np.random.default_rng(L)
    inputs = load(genome, rng, "sparse")
    for name, n in CHROMS:
        v = inputs[name]
        if n > 60000:
            v[5000:5000 + 3 * 4096 + 17] = 2.0            # whole tiles above the threshold
            v[30000:30031] = 1.0; v[30031:30063] = 0.0; v[30063:30100] = 3.0
            v[n - 9000:] = 1.0                            # a run that ends with the chromosome
            v[:700] = 1.0                                 # ... and one that starts with it
            v[40960 - 1] = 1.0; v[40960] = 0.0; v[40961:40999] = 1.0
        genome.set_chrom(name, v)
    genome.clump(0.5, L)
    compare(genome, inputs, lambda v: orc.clump(v, 0.5, L, True), what="clump long runs L=%d" % L)
    for name, n in CHROMS:
        genome.set_chrom(name, inputs[name])
    genome.anticlump(1.5, L, one=7.0, zero=0.5)
    compare(genome, inputs, lambda v: orc.clump(v, 1.5, L, False, 7.0, 0.5), what="anticlump long runs L=%d" % L)


@pytest.mark.parametrize("L", [512, 1000, 3000, 4096])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_clump_mostly_below_threshold(genome, orc, L, seed):
    """long stretches without a qualifying cell (whole 4096-cell tiles): k_clump_mark settles those tiles from the
    group records without reading the signal.  The stretches sit beside humps of every size, so that all three
    outcomes occur: no valid end and nothing marked; no valid end in the tile but cells marked from a hump on its
    right (the early-out must fall through); valid ends inside a falling stretch because an old, deeper prefix
    minimum lies behind it (the record test must fail)"""
    rng = np.random.default_rng(100 * L + seed)
    inputs = {}
    for name, n in CHROMS:
        v = np.full(n, -1.0)
        pos = int(rng.integers(0, 3000))
        while pos < n:
            stretch = int(rng.choice([700, 5000, 9000, 20000, 45000]))
            pos += stretch
            hump = int(rng.choice([40, 600, 1500, 6000, 12000]))
            height = float(rng.choice([0.5, 1.0, 4.0, 16.0]))
            v[pos:pos + hump] = height
            pos += hump
        if seed == 2:
            v[: n // 3] -= 3.0                      # a deep early minimum: later falling stretches stay above it
            v[n // 3: n // 3 + 8000] = 12.0
        inputs[name] = v
        genome.set_chrom(name, v)
    genome.clump(0.0, L, one=2.0, zero=-3.0)
    compare(genome, inputs, lambda v: orc.clump(v, 0.0, L, True, 2.0, -3.0), what="clump mostly-below L=%d" % L)
    for name, n in CHROMS:
        genome.set_chrom(name, -inputs[name])
    genome.anticlump(0.0, L)
    compare(genome, {k: -v for k, v in inputs.items()}, lambda v: orc.clump(v, 0.0, L, False), what="anticlump mostly-above L=%d" % L)


def test_clump_stored_prefix_passes_on_short_lengths(genome, orc, monkeypatch):
    """the stored-prefix passes (normally only for minimum lengths above 4096) on short ones"""
    monkeypatch.setenv("GDSP_CLUMP_STORED", "1")
    for L in (3, 300):
        inputs = load(genome, np.random.default_rng(L), "int")
        genome.clump(5.5, L)
        compare(genome, inputs, lambda v: orc.clump(v, 5.5, L, True), what="stored-prefix clump L=%d" % L)


def test_clump_all_below_and_relative_length(genome, orc):
    inputs = load(genome, np.random.default_rng(0), "int")
    genome.clump(1000.0, 10)
    compare(genome, inputs, lambda v: orc.clump(v, 1000.0, 10, True), what="clump all below")
    inputs = load(genome, np.random.default_rng(1), "int")
    genome.clump(5.5, 0, relative_length=0.001)
    compare(genome, inputs, lambda v: orc.clump(v, 5.5, max(0, int(0.001 * v.size)), True), what="clump CL*0.001")


def _sorted_post_state(genome, inputs):
    names = [genome.chroms[i][0] for i in genome.order]
    allv = np.sort(np.concatenate([inputs[n] for n in names]))
    out, pos = {}, 0
    for n in names:
        out[n] = allv[pos:pos + inputs[n].size]; pos += inputs[n].size
    return out


@pytest.mark.parametrize("kind", KINDS)
def test_percentile_values_and_sorted_post_state(genome, orc, kind):
    inputs = load(genome, np.random.default_rng(17), kind)
    names = [genome.chroms[i][0] for i in genome.order]
    srt = orc.sort(np.concatenate([orc.percentile_collect(inputs[n]) for n in names]).copy())
    got = genome.percentile(1.0, 99.0, step=7.0, destructive=False)
    for p in range(1000, 99001, 7000):
        want = srt[orc.percentile_rank(srt.size, p)]
        name = "percentile%d" % (p // 1000)
        assert np.float64(got[name]).view(np.uint64) == np.float64(want).view(np.uint64), (kind, p)
    compare(genome, inputs, lambda v: v, what="percentile must not modify the signal")
    for p_str, p in (("99", 99000), ("99.5", 99500), ("12.345", 12345), ("100", 100000), ("0", 0)):
        got = genome.percentile(float(p_str), destructive=False)
        want = srt[-1] if p == 100000 else srt[orc.percentile_rank(srt.size, p)]
        assert got["percentile" + p_str] == want, (kind, p_str)
    genome.percentile(99.0, destructive=True)
    post = _sorted_post_state(genome, inputs)
    for name, n in CHROMS:
        assert np.array_equal(genome.get_chrom(name), post[name]), (kind, name)


@pytest.mark.parametrize("ties_above", [False, True])
@pytest.mark.parametrize("kind", KINDS + ["negzero", "nan"])
def test_percentile_then_binarize_without_sorting(genome, orc, kind, ties_above):
    """`percentile P = binarize --threshold=percentileP`: the binarized SORTED genome is a step function;
    gdsp_sorted_binarize writes it from a count (no sort).  Must equal sort-then-binarize bit for bit --
    also with -0.0/+0.0 around a zero threshold, and with NaNs (which fall back to the real sort)."""
    base = "sparse" if kind in ("negzero", "nan") else kind
    inputs = load(genome, np.random.default_rng(23), base)
    if kind != base:
        for name, n in CHROMS:
            v = inputs[name]
            v[::7] = -0.0; v[1::11] = -1.5
            if kind == "nan" and n > 50:
                v[13] = np.nan; v[n - 2] = -np.nan
            genome.set_chrom(name, v)
    launches0 = genome.launches
    got = genome.percentile(90.0, destructive=True)
    thr = 0.0 if kind == "negzero" else got["percentile90"]
    genome.binarize(thr, ties_above, one=2.0, zero=-1.0)
    fused = {name: genome.get_chrom(name) for name, _ in CHROMS}
    fused_launches = genome.launches - launches0
    # the explicit route: select, sort, threshold
    for name, n in CHROMS:
        genome.set_chrom(name, inputs[name])
    launches0 = genome.launches
    genome.percentile(90.0, destructive=False)
    genome.sort_genome()
    genome.binarize(thr, ties_above, one=2.0, zero=-1.0)
    for name, n in CHROMS:
        assert np.array_equal(bits(fused[name]), bits(genome.get_chrom(name))), (kind, name)
    if kind != "nan":
        assert fused_launches < genome.launches - launches0, "the fused route must not have sorted"
        post = _sorted_post_state(genome, inputs)
        for name, n in CHROMS:
            want = np.where((post[name] >= thr) if ties_above else (post[name] > thr), 2.0, -1.0)
            assert np.array_equal(bits(fused[name]), bits(want)), (kind, name)


@pytest.mark.parametrize("W", [2, 5, 100])
def test_percentile_window_and_range(genome, orc, W):
    inputs = load(genome, np.random.default_rng(W), "int")
    names = [genome.chroms[i][0] for i in genome.order]
    srt = orc.sort(np.concatenate([orc.percentile_collect(inputs[n], W, 2.0, 9.0) for n in names]).copy())
    got = genome.percentile(10.0, 90.0, step=20.0, window=W, mn=2.0, mx=9.0, destructive=False)
    assert genome.num_samples == srt.size
    for p in (10, 30, 50, 70, 90):
        assert got["percentile%d" % p] == srt[orc.percentile_rank(srt.size, p * 1000)], (W, p)


def test_percentile_many_and_heavy_ties(genome, orc):
    rng = np.random.default_rng(3)
    inputs = {}
    for name, n in CHROMS:
        v = np.where(rng.random(n) < 0.7, 4.0, rng.integers(0, 3, n).astype(np.float64))   # 70 % ties
        inputs[name] = v; genome.set_chrom(name, v)
    names = [genome.chroms[i][0] for i in genome.order]
    srt = np.sort(np.concatenate([inputs[n] for n in names]))
    got = genome.percentile(0.5, 99.5, step=0.5, destructive=False)           # 199 percentiles
    for p in range(500, 99501, 500):
        want = srt[orc.percentile_rank(srt.size, p)]
        key = "percentile%d" % (p // 1000) if p % 1000 == 0 else "percentile%.1f" % (p / 1000.0)
        assert got[key] == want, p


def test_sort_genome_real_values(genome):
    inputs = load(genome, np.random.default_rng(23), "real")
    genome.sort_genome()
    post = _sorted_post_state(genome, inputs)
    for name, n in CHROMS:
        assert np.array_equal(genome.get_chrom(name), post[name]), name


def _key_order(v):
    """gdsp_sort_genome's total order (f64_key): by value, -0.0 before +0.0"""
    b = np.ascontiguousarray(v, np.float64).view(np.uint64)
    key = np.where(b >> np.uint64(63) != 0, ~b, b | np.uint64(1 << 63))
    return v[np.argsort(key, kind="stable")]


@pytest.mark.parametrize("kind", ["real", "int", "sparse"])
def test_sort_one_chromosome_leaves_the_others_alone(genome, kind):
    """the per-chromosome sorts of the percentile passes (percentile.c:611-621): a layout that is not the
    front of the buffer must not spill its intermediate radix passes over its neighbours (real values
    take more than two digit passes)"""
    inputs = load(genome, np.random.default_rng(57), kind)
    for k in range(genome.nseg):
        genome.piece_sort(k)
        name = genome.chroms[genome.seg_chrom[k]][0]
        inputs[name] = _key_order(inputs[name])
        for other, _ in CHROMS:
            got = genome.get_chrom(other)
            assert np.array_equal(bits(got), bits(inputs[other])), (kind, "after sorting", name, "chromosome", other)


@pytest.mark.parametrize("kind", ["real", "int", "ties"])
def test_merge_exchange_equals_joint_sort(genome, kind):
    """gdsp_merge_exchange = combine_sorted_vectors (percentile.c:820-864): two sorted chromosomes, the first
    ends up with the smallest cells of both, bytes of a joint sort; lengths 1, 63, tile multiples and +-1"""
    rng = np.random.default_rng(91)
    vals = {}
    for name, n in CHROMS:
        if kind == "ties":
            v = rng.integers(-1, 3, n).astype(np.float64)
            v[rng.random(n) < 0.1] = -0.0
        else:
            v = signal(rng, n, kind) + (rng.integers(-2, 3) if kind == "int" else rng.normal())
        vals[name] = _key_order(v)
        genome.set_chrom(name, vals[name])
    seg_of = {genome.chroms[genome.seg_chrom[k]][0]: k for k in range(genome.nseg)}
    for c, d in [("chr1", "chr7"), ("chr5", "chr6"), ("chr6", "chr5"), ("chr3", "chr4"), ("chr4", "chr2"), ("chr7", "chr5"),
                 ("chr2", "chr1"), ("chr1", "chr7")]:
        both = _key_order(np.concatenate([vals[c], vals[d]]))
        moved = genome.merge_exchange(seg_of[c], seg_of[d])
        want_c, want_d = both[:vals[c].size], both[vals[c].size:]
        assert moved == _moved(vals[c], vals[d]), (kind, c, d)
        vals[c], vals[d] = want_c, want_d
        for name, _ in CHROMS:
            assert np.array_equal(bits(genome.get_chrom(name)), bits(vals[name])), (kind, c, d, name)


def _moved(cv, dv):
    """the reference's count: first k with not (d[k] < c[len-1-k]) in key order"""
    def key(v):
        b = np.ascontiguousarray(v, np.float64).view(np.uint64)
        return np.where(b >> np.uint64(63) != 0, ~b, b | np.uint64(1 << 63))
    n = min(cv.size, dv.size)
    if n == 0:
        return 0
    less = key(dv)[:n] < key(cv)[::-1][:n]
    stop = np.nonzero(~less)[0]
    return int(stop[0]) if stop.size else n


def test_smooth_many_tiles_per_cta(orc):
    """chromosomes of many strips / tiles, one shorter than every window"""
    from genodsp_b200.genome import Genome
    chroms = [("big", 2500000), ("mid", 1300000 + 7), ("tiny", 50)]
    g = Genome(chroms)
    rng = np.random.default_rng(99)
    ins = {n: signal(rng, l, "real" if n != "mid" else "int") for n, l in chroms}
    for W in (101, 11, 513):
        for n, v in ins.items():
            g.set_chrom(n, v)
        g.smooth(W)
        for n, v in ins.items():
            want = orc.smooth(v.copy(), W)
            got = g.get_chrom(n)
            bad = np.nonzero(bits(got) != bits(want))[0]
            assert bad.size == 0, (W, n, bad[:5])
    g.close()


def test_sort_genome_mostly_zeros(genome):
    """zero-majority shortcut of gdsp_sort_genome (negatives, -0.0 and +0.0 included)"""
    rng = np.random.default_rng(31)
    inputs = {}
    for name, n in CHROMS:
        v = np.where(rng.random(n) < 0.93, 0.0, rng.normal(0, 3, n))
        if n > 10:
            v[3] = -0.0
        inputs[name] = v; genome.set_chrom(name, v)
    genome.sort_genome()
    post = _sorted_post_state(genome, inputs)
    for name, n in CHROMS:
        got = genome.get_chrom(name)
        assert np.array_equal(got, post[name]), name      # (-0.0 == +0.0: their relative order is unspecified in the reference too)


@pytest.mark.parametrize("kind", KINDS + ["nan"])
def test_percentiles_ranked_counts(genome, orc, kind):
    """gdsp_percentiles_ranked: for every reported value the exact number of cells below / equal to it
    (what lets `binarize` after `percentile` write its step function without a counting pass)"""
    base = "real" if kind == "nan" else kind
    inputs = load(genome, np.random.default_rng(31), base)
    if kind == "nan":
        v = inputs["chr1"]; v[5] = np.nan; v[777] = -np.nan
        genome.set_chrom("chr1", v)
    allv = np.concatenate([inputs[n] for n, _ in CHROMS])
    got = genome.percentile(5.0, 95.0, step=15.0, destructive=True)
    if kind == "nan":
        assert genome._sorted_known == {}, "NaN present: the step-function shortcut must be off"
        return
    assert len(genome._sorted_known) == len(set(got.values()))
    for val, (below, equal, _neg) in genome._sorted_known.items():
        assert below == int((allv < val).sum()) + (int(((allv == 0) & np.signbit(allv)).sum()) if val == 0 and not np.signbit(val) else 0) \
            or below == int((allv < val).sum()), (kind, val, below)
        if val != 0:
            assert equal == int((allv == val).sum()), (kind, val, equal)


@pytest.mark.parametrize("kind", ["int", "real"])
@pytest.mark.parametrize("args", [(7, None, None), (1, 2.0, None), (1, None, 4.0), (3, 1.5, 5.0), (1000, None, None),
                                  (1, None, None), (2, 1e9, None)])
def test_percentile_collect_permutation(genome, kind, args):
    """gdsp_percentile_collect == the reference's collect swaps (percentile.c:547-580) in closed form
    (tests/percentile_model.py, pinned to the reference by tests/test_oracle_percentile_state.py)"""
    import percentile_model as pm
    stride, mn, mx = args
    inputs = load(genome, np.random.default_rng(stride + 3), kind)
    names = [genome.chroms[i][0] for i in genome.order]
    lengths = [inputs[n].size for n in names]
    cat = np.concatenate([inputs[n] for n in names])
    lo = -np.finfo(np.float64).max if mn is None else mn
    hi = np.finfo(np.float64).max if mx is None else mx
    want, n_want = pm.collect(lengths, cat, stride, lo, hi)
    n = genome.percentile_collect(stride, lo, hi)
    assert n == n_want
    got = np.concatenate([genome.get_chrom(nm) for nm in names])
    bad = np.nonzero(bits(got) != bits(want))[0]
    assert bad.size == 0, (kind, args, n, bad[:8], got[bad[:8]], want[bad[:8]])


@pytest.mark.parametrize("path", ["host", "dev"])
@pytest.mark.parametrize("buckets", ["fixed", "exact"])
@pytest.mark.parametrize("shape", ["uniform", "hot_tiles", "pileup"])
def test_accumulate_fixed_capacity_buckets(orc, monkeypatch, path, buckets, shape):
    """round 2: reads go straight into fixed-capacity tile buckets (no count pass, no offset prefix); a tile
    that receives more than its capacity spills into a small overflow list that only the overflowing tiles
    scan, and a pile-up beyond BIN_OVF_MAX records sends the whole input down the exact-size path.  All three
    regimes, and the exact-size path forced by GDSP_ACCUMULATE_EXACT_BUCKETS, must give the reference loop's
    depth (genodsp.c:1325-1329) bit for bit."""
    import torch
    from genodsp_b200.genome import Genome
    if buckets == "exact":
        monkeypatch.setenv("GDSP_ACCUMULATE_EXACT_BUCKETS", "1")
    chroms = [("chrA", 3_000_001), ("chrB", 1_000_000), ("chrC", 8192)]
    g = Genome(chroms)
    try:
        rng = np.random.default_rng({"uniform": 1, "hot_tiles": 2, "pileup": 3}[shape])
        m = 400_000
        seg = rng.integers(0, 2, m).astype(np.uint32)
        lens = np.array([g.segs[k][5] for k in seg])
        if shape == "uniform":
            start = (rng.random(m) * (lens - 200)).astype(np.uint32)
        elif shape == "hot_tiles":          # a tenth of the reads in three narrow regions: a few tiles at 10-30x the mean
            start = (rng.random(m) * (lens - 200)).astype(np.uint32)
            hot = rng.random(m) < 0.1
            start[hot] = (rng.choice([100_000, 500_000, 777_777], int(hot.sum())) + rng.integers(0, 3000, int(hot.sum()))).astype(np.uint32)
        else:                               # half of all reads inside one tile: far more than the overflow list holds
            start = (rng.random(m) * (lens - 200)).astype(np.uint32)
            hot = rng.random(m) < 0.5
            start[hot] = (250_000 + rng.integers(0, 5000, int(hot.sum()))).astype(np.uint32)
        end = (start + rng.integers(1, 200, m)).astype(np.uint32)
        g.fill(3.0)
        if path == "host":
            g.accumulate(seg, start, end)
        else:
            t = [torch.from_numpy(a.astype(np.int64)).to(g.device).to(torch.int32) for a in (seg, start, end)]
            g.accumulate(t[0], t[1], t[2], host=False)
        for k in range(g.nseg):
            name, n = g.chroms[g.seg_chrom[k]]
            sel = seg == k
            want = orc.accumulate(np.zeros(n), start[sel], end[sel], None)
            got = g.get_chrom(name)
            bad = np.nonzero(bits(got) != bits(want))[0]
            assert bad.size == 0, (shape, buckets, path, name, bad[:5], got[bad[:5]], want[bad[:5]])
        # and on top of an existing signal
        before = {name: g.get_chrom(name) for name, _ in chroms}
        if path == "host":
            g.accumulate(seg, start, end, add=True)
        else:
            g.accumulate(t[0], t[1], t[2], host=False, add=True)
        for name, n in chroms:
            assert np.array_equal(g.get_chrom(name), 2 * before[name]), (shape, buckets, path, name)
    finally:
        g.close()


@pytest.mark.parametrize("zeros", ["plus", "mixed"])
def test_percentile_window_that_starts_on_a_zero_plateau(genome, orc, zeros):
    """the cfg3 shape (`sum --window=100` leaves 99 % zeros): the percentile-99 window starts ON the zero plateau,
    so the counting pass runs with a zero bound and nearly every cell tied with it -- with +0.0 only, and with
    -0.0 mixed in (equal to DSETP, different keys)"""
    rng = np.random.default_rng(5 if zeros == "plus" else 6)
    inputs = {}
    for name, n in CHROMS:
        v = np.zeros(n)
        k = max(1, n // 100)
        v[rng.choice(n, k, replace=False)] = rng.integers(1, 2000, k) / 100.0
        if zeros == "mixed":
            z = np.flatnonzero(v == 0)
            v[z[::3]] = -0.0
        inputs[name] = v
        genome.set_chrom(name, v)
    names = [genome.chroms[i][0] for i in genome.order]
    srt = orc.sort(np.concatenate([inputs[n] for n in names]).copy())
    for p_str, p in (("98.5", 98500), ("99", 99000), ("99.2", 99200), ("50", 50000), ("99.9", 99900)):
        got = genome.percentile(float(p_str), destructive=False)["percentile" + p_str]
        want = srt[orc.percentile_rank(srt.size, p)]
        assert got == want, (zeros, p_str, got, want)      # (inside a tie of -0.0 / +0.0 the reference's qsort order is unspecified)
    got = genome.percentile(99.0, destructive=True)["percentile99"]
    genome.binarize(got)
    allv = np.concatenate([inputs[n] for n in names])
    want_sorted = np.where(srt > got, 1.0, 0.0)
    cat = np.concatenate([genome.get_chrom(n) for n in names])
    assert np.array_equal(cat, want_sorted)
