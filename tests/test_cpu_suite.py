"""CPU-only tests (run with -m "not gpu"): the oracle against the golden fixtures produced by the
reference binary, the host-side helpers, the slab partition with a world_size-2 gloo exchange, and the
C-ABI library's exported symbols (no compute calls: there is no GPU here)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from checkers import Oracle, ROOT
from oracle_pipeline import Pipeline

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = json.load(open(os.path.join(GOLD, "cases.json")))


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_golden(case):
    p = Pipeline(GOLD)
    got = p.run(CASES[case]["args"], CASES[case]["stdin"])
    want = open(os.path.join(GOLD, case + ".out")).read()
    if got != want:
        g, w = got.split("\n"), want.split("\n")
        for i, (a, b) in enumerate(zip(g, w)):
            assert a == b, "%s: line %d: oracle %r, reference %r" % (case, i + 1, a, b)
        assert len(g) == len(w)
    err = open(os.path.join(GOLD, case + ".err")).read().strip().split("\n")
    assert [e for e in err if e] == p.stderr, case


def test_hann_taps_match_oracle():
    from genodsp_b200.genome import hann_taps
    orc = Oracle()
    for W in (3, 11, 101, 1001):
        assert np.array_equal(hann_taps(W).view(np.uint64), orc.hann_taps(W).view(np.uint64))


def test_percentile_rank_and_name():
    from genodsp_b200.genome import percentile_name, percentile_rank
    orc = Oracle()
    for n in (1, 7, 1000, 3088269832, 4294967295):
        for p in (1, 500, 12345, 50000, 99000, 99999):
            assert percentile_rank(n, p) == orc.percentile_rank(n, p)
    assert percentile_name(99000) == "percentile99"
    assert percentile_name(99500) == "percentile99.5"
    assert percentile_name(99720) == "percentile99.72"
    assert percentile_name(12345) == "percentile12.345"


def test_library_exports_every_declared_symbol():
    """include/gdsp_b200.h is the C-ABI contract: every function it declares must be exported by the built
    library and bound (with the same name) by the ctypes mirror."""
    from genodsp_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "gdsp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gdsp_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    lib = C.CDLL(capi.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), "libgdsp_b200.so does not export %s" % n
        assert n in capi.SIGNATURES, "capi.py does not bind %s" % n
    assert set(capi.SIGNATURES) <= names
    capi.load()


def test_no_cpu_fallback_without_device():
    """without a CUDA device the library must refuse, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genodsp_b200 import capi
    lib = capi.load()
    ctx = C.c_void_p()
    assert lib.gdsp_ctx_create(0, None, C.byref(ctx)) == -5
    assert b"no CPU path" in lib.gdsp_last_error()
    with pytest.raises(capi.GdspError):
        from genodsp_b200.genome import Genome
        Genome([("a", 10)])


def test_product_never_touches_the_oracle():
    """only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke may use oracle/"""
    for d, _, files in os.walk(os.path.join(ROOT, "genodsp_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "gdsp_oracle" not in txt and "oracle/" not in txt and "checkers" not in txt, os.path.join(d, f)


# ---------------------------------------------------------------------------------------- slabs
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 156040895, 145138636,
        138394717, 135086622, 133797422, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
        64444167, 58617616, 57227415, 50818468, 46709983]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slab_partition_covers_genome_once(world):
    from genodsp_b200 import slab
    owned = {}
    sizes = []
    for r in range(world):
        segs, cells = slab.partition(HG38, world, r, 50)
        sizes.append(sum(hi - lo for _, lo, hi, _, _, _ in segs))
        prev_end = 0
        for si, lo, hi, dlo, dhi, pos0 in segs:
            assert lo % 64 == 0 and dlo >= prev_end and dlo <= lo and hi <= dhi <= cells
            prev_end = dhi
            owned.setdefault(si, []).append((pos0, pos0 + hi - lo))
    assert sum(sizes) == sum(HG38) and max(sizes) - min(sizes) <= 1
    for si, pieces in owned.items():
        pieces.sort()
        assert pieces[0][0] == 0 and pieces[-1][1] == HG38[si]
        for a, b in zip(pieces, pieces[1:]):
            assert a[1] == b[0]


def _gloo_worker(rank, world, port, lengths, halo, out):
    import torch
    import torch.distributed as dist
    from genodsp_b200 import slab
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    segs, cells = slab.partition(lengths, world, rank, halo)
    buf = torch.full((cells,), -1.0, dtype=torch.float64)
    g0 = [0]
    for l in lengths:
        g0.append(g0[-1] + l)
    for si, lo, hi, dlo, dhi, pos0 in segs:       # cell value = its genome coordinate
        buf[lo:hi] = torch.arange(g0[si] + pos0, g0[si] + pos0 + (hi - lo), dtype=torch.float64)
    slab.exchange_halos(buf, slab.halo_plan(lengths, world, rank, halo), dist)
    ok = True
    for si, lo, hi, dlo, dhi, pos0 in segs:       # halo cells must hold the neighbour's coordinates
        want = torch.arange(g0[si] + pos0 - (lo - dlo), g0[si] + pos0 + (hi - lo) + (dhi - hi), dtype=torch.float64)
        ok = ok and bool(torch.equal(buf[dlo:dhi], want))
    out[rank] = ok
    dist.destroy_process_group()


def test_halo_exchange_world2_gloo():
    import torch.multiprocessing as mp
    lengths = [5000, 3000, 1200, 300]
    for world in (2, 3):
        mgr = mp.Manager()
        out = mgr.dict()
        port = 29500 + world + (os.getpid() % 1000)
        mp.spawn(_gloo_worker, args=(world, port, lengths, 37, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world)), dict(out)


def test_halo_plan_send_matches_peer_recv():
    """a cut within `halo` cells of a chromosome boundary gives the two sides different halos;
    what one rank sends must be what its peer posts as a receive (both directions)."""
    from genodsp_b200 import slab
    cases = [([990, 1010], 2, 50), ([1010, 990], 2, 50), ([5000, 3000, 1200, 300], 3, 37), ([1000] * 7, 4, 120),
             ([2000, 30, 1970], 2, 64), ([300, 300, 300, 300], 3, 80)]
    for lengths, world, halo in cases:
        plans = [slab.halo_plan(lengths, world, r, halo) for r in range(world)]
        for r, plan in enumerate(plans):
            for peer, s_lo, s_hi, r_lo, r_hi in plan:
                theirs = [p for p in plans[peer] if p[0] == r]
                assert len(theirs) == 1, (lengths, world, r, peer)
                assert s_hi - s_lo == theirs[0][4] - theirs[0][3], (lengths, world, halo, r, peer)
                assert r_hi - r_lo == theirs[0][2] - theirs[0][1], (lengths, world, halo, r, peer)
    # a piece shorter than the halo its neighbour needs is refused, not silently mis-sized
    with pytest.raises(ValueError):
        for r in range(3):
            slab.halo_plan([3000], 3, r, 1500)


def test_halo_exchange_cut_near_chromosome_boundary_gloo():
    import torch.multiprocessing as mp
    for lengths in ([990, 1010], [1010, 990]):
        mgr = mp.Manager()
        out = mgr.dict()
        port = 29650 + (os.getpid() % 1000) + len(out) + lengths[0] % 7
        mp.spawn(_gloo_worker, args=(2, port, lengths, 50, out), nprocs=2, join=True)
        assert all(out[r] for r in range(2)), (lengths, dict(out))


# ----------------------------------------------------------------------------- slab-level operators
def _slab_signal(chroms, seed):
    rng = np.random.default_rng(seed)
    sig = {}
    for name, n in chroms:
        v = np.repeat(rng.integers(0, 6, n // 7 + 1), 7)[:n].astype(np.float64)      # runs of 7 equal cells
        sig[name] = v
    return sig


def _slab_expected(chroms, sig, plist):
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    allv = np.sort(np.concatenate([sig[chroms[i][0]] for i in order]))
    n = allv.size
    from genodsp_b200 import slab
    want_p = [float(allv[slab._pct_rank(n, p)]) for p in plist]
    want_c = {name: np.cumsum(sig[name]) for name, _ in chroms}
    want_r = {}
    for name, _ in chroms:
        v = sig[name]
        head = np.concatenate([[True], v[1:] != v[:-1]])
        st = np.nonzero(head)[0]; en = np.concatenate([st[1:], [v.size]]); val = v[st]
        keep = val != 0
        want_r[name] = (st[keep], en[keep], val[keep])
    return want_p, want_c, want_r


def _slab_ops_check(parts, gather, chroms, sig, plist):
    from genodsp_b200 import slab
    want_p, want_c, want_r = _slab_expected(chroms, sig, plist)
    got_p, n = slab.slab_percentiles(parts, gather, plist, sample_per_rank=512)
    assert n == sum(l for _, l in chroms) and got_p == want_p, (got_p, want_p)
    got_p, n, ranks, nan = slab.slab_percentiles(parts, gather, plist, sample_per_rank=512, ranked=True)
    allv = np.concatenate([sig[name] for name, _ in chroms])
    assert got_p == want_p and nan == 0
    for v, (below, equal) in zip(got_p, ranks):
        assert below == int((allv < v).sum()) and equal == int((allv == v).sum()), (v, below, equal)
    lo, hi, cnt = slab.slab_minmax(parts, gather)
    assert (lo, hi, cnt) == (min(v.min() for v in sig.values()), max(v.max() for v in sig.values()), n)
    runs = slab.slab_runs(parts, gather)
    for name, _ in chroms:
        if want_r[name][0].size == 0:
            continue
        for a, b in zip(runs[name], want_r[name]):
            assert np.array_equal(np.asarray(a, np.float64), np.asarray(b, np.float64)), name
    slab.slab_cumulativesum(parts, gather)
    for g in parts:
        for k in range(g.nseg):
            name = g.chroms[g.seg_chrom[k]][0]
            pos0 = g.segs[k][4]
            assert np.array_equal(g.piece[k], want_c[name][pos0:pos0 + g.piece[k].size]), name
    return True


def _host_parts(chroms, world, ranks, sig):
    from genodsp_b200 import slab
    from host_genome import HostGenome
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    lengths = [chroms[i][1] for i in order]
    parts = []
    for r in ranks:
        segs_s, _ = slab.partition(lengths, world, r, 0)
        parts.append(HostGenome(chroms, [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s], sig))
    return parts


SLAB_CHROMS = [("chrA", 9001), ("chrB", 20000), ("chrC", 313), ("chrD", 4500)]
SLAB_PCTS = [0, 1000, 50000, 99000, 99900, 100000]


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_slab_operators_virtual_ranks(world):
    """cross-rank logic of slab_percentiles / slab_cumulativesum / slab_runs / slab_minmax with every
    rank's piece held by a numpy stand-in (tests/host_genome.py)"""
    from genodsp_b200 import slab
    sig = _slab_signal(SLAB_CHROMS, 5)
    parts = _host_parts(SLAB_CHROMS, world, range(world), sig)
    assert _slab_ops_check(parts, slab.virtual_gather, SLAB_CHROMS, sig, SLAB_PCTS)


def _slab_gloo_worker(rank, world, port, out):
    import sys
    import torch.distributed as dist
    from genodsp_b200 import slab
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    sig = _slab_signal(SLAB_CHROMS, 5)
    parts = _host_parts(SLAB_CHROMS, world, [rank], sig)
    try:
        import torch
        out[rank] = _slab_ops_check(parts, slab.DistComm(dist, torch.device("cpu")), SLAB_CHROMS, sig, SLAB_PCTS)
    finally:
        dist.destroy_process_group()


def test_slab_operators_world2_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_slab_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    assert all(out[r] for r in range(2)), dict(out)


def test_merge_runs_chain_across_several_cuts():
    from genodsp_b200 import slab
    u = lambda *a: np.array(a, np.uint32)
    chunks = [(0, (u(0, 10), u(5, 20), np.array([1., 2.]))), (20, (u(20), u(40), np.array([2.]))),
              (40, (u(40, 50), u(45, 51), np.array([2., 3.])))]
    s, e, v = slab.merge_runs(chunks)
    assert s.tolist() == [0, 10, 50] and e.tolist() == [5, 45, 51] and v.tolist() == [1., 2., 3.]
    s, e, v = slab.merge_runs(chunks, collapse=False)
    assert s.size == 5


# ----------------------------------------------------------------------------- host CLI without a GPU
OURS_BIN = os.path.join(ROOT, "genodsp_b200", "bin", "genodsp")


def _run_cli(binary, args, stdin=b"chr1 1 5\n", cwd=None):
    import subprocess
    return subprocess.run([binary] + args, input=stdin, capture_output=True, cwd=cwd, timeout=60)


@pytest.mark.skipif(not os.path.exists(OURS_BIN), reason="host CLI not built")
def test_cli_parser_errors_match_reference_without_gpu(tmp_path):
    """command-line parsing happens before the device is opened: wrong arguments produce the reference's
    first message line and exit status (the usage text that follows is worded independently)"""
    from checkers import REF_BIN, have_ref
    (tmp_path / "g.chroms").write_text("chr1 100\nchr2 50\n")
    cases = [
        ["--chromosomes=g.chroms", "=", "binarize", "--bogus"],
        ["--chromosomes=g.chroms", "=", "nosuchoperator"],
        ["--chromosomes=g.chroms", "=", "smooth", "--window=abc"],
        ["--chromosomes=g.chroms", "=", "multiply"],
        ["--chromosomes=g.chroms", "=", "minover"],
        ["--chromosomes=g.chroms", "=", "map"],
        ["--chromosomes=nosuchfile"],
        ["--value=2"],
    ]
    for args in cases:
        ours = _run_cli(OURS_BIN, args, cwd=tmp_path)
        assert ours.returncode != 0, args
        if have_ref():
            ref = _run_cli(REF_BIN, args, cwd=tmp_path)
            assert ref.returncode == ours.returncode, (args, ref.returncode, ours.returncode)
            first = lambda b: b.decode(errors="replace").splitlines()[0] if b.strip() else ""
            assert first(ours.stderr) == first(ref.stderr), (args, ours.stderr[:200], ref.stderr[:200])


@pytest.mark.skipif(not os.path.exists(OURS_BIN), reason="host CLI not built")
def test_cli_refuses_to_run_without_a_device(tmp_path):
    """no CPU fallback: with no CUDA device the program stops with a message and a failure status"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    (tmp_path / "g.chroms").write_text("chr1 100\n")
    p = _run_cli(OURS_BIN, ["--chromosomes=g.chroms", "=", "binarize"], cwd=tmp_path)
    assert p.returncode != 0 and b"no CUDA device" in p.stderr and p.stdout == b""


# ----------------------------------------------------------------------------- parallel tokenizer (no GPU needed)
def _parse_only(args, text, threads, cwd):
    import subprocess
    env = dict(os.environ, GENODSP_PARSE_ONLY="1", GENODSP_THREADS=str(threads))
    return subprocess.run([OURS_BIN] + args, input=text, capture_output=True, cwd=cwd, env=env, timeout=120)


@pytest.mark.skipif(not os.path.exists(OURS_BIN), reason="host CLI not built")
def test_parallel_tokenizer_equals_line_at_a_time_reader(tmp_path):
    """the block tokenizer on T threads (gd_core.c:read_intervals_parallel) must accept, reject and number
    lines exactly as the one-line-at-a-time read_interval does (reference grammar genodsp.c:1384-1534):
    same intervals in the same order, same first error message and exit status"""
    rng = np.random.default_rng(11)
    (tmp_path / "g.chroms").write_text("chr1 2000000\nchr2 900000\nchrX 50\n")
    lines = ["track name=x", "# comment", "", "   ", "\t"]
    for _ in range(200000):
        c = ["chr1", "chr2", "chrUn", "chrX"][int(rng.integers(0, 4))]
        L = {"chr1": 2000000, "chr2": 900000, "chrUn": 1000, "chrX": 50}[c]
        a = int(rng.integers(0, L))
        b = min(L, a + int(rng.integers(0, 300)))
        sep = ["\t", " ", "  \t "][int(rng.integers(0, 3))]
        v = ["1", "2.5", "-0.125", "1e3", "+4", ".5", "7"][int(rng.integers(0, 7))]
        lines.append(sep.join([c, str(a), str(b), v, "extra"]) + ("  " if rng.random() < 0.1 else ""))
        if rng.random() < 0.001:
            lines.append("# interleaved comment")
    good = ("\n".join(lines) + "\n").encode()
    C = ["--chromosomes=g.chroms"]
    for args in (C, C + ["--novalue"], C + ["--value=5"], C + ["--origin=one"]):
        if "--origin=one" in args:
            text = good.replace(b"\t0\t", b"\t1\t").replace(b" 0 ", b" 1 ")
        else:
            text = good
        one = _parse_only(args, text, 1, tmp_path)
        many = _parse_only(args, text, 7, tmp_path)
        assert one.returncode == many.returncode, (args, one.stderr[-300:], many.stderr[-300:])
        assert one.stdout == many.stdout and one.stderr == many.stderr, (args, one.stdout, many.stdout, many.stderr[-300:])
    assert b"intervals=" in _parse_only(C, good, 7, tmp_path).stdout
    # no trailing newline on the last line; a value field the fast path declines ("nan", "inf", hex) is replayed
    odd = good[:-1] + b"\nchr1\t5\t9\tnan\nchr1\t5\t9\t0x10\nchr2 1 2 inf"
    a, b = _parse_only(C, odd, 1, tmp_path), _parse_only(C, odd, 5, tmp_path)
    assert a.returncode == b.returncode == 0 and a.stdout == b.stdout, (a.stdout, b.stdout, b.stderr[-300:])
    # every kind of malformed line: same message (with its line number) and status from both readers
    bad_lines = [b" chr1 5 9 1", b"chr1", b"chr1 5", b"chr1 5 9", b"chr1 x 9 1", b"chr1 5 y 1", b"chr1 5 3000000 1",
                 b"chr1 5 9 zzz", b"chr1 99999999999 9 1", b"chr1 " + b"9" * 1200 + b" 5 1", b"chrX 10 51 1", b"chr1 -5 9 1"]
    for pos in (10, len(lines) // 2, len(lines) - 3):
        for bad in bad_lines:
            text = b"\n".join([l.encode() for l in lines[:pos]] + [bad] + [l.encode() for l in lines[pos:]]) + b"\n"
            a, b = _parse_only(C, text, 1, tmp_path), _parse_only(C, text, 6, tmp_path)
            assert a.returncode == b.returncode, (bad, pos, a.stderr[-200:], b.stderr[-200:])
            assert a.stderr == b.stderr and a.stdout == b.stdout, (bad, pos, a.stderr[-200:], b.stderr[-200:])
            if pos != 10:
                break


@pytest.mark.skipif(not os.path.exists(OURS_BIN), reason="host CLI not built")
def test_parallel_tokenizer_errors_match_the_reference(tmp_path):
    """the first malformed line is reported with the reference's own words and line number"""
    from checkers import REF_BIN, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    import subprocess
    (tmp_path / "g.chroms").write_text("chr1 2000000\n")
    body = "".join("chr1\t%d\t%d\n" % (i, i + 10) for i in range(0, 300000, 3))
    for bad in ("chr1 7\n", " chr1 1 2\n", "chr1 5 3000000\n", "chr1 " + "9" * 1100 + " 5\n"):
        text = (body + bad + body).encode()
        ref = subprocess.run([REF_BIN, "--chromosomes=g.chroms", "--novalue"], input=text, capture_output=True, cwd=tmp_path)
        ours = _parse_only(["--chromosomes=g.chroms", "--novalue"], text, 8, tmp_path)
        assert ref.returncode != 0 and ours.returncode == ref.returncode, (bad[:30], ours.stderr[-200:])
        assert ours.stderr == ref.stderr, (bad[:30], ref.stderr[-200:], ours.stderr[-200:])
