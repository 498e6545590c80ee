#!/usr/bin/env python
"""Per-stage throughput of the operator kernels on an hg38-shaped genome (or hg38/--scale).

Every stage is timed with CUDA events on the stream the kernels run on, after warm-up, on a
signal much larger than L2.  Prints one JSON object: stage -> ms, Gbp/s, algorithmic GB/s and the
fraction of the measured HBM peak (MEASURED_PEAKS.json)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--stages", default="all")
    args = ap.parse_args()
    import torch
    from genodsp_b200.genome import Genome
    chroms = bench.scaled_genome(args.scale)
    g = Genome(chroms)
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
    N = g.cells
    peak, _ = bench.measured_hbm_peak()
    G = type(g)
    depth = None

    def reset():
        g.sig.copy_(depth)

    def timed(fn, setup=None):
        best = None
        for r in range(args.reps + 1):
            if setup:
                setup()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            if r > 0:
                best = ms if best is None else min(best, ms)
        return best

    out = {}

    def rec(name, ms, bytes_per_bp, extra_bytes=0):
        gbs = (bytes_per_bp * N + extra_bytes) / (ms / 1e3) / 1e9
        out[name] = {"ms": round(ms, 3), "gbp_s": round(N / (ms / 1e3) / 1e9, 2), "alg_bytes_per_bp": bytes_per_bp,
                     "achieved_gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 3)}
        print(name, out[name], flush=True)

    want = None if args.stages == "all" else set(args.stages.split(","))

    def on(name):
        return want is None or name in want

    ms = timed(lambda: g.accumulate(seg, st, en, host=False))
    rec("depth_accumulate", ms, 16, 28 * int(seg.shape[0]))
    depth = g.sig.clone()
    if on("slidingsum"):
        rec("slidingsum_w101", timed(lambda: g.slidingsum(101), reset), 16)
    if on("sum"):
        rec("sum_w100", timed(lambda: g.sum(100, denom=100.0), reset), 16)
    if on("smooth"):
        rec("smooth_w101", timed(lambda: g.smooth(101), reset), 16)
    if on("localmax"):
        rec("localmax_n11", timed(lambda: g.localmax(11), reset), 16)
    if on("bestmax"):
        rec("bestmax_w101", timed(lambda: g.bestmax(101), reset), 16)
    if on("binarize"):
        rec("binarize", timed(lambda: g.binarize(6.0), reset), 16)
        rec("pointwise_chain5", timed(lambda: g.pointwise([G.op_addconst(-1.5), G.op_abs(), G.op_clip(0.5, 6.0), G.op_invert(2.0),
                                                           G.op_binarize(-1.0)]), reset), 16)
    if on("ivl"):
        # cfg5's second track: sorted disjoint intervals covering ~50 %, values k/1024; the whole chain
        # add B = multiply B = mask B = and B = binarize is ONE launch
        import numpy as np
        from genodsp_b200 import capi
        rng = np.random.default_rng(99)
        bs, bstart, bend, bval = [], [], [], []
        for k, (lo, hi, dlo, dhi, pos0, clen) in enumerate(g.segs):
            n = hi - lo
            m = max(1, n // 2000)
            cuts = np.sort(rng.choice(np.arange(pos0, pos0 + n, dtype=np.int64), size=min(2 * m, n), replace=False))
            a, b = cuts[0::2], cuts[1::2]
            m2 = min(a.size, b.size)
            bs.append(np.full(m2, k, np.uint32)); bstart.append(a[:m2].astype(np.uint32)); bend.append(b[:m2].astype(np.uint32))
            bval.append(rng.integers(1, 2048, m2) / 1024.0)
        tableB = g.interval_table(np.concatenate(bs), np.concatenate(bstart), np.concatenate(bend), np.concatenate(bval))
        chain5 = [(capi.PW_IVL_ADD, 0.0, 0, 0, 0, tableB), (capi.PW_IVL_MUL, 0.0, 0, 0, 0, tableB),
                  (capi.PW_IVL_SET, 0.0, 0, 0, 0, tableB), (capi.PW_NONZERO_TO_ONE, 0.0),
                  (capi.PW_IVL_SET_OUTSIDE, 0.0, 0, 0, 0, tableB), G.op_binarize(0.5)]
        rec("ivl_chain_cfg5", timed(lambda: g.pointwise(chain5), reset), 16)
        rec("ivl_multiply_only", timed(lambda: g.pointwise(chain5[1:2]), reset), 16)
        tableB.close()
    if on("cumulativesum"):
        rec("cumulativesum", timed(lambda: g.cumulativesum(), reset), 16)
    if on("percentile"):
        rec("percentile99_select", timed(lambda: g.percentile(99.0, destructive=False), reset), 8)
        rec("percentile_1to99by1_select", timed(lambda: g.percentile(1.0, 99.0, 1.0, destructive=False), reset), 8)
    if on("sort"):
        rec("sort_genome_depth", timed(lambda: g.sort_genome(), reset), 16)
        def smooth_setup():
            reset(); g.smooth(101)
        rec("percentile99_select_smoothed", timed(lambda: g.percentile(99.0, destructive=False), smooth_setup), 8)
        rec("sort_genome_smoothed", timed(lambda: g.sort_genome(), smooth_setup), 16)
    if on("bubble"):
        # the post-state of `percentile 50` with every position qualifying (percentile.c:611-651): every chromosome
        # sorted on its own, then chromosome c = 0..K takes the smallest cells of {c, d} for every later d
        # (gdsp_merge_exchange per step); K = the chromosome that holds the median rank
        import time
        lens = [hi - lo for (lo, hi, *_r) in g.segs]
        acc, K = 0, len(lens) - 1
        for i, l in enumerate(lens):
            acc += l
            if N // 2 < acc:
                K = i
                break
        steps = moved = 0

        def passes():
            nonlocal steps, moved
            steps = moved = 0
            for k in range(g.nseg):
                g.piece_sort(k)
            for c in range(K + 1):
                for d in range(c + 1, g.nseg):
                    moved += g.merge_exchange(c, d)
                    steps += 1
        best = None
        for r in range(2):
            reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
            passes(); torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        srt = True
        for k in range(K + 1):                       # chromosomes 0..K: one ascending sequence
            lo, hi = g.segs[k][0], g.segs[k][1]
            v = g.sig[lo:hi]
            srt = srt and bool((v[1:] >= v[:-1]).all().item())
            if k:
                srt = srt and bool((v[0] >= g.sig[g.segs[k - 1][1] - 1]).item())
        out["percentile50_bubble_passes"] = {"ms": round(best, 1), "K": K, "steps": steps, "cells_exchanged": moved,
                                             "front_sorted": srt, "timing": "wall clock around the host loop (one sync per step)"}
        print("percentile50_bubble_passes", out["percentile50_bubble_passes"], flush=True)
    if on("morph"):
        rec("open_1001", timed(lambda: g.open_(1001, 6.0), reset), 16)
        rec("close_1001", timed(lambda: g.close_(1001, 6.0), reset), 16)
        rec("dilate_1001", timed(lambda: g.dilate(1001, threshold=6.0), reset), 16)
    if on("clump"):
        rec("clump_L1000", timed(lambda: g.clump(6.5, 1000), reset), 16)
    if on("runs"):
        def bin_setup():
            reset(); g.binarize(9.0)
        # kernel only: results stay on the device (the CLI formats them there, gdsp_format_runs)
        capr = max(1024, N // 4)
        bufs = (torch.empty(capr, dtype=torch.int32, device=g.device), torch.empty(capr, dtype=torch.int32, device=g.device),
                torch.empty(capr, dtype=torch.float64, device=g.device))
        nr = [0]
        def do_runs():
            nr[0] = g.runs_device(bufs)[0]
        ms = timed(do_runs, bin_setup)
        rec("runs_binarized", ms, 8, 16 * nr[0])
        out["runs_binarized"]["runs"] = nr[0]
        def few_setup():
            reset(); g.binarize(14.0)
        ms = timed(do_runs, few_setup)
        rec("runs_sparse", ms, 8, 16 * nr[0])
        out["runs_sparse"]["runs"] = nr[0]
        del bufs
    # whole pipelines: wall time of the chain, with per-operator CUDA-event splits
    def chain(name, ops):
        evs = None
        best = None
        for r in range(args.reps + 1):
            torch.cuda.synchronize()
            marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
            marks[0].record()
            for k, (_, fn) in enumerate(ops):
                fn(); marks[k + 1].record()
            torch.cuda.synchronize()
            ms = marks[0].elapsed_time(marks[-1])
            if r > 0 and (best is None or ms < best):
                best = ms; evs = [marks[k].elapsed_time(marks[k + 1]) for k in range(len(ops))]
        out[name] = {"ms": round(best, 3), "gbp_s": round(N / (best / 1e3) / 1e9, 2),
                     "ops": {ops[k][0]: round(evs[k], 3) for k in range(len(ops))}}
        print(name, out[name], flush=True)

    if on("pipe5"):
        chain("pipe5_depth_smooth_localmax_percentile_binarize", [
            ("depth", lambda: g.accumulate(seg, st, en, host=False)),
            ("smooth101", lambda: g.smooth(101)),
            ("localmax11", lambda: g.localmax(11)),
            ("percentile99", lambda: g.percentile(99.0)),
            ("binarize", lambda: g.binarize(g.variables["percentile99"])),
            ("runs", lambda: g.runs(cap=max(1024, N // 8)))])
    if on("cfg3"):
        chain("cfg3_depth_sum100_percentile99_binarize", [
            ("depth", lambda: g.accumulate(seg, st, en, host=False)),
            ("sum100", lambda: g.sum(100, denom=100.0)),
            ("percentile99", lambda: g.percentile(99.0)),
            ("binarize", lambda: g.binarize(g.variables["percentile99"])),
            ("runs", lambda: g.runs(cap=max(1024, N // 8)))])
    if on("cfg4"):
        chain("cfg4_depth_binarize_open_close_clump", [
            ("depth", lambda: g.accumulate(seg, st, en, host=False)),
            ("binarize6", lambda: g.binarize(6.0)),
            ("open1001", lambda: g.open_(1001, 0.5)),
            ("close1001", lambda: g.close_(1001, 0.5)),
            ("clump", lambda: g.clump(0.5, 1000)),
            ("runs", lambda: g.runs(cap=max(1024, N // 8)))])
    print(json.dumps({"scale": args.scale, "bases": N, "hbm_peak_gbs": peak, "stages": out}))


if __name__ == "__main__":
    main()
