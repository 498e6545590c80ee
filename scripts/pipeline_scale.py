#!/usr/bin/env python
"""BASELINE configs 3-5 on N GPUs (slab-sharded, one process per GPU over NCCL), hg38/--scale:
  cfg3  depth -> sum 100 /100 -> percentile 99 -> binarize(percentile99)
  cfg4  depth -> binarize 6 -> open 1001 -> close 1001 -> clump 0.5 L=1000 -> run-length detection
  cfg5  depth -> add B -> multiply B -> mask M -> and B -> binarize 0.5   (one fused pointwise launch)
Each pipeline is timed with CUDA events between barriers (max over ranks), best of --reps.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29580 scripts/pipeline_scale.py --scale 1
Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from genodsp_b200 import capi, slab
    from genodsp_b200.genome import Genome
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    chroms = bench.scaled_genome(args.scale)
    order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
    sorted_chroms = [chroms[i] for i in order]
    lengths = [l for _, l in sorted_chroms]
    HALO = 1024                                        # open/close 1001 reach 1003 cells
    segs_s, cells = slab.partition(lengths, world, rank, HALO)
    segs = [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s]
    g = Genome(chroms, device=local, segs=segs, buffer_cells=cells)
    plan = slab.halo_plan(lengths, world, rank, HALO)
    gather = slab.dist_gather(dist)
    cs, st, en = bench.synth_intervals(torch, device, sorted_chroms)
    ks, ss, es = [], [], []
    for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
        si = order.index(ci)
        msk = (cs == si) & (en.to(torch.int64) > pos0) & (st.to(torch.int64) < pos0 + (hi - lo))
        ks.append(torch.full((int(msk.sum()),), k, dtype=torch.int32, device=device)); ss.append(st[msk]); es.append(en[msk])
    seg_t, start_t, end_t = torch.cat(ks), torch.cat(ss), torch.cat(es)
    del cs, st, en

    # second track for cfg5: sorted disjoint intervals covering ~50 %, values k/1024 (per owned piece)
    rng = np.random.default_rng(99)
    bs, bstart, bend, bval = [], [], [], []
    for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
        n = hi - lo
        m = max(1, n // 2000)
        cuts = np.sort(rng.choice(np.arange(pos0, pos0 + n, dtype=np.int64), size=min(2 * m, n), replace=False))
        a, b = cuts[0::2], cuts[1::2]
        m2 = min(a.size, b.size)
        bs.append(np.full(m2, k, np.uint32)); bstart.append(a[:m2].astype(np.uint32)); bend.append(b[:m2].astype(np.uint32))
        bval.append(rng.integers(1, 2048, m2) / 1024.0)
    tableB = g.interval_table(np.concatenate(bs), np.concatenate(bstart), np.concatenate(bend), np.concatenate(bval))
    capr = max(1024, g.cells // 4)
    rbufs = (torch.empty(capr, dtype=torch.int32, device=device), torch.empty(capr, dtype=torch.int32, device=device),
             torch.empty(capr, dtype=torch.float64, device=device))
    factory = lambda name, clen, r: Genome([(name, clen)], device=local)

    def depth():
        g.accumulate(seg_t, start_t, end_t, host=False)

    def cfg3():
        depth(); g.sum(100, denom=100.0)
        slab.slab_percentile_then_binarize([g], gather, 99000)

    def cfg4():
        depth(); g.binarize(6.0)
        slab.exchange_halos(g.sig, plan, dist); g.open_(1001, 0.5)
        slab.exchange_halos(g.sig, plan, dist); g.close_(1001, 0.5)
        slab.slab_clump(slab.DistTransport(g, dist), gather, factory, average=0.5, length=1000)
        g.runs_device(rbufs)

    def cfg5():
        depth()
        g.pointwise([(capi.PW_IVL_ADD, 0.0, 0, 0, 0, tableB), (capi.PW_IVL_MUL, 0.0, 0, 0, 0, tableB),
                     (capi.PW_IVL_SET, 0.0, 0, 0, 0, tableB), (capi.PW_NONZERO_TO_ONE, 0.0),
                     (capi.PW_IVL_SET_OUTSIDE, 0.0, 0, 0, 0, tableB), type(g).op_binarize(0.5)])

    total_bases = sum(lengths)
    out = {}
    for name, fn in (("cfg3", cfg3), ("cfg4", cfg4), ("cfg5", cfg5)):
        best = None
        for r in range(args.reps + 1):
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize(); dist.barrier()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if r > 0:
                best = float(t.item()) if best is None else min(best, float(t.item()))
        out[name] = {"ms": round(best, 3), "gbp_s": round(total_bases / (best / 1e3) / 1e9, 2)}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "scale": args.scale, "bases": total_bases, "pipelines": out}), flush=True)
    tableB.close(); g.close()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
