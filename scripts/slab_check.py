#!/usr/bin/env python
"""N-GPU check of the slab-sharded pipeline with REAL ranks (NCCL):
  depth -> smooth 101 -> localmax 11 -> percentile 99 -> binarize -> dilate 150 -> close 200 -> clump -> run-length output
on hg38/--scale, one process per GPU, against the same pipeline on one whole-genome Genome on rank 0.
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29533 scripts/slab_check.py --scale 16
Prints one line per rank-0 check and `SLAB_CHECK PASS` / `FAIL`."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=16)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from genodsp_b200 import slab
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
    HALO = 256
    segs_s, cells = slab.partition(lengths, world, rank, HALO)
    segs = [(order[si], lo, hi, dlo, dhi, pos0) for si, lo, hi, dlo, dhi, pos0 in segs_s]
    g = Genome(chroms, device=local, segs=segs, buffer_cells=cells)
    plan = slab.halo_plan(lengths, world, rank, HALO)
    cs, st, en = bench.synth_intervals(torch, device, sorted_chroms)
    keep_seg, keep_s, keep_e = [], [], []
    for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
        si = order.index(ci)
        msk = (cs == si) & (en.to(torch.int64) > pos0) & (st.to(torch.int64) < pos0 + (hi - lo))
        keep_seg.append(torch.full((int(msk.sum()),), k, dtype=torch.int32, device=device))
        keep_s.append(st[msk]); keep_e.append(en[msk])
    gather = slab.dist_gather(dist)

    g.accumulate(torch.cat(keep_seg), torch.cat(keep_s), torch.cat(keep_e), host=False)
    slab.exchange_halos(g.sig, plan, dist)
    g.smooth(101)
    slab.exchange_halos(g.sig, plan, dist)
    g.localmax(11)
    (p99,), n = slab.slab_percentiles([g], gather, [99000])
    g.binarize(p99)
    slab.exchange_halos(g.sig, plan, dist)
    g.dilate(150, threshold=0.5)                   # reach 76 cells: inside the halo
    slab.exchange_halos(g.sig, plan, dist)
    g.close_(200, 0.5)                             # reach 202 cells
    # clump has no bounded reach: chromosomes cut by a slab boundary are reassembled on one rank
    slab.slab_clump(slab.DistTransport(g, dist), gather, lambda name, clen, r: Genome([(name, clen)], device=local),
                    average=0.97, length=300)
    runs = slab.slab_runs([g], gather)
    total_cum = None
    # cumulative sum of the binary track: the last cell of every chromosome = number of ones
    slab.slab_cumulativesum([g], gather)
    ends = {}
    for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
        if pos0 + (hi - lo) == chroms[ci][1]:
            ends[chroms[ci][0]] = float(g.sig[hi - 1].item())
    all_ends = {}
    for d in gather([ends]):
        all_ends.update(d)

    ok = True
    if rank == 0:
        w = Genome(chroms, device=local)
        w.accumulate(cs, st, en, host=False)
        w.smooth(101); w.localmax(11)
        want = w.percentile(99.0, destructive=False)["percentile99"]
        w.binarize(want)
        w.dilate(150, threshold=0.5); w.close_(200, 0.5)
        w.clump(0.97, 300)
        wr = w.runs()
        print("percentile99 slabs=%r whole=%r samples=%d" % (p99, want, n), flush=True)
        ok = ok and (p99 == want) and n == sum(lengths)
        nruns = 0
        for name, _ in chroms:
            a, b = runs.get(name), wr.get(name)
            if b is None or b[0].size == 0:
                ok = ok and (a is None or a[0].size == 0)
                continue
            same = all(np.array_equal(np.asarray(x, np.float64), np.asarray(y, np.float64)) for x, y in zip(a, b))
            ok = ok and same
            nruns += b[0].size
            ones = float(np.sum((b[1].astype(np.int64) - b[0].astype(np.int64)) * b[2]))
            ok = ok and (all_ends.get(name) == ones)
        print("runs compared: %d over %d chromosomes; cumulative-sum chromosome totals checked" % (nruns, len(chroms)), flush=True)
        print("SLAB_CHECK %s (world %d, scale %d)" % ("PASS" if ok else "FAIL", world, args.scale), flush=True)
        w.close()
    g.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
