"""Depth accumulation of the bench's reads on the hg38 layout: as generated (random order inside a chromosome) and
sorted by position (what a coordinate-sorted alignment file gives).  CUDA events, best of 3 after warm-up."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch
from genodsp_b200.genome import Genome
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 1
chroms = bench.scaled_genome(scale)
g = Genome(chroms)
order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
def timed(fn, reps=3):
    best = None
    for r in range(reps + 1):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r: best = ms if best is None else min(best, ms)
    return best
t_rand = timed(lambda: g.accumulate(seg, st, en, host=False)); ref = g.sig.clone()
key = seg.to(torch.int64) * (1 << 32) + st.to(torch.int64)
idx = torch.argsort(key)
seg2, st2, en2 = seg[idx].contiguous(), st[idx].contiguous(), en[idx].contiguous()
t_sort = timed(lambda: g.accumulate(seg2, st2, en2, host=False))
same = bool(torch.equal(ref.view(torch.int64), g.sig.view(torch.int64)))
print("accumulate %d reads: random order %.3f ms, position-sorted %.3f ms, same depth %s" % (int(seg.shape[0]), t_rand, t_sort, same))
