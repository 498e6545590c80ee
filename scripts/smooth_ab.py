"""A/B timing of gdsp_smooth's two kernels on the hg38 layout (CUDA events, best of 3 after warm-up), with a bit
comparison of their results.  usage: smooth_ab.py <scale> <W> ...   (GDSP_SYM_STRIP=<cells> overrides the strip length)"""
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
g.accumulate(seg, st, en, host=False)
depth = g.sig.clone()
def timed(fn, reps=3):
    best = None
    for r in range(reps + 1):
        g.sig.copy_(depth); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if r: best = ms if best is None else min(best, ms)
    return best
last = None
for spec in (sys.argv[2:] or ["101", "31", "11", "1001", "55"]):
    W = int(spec)
    if last != W:
        g.sig.copy_(depth); g.smooth(W, direct=True); ref = g.sig.clone()
        td = timed(lambda: g.smooth(W, direct=True)); last = W
    g.sig.copy_(depth); g.smooth(W); same = bool(torch.equal(ref.view(torch.int64), g.sig.view(torch.int64)))
    ts = timed(lambda: g.smooth(W))
    print("%-12s direct %.3f ms  shared-product %.3f ms  ratio %.3f  bit_equal %s" % (spec, td, ts, td / ts, same), flush=True)
