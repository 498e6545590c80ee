# scratch: pipe5 up to localmax, then percentile(destructive) once -- for an ncu launch list of the post-state path
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
from genodsp_b200.genome import Genome
chroms = bench.scaled_genome(1)
g = Genome(chroms)
order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
g.accumulate(seg, st, en, host=False); g.smooth(101); g.localmax(11)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = g.percentile(99.0); b.record(); torch.cuda.synchronize()
print(out, a.elapsed_time(b), "ms", flush=True)
nz = int((g.sig != 0).sum().item()); print("nonzero cells", nz, "of", g.cells)
