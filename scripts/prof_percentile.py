#!/usr/bin/env python
"""depth -> smooth -> percentile 99 (select) -> sort_genome on hg38/--scale; used under ncu to list
the per-kernel times of the percentile path."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=16); ap.add_argument("--smooth", type=int, default=1)
a = ap.parse_args()
import torch
from genodsp_b200.genome import Genome
chroms = bench.scaled_genome(a.scale)
g = Genome(chroms)
order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
g.accumulate(seg, st, en, host=False)
if a.smooth:
    g.smooth(101)
torch.cuda.synchronize()
print(g.percentile(99.0, destructive=False))
g.sort_genome()
torch.cuda.synchronize()
print("done")
