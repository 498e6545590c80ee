#!/usr/bin/env python
"""percentile 99 on the depth track and on the smoothed track (for an ncu launch list)"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=1)
a = ap.parse_args()
import torch
from genodsp_b200.genome import Genome
chroms = bench.scaled_genome(a.scale)
g = Genome(chroms)
order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
g.accumulate(seg, st, en, host=False)
g.percentile(99.0, destructive=False)
g.smooth(101)
torch.cuda.synchronize()
print("MARK smoothed percentile starts", flush=True)
g.percentile(99.0, destructive=False)
torch.cuda.synchronize()
print("done")
