#!/usr/bin/env python
"""one launch of each stage kernel on hg38/--scale (for ncu)"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=16)
a = ap.parse_args()
import torch
from genodsp_b200.genome import Genome
chroms = bench.scaled_genome(a.scale)
g = Genome(chroms)
order = sorted(range(len(chroms)), key=lambda i: -chroms[i][1])
seg, st, en = bench.synth_intervals(torch, g.device, [chroms[i] for i in order])
g.accumulate(seg, st, en, host=False)
depth = g.sig.clone()
tableB = bench.second_track(g, __import__("numpy"))
from genodsp_b200 import capi as c


def chain():
    g.pointwise([(c.PW_IVL_ADD, 0.0, 0, 0, 0, tableB), (c.PW_IVL_MUL, 0.0, 0, 0, 0, tableB), (c.PW_IVL_SET, 0.0, 0, 0, 0, tableB),
                 (c.PW_NONZERO_TO_ONE, 0.0), (c.PW_IVL_SET_OUTSIDE, 0.0, 0, 0, 0, tableB), type(g).op_binarize(0.5)])


def smoothed_percentile():
    g.smooth(101)
    g.percentile(99.0, destructive=False)


for fn in (lambda: g.localmax(11), lambda: g.bestmax(101), lambda: g.slidingsum(101), lambda: g.sum(100),
           lambda: g.cumulativesum(), lambda: g.open_(1001, 6.0), lambda: g.clump(6.5, 1000),
           lambda: g.binarize(9.0), lambda: g.runs(), lambda: g.percentile(99.0, destructive=False), chain,
           smoothed_percentile):
    g.sig.copy_(depth)
    fn()
torch.cuda.synchronize()
print("done")
