import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from genodsp_b200.genome import Genome
g = Genome([("a", 30000), ("b", 100)])
rng = np.random.default_rng(0)
v = rng.normal(0,3,30000); g.set_chrom("a", v); g.set_chrom("b", rng.normal(0,3,100))
for W in (6145, 6147, 8001):
    g.set_chrom("a", v)
    g.bestmax(W); g.sync()
    l=(W-1)//2; r=W-1-l
    print("bestmax", W, g.get_chrom("a")[:3], "expected", v[:r+1].max(), flush=True)
