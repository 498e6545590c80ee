"""The accumulator schedule of the shared-product smooth kernel (k_smooth_sym), restated on the CPU in
scripts/smooth_sym_model.py, against the reference's ascending-tap fold (sum.c:651-664) and the oracle:
same bits for every output, including strips that start or end inside a chromosome."""
import os
import sys

import numpy as np
import pytest

from checkers import Oracle
from genodsp_b200.genome import hann_taps

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import smooth_sym_model as M  # noqa: E402


@pytest.mark.parametrize("W,T,K", [(9, 1, 4), (31, 1, 15), (101, 1, 50), (101, 2, 25), (21, 3, 4)])
def test_schedule_matches_direct_fold(W, T, K):
    M.check(W, T, K, n=260, seed=W)


def test_direct_fold_is_the_oracle():
    orc = Oracle()
    rng = np.random.default_rng(3)
    v = rng.normal(size=400) * 10.0 ** rng.integers(-3, 4, 400)
    for W in (9, 31, 101):
        want = orc.smooth(v.copy(), W)
        got = M.direct(v, hann_taps(W))             # the window exactly as sum.c:634-650 builds it
        assert want.tobytes() == got.tobytes(), W
