"""CPU model of gdsp_merge.cu (the percentile bubble step as two merges): the split count and the
per-tile / per-thread merge-path corners are restated with the kernel's own index arithmetic and checked
against "sort the pair jointly" -- which is what combine_sorted_vectors (percentile.c:820-864) amounts
to -- on sorted inputs with ties, empty sides and lengths around the tile size."""
import numpy as np
import pytest


def split_count(C, D):                       # k_mx_split
    lo, hi = 0, min(len(C), len(D))
    while lo < hi:
        mid = (lo + hi) >> 1
        if D[mid] < C[len(C) - 1 - mid]:
            lo = mid + 1
        else:
            hi = mid
    return lo


def reference_count(v, w):                   # the scan of combine_sorted_vectors
    if len(v) == 0 or len(w) == 0 or v[-1] <= w[0]:
        return 0
    wi, vi = 0, len(v) - 1
    while wi + 1 < len(w) and vi > 0:
        wi += 1; vi -= 1
        if v[vi] > w[wi]:
            continue
        wi -= 1; vi += 1
        break
    return wi + 1


def corner(A, B, d):                         # merge_corner and the per-thread search
    lo, hi = max(0, d - len(B)), min(d, len(A))
    while lo < hi:
        mid = (lo + hi) >> 1
        assert 0 <= d - 1 - mid < len(B) and 0 <= mid < len(A)
        if not (B[d - 1 - mid] < A[mid]):
            lo = mid + 1
        else:
            hi = mid
    return lo


def merge_tiled(A, B, tile, per):            # k_merge, block by block and thread by thread
    n = len(A) + len(B)
    out = np.full(n, -1, np.int64)
    threads = tile // per
    for blk in range((n + tile - 1) // tile):
        d0 = blk * tile
        d1 = min(d0 + tile, n)
        a0, a1 = corner(A, B, d0), corner(A, B, d1)
        b0, b1 = d0 - a0, d1 - a1
        sA, sB = A[a0:a1], B[b0:b1]
        la, lb = len(sA), len(sB)
        total = la + lb
        assert total == d1 - d0 and b0 >= 0 and b1 >= b0 and a1 >= a0
        for t in range(threads):
            ld = min(t * per, total)
            ai = corner(sA, sB, ld)
            bi = ld - ai
            for k in range(per):
                o = ld + k
                if o >= total:
                    break
                take_a = bi >= lb or (ai < la and not (sB[bi] < sA[ai]))
                if take_a:
                    out[d0 + o] = sA[ai]; ai += 1
                else:
                    out[d0 + o] = sB[bi]; bi += 1
    return out


@pytest.mark.parametrize("seed", range(40))
def test_exchange_by_merges_equals_joint_sort(seed):
    rng = np.random.default_rng(seed)
    tile, per = 16, 4
    nc = int(rng.choice([0, 1, 2, 15, 16, 17, 31, 33, 100]))
    nd = int(rng.choice([0, 1, 3, 16, 32, 47, 64, 90]))
    hi = int(rng.choice([2, 5, 1000]))                       # few distinct values: many ties
    C = np.sort(rng.integers(0, hi, nc) + int(rng.integers(-3, 4)))
    D = np.sort(rng.integers(0, hi, nd))
    m = split_count(C, D) if nc and nd else 0
    assert m == reference_count(list(C), list(D))
    both = np.sort(np.concatenate([C, D]))
    c2 = merge_tiled(C[:nc - m], D[:m], tile, per)
    d2 = merge_tiled(C[nc - m:], D[m:], tile, per)
    assert np.array_equal(c2, both[:nc]) and np.array_equal(d2, both[nc:])
