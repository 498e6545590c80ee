"""numpy stand-in for genodsp_b200.genome.Genome -- TEST INFRASTRUCTURE ONLY.

Implements just the per-rank calls the slab-level operators of genodsp_b200/slab.py make
(pct_sample, pct_count, minmax, cumulativesum, piece_last_values, piece_add_constant, runs,
pointwise/op_invert) so that their cross-rank logic (what is gathered, how brackets narrow, how
carries and cut runs are merged) can be exercised on CPU with world_size-2 gloo.  The product path
never imports this."""
import numpy as np

from genodsp_b200 import slab


class HostGenome:
    def __init__(self, chroms, segs, signal):
        """segs: [(chrom_index, lo, hi, dlo, dhi, pos0)]; signal: {name: whole-chromosome array}"""
        self.chroms = chroms
        self.seg_chrom = [s[0] for s in segs]
        self.segs = [(s[1], s[2], s[3], s[4], s[5], chroms[s[0]][1]) for s in segs]
        self.nseg = len(segs)
        self.piece = [np.array(signal[chroms[s[0]][0]][s[5]:s[5] + s[2] - s[1]], dtype=np.float64) for s in segs]
        import torch
        self.torch, self.device = torch, torch.device("cpu")

    def seg_index(self, name):
        ci = [i for i, (n, _) in enumerate(self.chroms) if n == name][0]
        return [k for k, c in enumerate(self.seg_chrom) if c == ci]

    def _samples(self, stride, mn, mx):
        out = []
        for k, v in enumerate(self.piece):
            pos0 = self.segs[k][4]
            first = (-pos0) % stride
            s = v[first::stride]
            out.append(s[~(s < mn) & ~(s > mx)])
        return np.concatenate(out) if out else np.zeros(0)

    def minmax(self, stride=1, mn=-np.inf, mx=np.inf):
        s = self._samples(stride, mn, mx)
        return (float(s.min()), float(s.max()), int(s.size)) if s.size else (0.0, 0.0, 0)

    def pct_sample(self, m, stride=1, mn=-np.inf, mx=np.inf, key_lo=0, key_hi=2 ** 64 - 1, seed=1):
        s = self._samples(stride, mn, mx)
        rng = np.random.default_rng(seed)
        pick = s[rng.integers(0, s.size, min(m, 4096))] if s.size else s
        k = slab.f64_keys(pick)
        return pick[(k >= np.uint64(key_lo)) & (k <= np.uint64(key_hi))], int(s.size)

    def pct_count(self, bound_keys, compact, stride=1, mn=-np.inf, mx=np.inf, cap=None):
        s = self._samples(stride, mn, mx)
        k = slab.f64_keys(s)
        b = np.array(bound_keys, dtype=np.uint64)
        lo = np.searchsorted(b, k, "left")
        is_b = (lo < b.size) & (b[np.minimum(lo, max(b.size - 1, 0))] == k) if b.size else np.zeros(k.size, bool)
        reg = 2 * lo + is_b
        counts = np.bincount(reg, minlength=2 * b.size + 1).astype(np.uint64)
        sel = ~is_b & np.array(compact, bool)[lo]
        cand = s[sel]
        if cap is not None and cand.size > cap:
            cand = None
        return counts, cand

    def pct_sample_dev(self, m, stride=1, mn=-np.inf, mx=np.inf, key_lo=0, key_hi=2 ** 64 - 1, seed=1):
        pick, slots = self.pct_sample(m, stride, mn, mx, key_lo, key_hi, seed)
        return self.torch.from_numpy(np.ascontiguousarray(pick)), slots

    def pct_count_dev(self, bound_keys, compact, stride=1, mn=-np.inf, mx=np.inf, cap=None):
        s = self._samples(stride, mn, mx)
        counts, cand = self.pct_count(bound_keys, compact, stride, mn, mx, None)
        n = int(cand.size)
        fits = cap is None or n <= cap
        self.last_nan_count = int(np.isnan(s).sum())
        return counts, n, (self.torch.from_numpy(np.ascontiguousarray(cand)) if fits else None)

    def equal_range(self, sorted_t, value):
        k = slab.f64_keys(sorted_t.numpy())
        kv = slab.f64_keys(np.array([value]))[0]
        return int(np.searchsorted(k, kv, "left")), int(np.searchsorted(k, kv, "right"))

    def sort_array(self, a):
        v = a.numpy()
        return self.torch.from_numpy(np.ascontiguousarray(v[np.argsort(slab.f64_keys(v), kind="stable")]))

    def cumulativesum(self):
        self.piece = [np.cumsum(v) for v in self.piece]

    def piece_last_values(self):
        return np.array([v[-1] for v in self.piece])

    def piece_add_constant(self, k, value):
        self.piece[k] = self.piece[k] + value

    @staticmethod
    def op_invert(mid):
        return ("invert", mid)

    def pointwise(self, ops):
        for code, a in ops:
            assert code == "invert"
            self.piece = [2 * a - v for v in self.piece]

    def runs(self, collapse=True, show_uncovered=0):
        out = {}
        for k, v in enumerate(self.piece):
            pos0 = self.segs[k][4]
            if collapse:
                head = np.concatenate([[True], v[1:] != v[:-1]])
            else:
                head = np.ones(v.size, bool)
            st = np.nonzero(head)[0]
            en = np.concatenate([st[1:], [v.size]])
            val = v[st]
            keep = (val != 0) | bool(show_uncovered)
            out[self.chroms[self.seg_chrom[k]][0]] = ((st[keep] + pos0).astype(np.uint32), (en[keep] + pos0).astype(np.uint32), val[keep])
        return out
