"""Host-side mirror of the reference's per-base operator set on top of the C-ABI.

`Genome` plays the role of the reference's chromosome table + value vectors
(`spec`, `chromsOfInterest`, `chromsSorted`; genodsp_interface.h:37-57): one
device buffer holds every chromosome back to back in `chromsSorted` order
(descending length, genodsp.c:1104-1143), and each method is one reference
operator (same names, same argument meaning) executed by the sm_100a kernels.
torch is used only for device memory and streams.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import check

DBL_MAX = float(np.finfo(np.float64).max)


def hann_taps(window):
    """Hann window exactly as sum.c:634-645 builds it (host libm cos, symmetric
    fill from both ends, normalised by the left-to-right total)."""
    import math
    W = int(window)
    h = (W - 1) // 2
    w = np.zeros(W, np.float64)
    for k in range(h + 1):
        x = (k + 1) / float(W + 1)
        w[k] = w[W - 1 - k] = (1 - math.cos(2 * math.pi * x)) / 2
    tot = 0.0
    for k in range(W):
        tot += float(w[k])
    for k in range(W):
        w[k] = float(w[k]) / tot
    return w


def percentile_rank(num_values, p_milli):
    """(u32) ((u64) numValues * p / 100000.0), percentile.c:588/:686."""
    return int(np.uint32(np.float64(np.uint64(np.uint32(num_values)) * np.uint64(p_milli)) / np.float64(100000.0)))


def percentile_name(p_milli):
    """set_percentile_name, percentile.c:756-780."""
    if p_milli % 1000 == 0:
        return "percentile%d" % (p_milli // 1000)
    pct = np.float32(p_milli) / np.float32(1000)
    precision, denom = 1, 100
    while denom >= 1:
        if p_milli % denom == 0:
            return "percentile%.*f" % (precision, float(pct))
        precision += 1
        denom //= 10
    return "percentile%f" % float(pct)


class Genome:
    def __init__(self, chroms, device=0, segs=None, buffer_cells=None):
        """chroms: [(name, length), ...] in input ("file") order.

        segs/buffer_cells: explicit segment table for slab-sharded runs (one entry
        per chromosome piece owned by this GPU, in sorted order, with the piece's
        chromosome index); when omitted every chromosome is whole on this GPU."""
        import torch
        self.torch = torch
        if not torch.cuda.is_available():
            raise capi.GdspError("genodsp_b200 needs a CUDA device (no CPU fallback)")
        self.lib = capi.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.chroms = [(str(n), int(l)) for n, l in chroms]
        # chromsSorted: descending length; glibc qsort is a stable merge sort for small arrays
        self.order = sorted(range(len(self.chroms)), key=lambda i: -self.chroms[i][1])
        self.variables = {}
        ctx = C.c_void_p()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.gdsp_ctx_create(device, C.c_void_p(stream), C.byref(ctx)))
        self.ctx = ctx
        if segs is None:
            lens = (C.c_uint32 * len(self.order))(*[self.chroms[i][1] for i in self.order])
            seg_arr = (capi.Seg * len(self.order))()
            total = C.c_uint64()
            check(self.lib.gdsp_layout_pack(lens, len(self.order), seg_arr, C.byref(total)))
            self.seg_chrom = list(self.order)          # segment -> chromosome index (file order)
            self.buffer_cells = int(total.value)
        else:
            seg_arr = (capi.Seg * len(segs))()
            self.seg_chrom = []
            for k, (ci, lo, hi, dlo, dhi, pos0) in enumerate(segs):
                seg_arr[k].lo, seg_arr[k].hi, seg_arr[k].dlo, seg_arr[k].dhi = lo, hi, dlo, dhi
                seg_arr[k].pos0, seg_arr[k].chrom_len = pos0, self.chroms[ci][1]
                self.seg_chrom.append(ci)
            self.buffer_cells = int(buffer_cells)
        self.nseg = len(seg_arr)
        self.segs = [(s.lo, s.hi, s.dlo, s.dhi, s.pos0, s.chrom_len) for s in seg_arr]
        lay = C.c_void_p()
        check(self.lib.gdsp_layout_create(self.ctx, seg_arr, self.nseg, C.byref(lay)))
        self.layout = lay
        self._pending_sort = False
        self._sig = torch.zeros(self.buffer_cells, dtype=torch.float64, device=self.device)
        self.tmp = torch.zeros(self.buffer_cells, dtype=torch.float64, device=self.device)
        self._work = None
        self.cells = int(self.lib.gdsp_layout_cells(self.layout))
        self._launch0 = self.lib.gdsp_launch_count()

    # ------------------------------------------------------------------ plumbing
    @property
    def sig(self):
        """the current signal.  A destructive percentile leaves the genome notionally sorted
        (percentile.c:611-651); the sort only runs when somebody looks -- here -- unless the next
        operator is `binarize`, which never needs it (gdsp_sorted_binarize)."""
        if self._pending_sort:
            self._pending_sort = False
            self.sort_genome()
        return self._sig

    @sig.setter
    def sig(self, t):
        self._sig = t

    @property
    def launches(self):
        """kernels launched by the library since this genome was created (gdsp_launch_count)"""
        return int(self.lib.gdsp_launch_count() - self._launch0)

    def close(self):
        for lay in getattr(self, "_piece_layouts", None) or []:
            self.lib.gdsp_layout_destroy(lay)
        self._piece_layouts = None
        if getattr(self, "layout", None):
            self.lib.gdsp_layout_destroy(self.layout); self.layout = None
        if getattr(self, "ctx", None):
            self.lib.gdsp_ctx_destroy(self.ctx); self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _p(self, t):
        return C.c_void_p(t.data_ptr())

    def _swap(self):
        self.sig, self.tmp = self.tmp, self.sig

    def work(self, nbytes):
        if self._work is None or self._work.numel() < nbytes:
            self._work = None
            self._work = self.torch.empty(int(nbytes), dtype=self.torch.uint8, device=self.device)
        return self._work

    def exact_order(self, on=True):
        """context manager: slidingsum / cumulativesum / clump in the reference's sequential summation order
        (gdsp_ctx_set_exact_order) -- bit-identical on general reals, slow"""
        g = self

        class _Mode:
            def __enter__(self_inner):
                check(g.lib.gdsp_ctx_set_exact_order(g.ctx, 1 if on else 0))
                return g

            def __exit__(self_inner, *exc):
                check(g.lib.gdsp_ctx_set_exact_order(g.ctx, 0))
                return False
        return _Mode()

    def sync(self):
        check(self.lib.gdsp_sync(self.ctx))

    def seg_index(self, name):
        """layout (sorted-order) segment indices of chromosome `name`"""
        ci = [i for i, (n, _) in enumerate(self.chroms) if n == name][0]
        return [k for k, c in enumerate(self.seg_chrom) if c == ci]

    def set_chrom(self, name, values):
        """upload a whole chromosome vector (single-GPU layouts)"""
        k = self.seg_index(name)[0]
        lo, hi = self.segs[k][0], self.segs[k][1]
        v = np.ascontiguousarray(values, np.float64)
        assert v.size == hi - lo
        self.sig[lo:hi].copy_(self.torch.from_numpy(v))

    def get_chrom(self, name):
        k = self.seg_index(name)[0]
        lo, hi = self.segs[k][0], self.segs[k][1]
        return self.sig[lo:hi].cpu().numpy()

    def fill(self, value=0.0):
        check(self.lib.gdsp_fill(self.ctx, self.layout, self._p(self.sig), float(value)))

    # ------------------------------------------------------------------ input
    def accumulate(self, seg, start, end, val=None, mode=None, add=False, host=True):
        """read_intervals accumulate loops (genodsp.c:1307-1330), sum overlap.
        seg/start/end[/val]: numpy arrays (host=True) or torch cuda tensors (host=False);
        seg holds layout segment indices, start/end chromosome coordinates."""
        if mode is None:
            mode = capi.ACC_I32 if val is None else capi.ACC_F64
        wb = self.lib.gdsp_accumulate_work_bytes(self.layout, self.buffer_cells, mode)
        work = self.work(wb)
        n = int(seg.shape[0])
        if host:
            seg = np.ascontiguousarray(seg, np.uint32); start = np.ascontiguousarray(start, np.uint32)
            end = np.ascontiguousarray(end, np.uint32)
            vp = None
            if val is not None:
                val = np.ascontiguousarray(val, np.float64); vp = C.c_void_p(val.ctypes.data)
            check(self.lib.gdsp_accumulate_host(self.ctx, self.layout, self._p(self.sig), self.buffer_cells,
                                                self._p(work), C.c_void_p(seg.ctypes.data),
                                                C.c_void_p(start.ctypes.data), C.c_void_p(end.ctypes.data),
                                                vp, n, mode, int(add)))
        else:
            vp = None if val is None else self._p(val)
            check(self.lib.gdsp_accumulate_dev(self.ctx, self.layout, self._p(self.sig), self.buffer_cells,
                                               self._p(work), self._p(seg), self._p(start), self._p(end),
                                               vp, n, mode, int(add)))

    def accumulate_pinned(self, seg, start, end, val=None, mode=None, add=False):
        """like accumulate(host=True) for torch CPU tensors in pinned memory"""
        if mode is None:
            mode = capi.ACC_I32 if val is None else capi.ACC_F64
        wb = self.lib.gdsp_accumulate_work_bytes(self.layout, self.buffer_cells, mode)
        work = self.work(wb)
        vp = None if val is None else C.c_void_p(val.data_ptr())
        check(self.lib.gdsp_accumulate_host(self.ctx, self.layout, self._p(self.sig), self.buffer_cells,
                                            self._p(work), C.c_void_p(seg.data_ptr()), C.c_void_p(start.data_ptr()),
                                            C.c_void_p(end.data_ptr()), vp, int(seg.shape[0]), mode, int(add)))

    # ------------------------------------------------------------------ sum.c
    def sum(self, window=100, denom=1.0, denom_actual=False, zero=0.0, window_is_chromosome=False):
        check(self.lib.gdsp_block_sum(self.ctx, self.layout, self._p(self.sig), int(window),
                                      int(window_is_chromosome), float(denom), int(denom_actual), float(zero)))

    def slidingsum(self, window=100, denom=1.0):
        check(self.lib.gdsp_sliding_sum(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp),
                                        int(window), float(denom)))
        self._swap()

    def smooth(self, window=101, direct=False):
        """direct=True forces the direct FIR kernel (gdsp_ctx_set_smooth_direct); the default lets gdsp_smooth
        use the shared-product kernel for symmetric windows -- both give the same bits"""
        W = int(window)
        if W % 2 == 0:
            W += 1                                  # sum.c:565-570
        taps = hann_taps(W)
        if direct:
            check(self.lib.gdsp_ctx_set_smooth_direct(self.ctx, 1))
        try:
            check(self.lib.gdsp_smooth(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), W,
                                       taps.ctypes.data_as(C.POINTER(C.c_double))))
        finally:
            if direct:
                check(self.lib.gdsp_ctx_set_smooth_direct(self.ctx, 0))
        self._swap()

    def _ensure_piece_layouts(self):
        """one single-segment layout per owned chromosome piece (for per-piece calls)"""
        if getattr(self, "_piece_layouts", None) is None:
            self._piece_layouts = []
            for k in range(self.nseg):
                one = (capi.Seg * 1)()
                lo, hi, dlo, dhi, pos0, clen = self.segs[k]
                one[0].lo, one[0].hi, one[0].dlo, one[0].dhi, one[0].pos0, one[0].chrom_len = lo, hi, dlo, dhi, pos0, clen
                lay = C.c_void_p()
                check(self.lib.gdsp_layout_create(self.ctx, one, 1, C.byref(lay)))
                self._piece_layouts.append(lay)
        return self._piece_layouts

    # ------------------------------------------------------------------ per-rank pieces of the slab-sharded operators
    def piece_add_constant(self, k, value):
        """v += value on owned piece k only (carry fix-up of slab-sharded scans)"""
        lay = self._ensure_piece_layouts()[k]
        arr = (capi.PwOp * 1)()
        arr[0].code, arr[0].a, arr[0].b, arr[0].c, arr[0].flags, arr[0].table = capi.PW_ADDCONST, float(value), 0.0, 0.0, 0, None
        check(self.lib.gdsp_pointwise(self.ctx, lay, self._p(self.sig), self._p(self.sig), arr, 1))

    def piece_last_values(self):
        """value of the last owned cell of every piece (host floats)"""
        idx = self.torch.tensor([hi - 1 for (lo, hi, *_r) in self.segs], dtype=self.torch.int64, device=self.device)
        return self.sig[idx].cpu().numpy()

    def pct_sample(self, m, stride=1, mn=-DBL_MAX, mx=DBL_MAX, key_lo=0, key_hi=2 ** 64 - 1, seed=1):
        """gdsp_pct_sample -> (numpy samples, number of sample slots this rank owns)"""
        m = min(int(m), int(self.tmp.numel()))          # the scratch buffer bounds the sample
        cnt, slots = C.c_uint32(), C.c_uint64()
        check(self.lib.gdsp_pct_sample(self.ctx, self.layout, self._p(self.sig), int(stride), float(mn), float(mx),
                                       int(key_lo), int(key_hi), m, int(seed), self._p(self.tmp), C.byref(cnt), C.byref(slots)))
        return self.tmp[:int(cnt.value)].cpu().numpy(), int(slots.value)

    def pct_count(self, bound_keys, compact, stride=1, mn=-DBL_MAX, mx=DBL_MAX, cap=None):
        """gdsp_pct_count -> (region counts numpy u64[2*nb+1], candidates numpy or None if over cap)"""
        nb = len(bound_keys)
        cap = int(cap) if cap else int(self.tmp.numel())
        cap = min(cap, int(self.tmp.numel()))
        bk = (C.c_uint64 * max(nb, 1))(*[int(k) for k in bound_keys])
        cp = (C.c_uint8 * (nb + 1))(*[int(bool(x)) for x in compact])
        counts = (C.c_uint64 * (2 * nb + 1))()
        ncand = C.c_uint64()
        check(self.lib.gdsp_pct_count(self.ctx, self.layout, self._p(self.sig), int(stride), float(mn), float(mx),
                                      bk, nb, cp, counts, self._p(self.tmp), cap, C.byref(ncand)))
        n = int(ncand.value)
        cand = self.tmp[:n].cpu().numpy() if n <= cap else None
        return np.array(list(counts), dtype=np.uint64), cand

    def pct_sample_dev(self, m, stride=1, mn=-DBL_MAX, mx=DBL_MAX, key_lo=0, key_hi=2 ** 64 - 1, seed=1):
        """gdsp_pct_sample -> (device view of the samples inside self.tmp, sample slots this rank owns)"""
        m = min(int(m), int(self.tmp.numel()))
        cnt, slots = C.c_uint32(), C.c_uint64()
        check(self.lib.gdsp_pct_sample(self.ctx, self.layout, self._p(self.sig), int(stride), float(mn), float(mx),
                                       int(key_lo), int(key_hi), m, int(seed), self._p(self.tmp), C.byref(cnt), C.byref(slots)))
        return self.tmp[:int(cnt.value)], int(slots.value)

    def pct_count_dev(self, bound_keys, compact, stride=1, mn=-DBL_MAX, mx=DBL_MAX, cap=None):
        """gdsp_pct_count -> (region counts numpy u64[2*nb+1], number of candidates, device view of them inside
        self.tmp -- None when they did not fit `cap`)"""
        nb = len(bound_keys)
        cap = int(cap) if cap else int(self.tmp.numel())
        cap = min(cap, int(self.tmp.numel()))
        bk = (C.c_uint64 * max(nb, 1))(*[int(k) for k in bound_keys])
        cp = (C.c_uint8 * (nb + 1))(*[int(bool(x)) for x in compact])
        counts = (C.c_uint64 * (2 * nb + 1))()
        ncand, nnan = C.c_uint64(), C.c_uint64()
        check(self.lib.gdsp_pct_count_nan(self.ctx, self.layout, self._p(self.sig), int(stride), float(mn), float(mx),
                                          bk, nb, cp, counts, self._p(self.tmp), cap, C.byref(ncand), C.byref(nnan)))
        n = int(ncand.value)
        self.last_nan_count = int(nnan.value)
        return np.array(list(counts), dtype=np.uint64), n, (self.tmp[:n] if n <= cap else None)

    def sort_array(self, a):
        """gdsp_sort_array on a 1-D float64 device tensor -> sorted tensor (key order: -0.0 before +0.0)"""
        if a.numel() < 2:
            return a
        a = a.contiguous()
        b = self.torch.empty_like(a)
        flag = C.c_int()
        check(self.lib.gdsp_sort_array(self.ctx, self._p(a), self._p(b), int(a.numel()), C.byref(flag)))
        return b if flag.value else a

    def equal_range(self, sorted_t, value):
        """gdsp_equal_range: [lo, hi) of the cells with value's key in a tensor sorted by sort_array"""
        lo, hi = C.c_uint64(), C.c_uint64()
        check(self.lib.gdsp_equal_range(self.ctx, self._p(sorted_t), int(sorted_t.numel()), float(value), C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def smooth_to_host(self, window, out_host):
        """smooth + device->host delivery of the result, pipelined per chromosome piece: while piece
        k is on its way to `out_host` (a pinned float64 tensor of buffer_cells) on a copy stream, the
        FIR of piece k+1 runs.  Same arithmetic as smooth(); returns the bytes copied."""
        t = self.torch
        W = int(window)
        if W % 2 == 0:
            W += 1
        taps = hann_taps(W)
        self._ensure_piece_layouts()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = t.cuda.Stream(device=self.device)
        main = t.cuda.current_stream(self.device)
        copied = 0
        for k in range(self.nseg):
            lo, hi = self.segs[k][0], self.segs[k][1]
            check(self.lib.gdsp_smooth(self.ctx, self._piece_layouts[k], self._p(self.sig), self._p(self.tmp), W,
                                       taps.ctypes.data_as(C.POINTER(C.c_double))))
            done = t.cuda.Event(); done.record(main)
            self._copy_stream.wait_event(done)
            with t.cuda.stream(self._copy_stream):
                out_host[lo:hi].copy_(self.tmp[lo:hi], non_blocking=True)
            copied += 8 * (hi - lo)
        fin = t.cuda.Event(); fin.record(self._copy_stream)
        main.wait_event(fin)
        self._swap()
        return copied

    def cumulativesum(self):
        check(self.lib.gdsp_cumulative_sum(self.ctx, self.layout, self._p(self.sig), self._p(self.sig)))

    # ------------------------------------------------------------------ minmax.c
    @staticmethod
    def _odd3(n):
        n = max(int(n), 3)
        return n + 1 if n % 2 == 0 else n            # minmax.c:921-934

    def localmax(self, neighborhood=3, zero=0.0):
        check(self.lib.gdsp_local_extrema(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp),
                                          self._odd3(neighborhood), 1, float(zero)))
        self._swap()

    def localmin(self, neighborhood=3, infinity=DBL_MAX):
        check(self.lib.gdsp_local_extrema(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp),
                                          self._odd3(neighborhood), 0, float(infinity)))
        self._swap()

    def bestmax(self, window=100):
        check(self.lib.gdsp_best_extrema(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), int(window), 1))
        self._swap()

    def bestmin(self, window=100):
        check(self.lib.gdsp_best_extrema(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), int(window), 0))
        self._swap()

    # ------------------------------------------------------------------ morphology.c
    def _morph(self, kind, length, left, right, threshold, one, zero):
        wb = self.lib.gdsp_morph_work_bytes(self.buffer_cells)
        work = self.work(wb)
        check(self.lib.gdsp_morphology(self.ctx, self.layout, self._p(self.sig), self.buffer_cells, self._p(work),
                                       kind, float(length), int(left), int(right), float(threshold),
                                       float(one), float(zero)))

    def close_(self, length, threshold=0.0, one=1.0, zero=0.0):
        self._morph(capi.MORPH_CLOSE, length, 0, 0, threshold, one, zero)

    def open_(self, length, threshold=0.0, one=1.0, zero=0.0):
        self._morph(capi.MORPH_OPEN, length, 0, 0, threshold, one, zero)

    def dilate(self, length=None, left=0, right=0, threshold=0.0, one=1.0, zero=0.0):
        if left == 0 and right == 0:                 # morphology.c:924-928
            left = int(float(length) / 2); right = int(float(length) - left)
        self._morph(capi.MORPH_DILATE, length or 0, left, right, threshold, one, zero)

    def erode(self, length=None, left=0, right=0, threshold=0.0, one=1.0, zero=0.0):
        if left == 0 and right == 0:                 # morphology.c:1374-1378
            left = int(float(length) / 2); right = int(float(length) - left)
        self._morph(capi.MORPH_ERODE, length or 0, left, right, threshold, one, zero)

    # ------------------------------------------------------------------ pointwise chains
    def pointwise(self, ops):
        """ops: list of (code, a, b, c, flags, table) tuples (missing trailing fields = 0/None);
        the whole chain runs as ONE kernel."""
        arr = (capi.PwOp * len(ops))()
        for i, op in enumerate(ops):
            op = tuple(op) + (0.0, 0.0, 0.0, 0, None)[len(op) - 1:]
            arr[i].code, arr[i].a, arr[i].b, arr[i].c, arr[i].flags = int(op[0]), float(op[1]), float(op[2]), float(op[3]), int(op[4])
            arr[i].table = op[5].handle if op[5] is not None else None
        check(self.lib.gdsp_pointwise(self.ctx, self.layout, self._p(self.sig), self._p(self.sig), arr, len(ops)))

    @staticmethod
    def op_binarize(threshold=0.0, ties_above=False, one=1.0, zero=0.0):
        return (capi.PW_BINARIZE_GE if ties_above else capi.PW_BINARIZE_GT, threshold, one, zero)

    @staticmethod
    def op_addconst(value):
        return (capi.PW_ADDCONST, value)

    @staticmethod
    def op_abs():
        return (capi.PW_ABS, 0.0)

    @staticmethod
    def op_clip(mn=None, mx=None):
        if mn is not None and mx is not None:
            return (capi.PW_CLIP_BOTH, mn, mx)
        return (capi.PW_CLIP_MIN, mn) if mx is None else (capi.PW_CLIP_MAX, mx)

    @staticmethod
    def op_erase(mn=None, mx=None, keep_inside=False, zero=0.0):
        fl = (capi.PW_ERASE_HAVE_MIN if mn is not None else 0) | (capi.PW_ERASE_HAVE_MAX if mx is not None else 0) \
            | (capi.PW_ERASE_KEEP_INSIDE if keep_inside else 0)
        return (capi.PW_ERASE, mn or 0.0, mx or 0.0, zero, fl)

    @staticmethod
    def op_invert(mid):
        return (capi.PW_INVERT, 2 * mid)

    def binarize(self, threshold=0.0, ties_above=False, one=1.0, zero=0.0):
        if isinstance(threshold, str):
            threshold = self.variables[threshold]     # logical.c:232-244
        # (+0.0 and -0.0 compare equal but have different sort keys: the key counts give the step only when
        # the zeros that tie with the threshold all lie on the side the counts put them)
        zero_ok = float(threshold) != 0.0 or (np.signbit(threshold) == bool(ties_above))
        if self._pending_sort and zero_ok and float(threshold) in getattr(self, "_sorted_known", {}) \
                and np.signbit(threshold) == self._sorted_known[float(threshold)][2]:
            # the selection pass already counted the cells below / equal to this percentile: fill only
            below, equal, _neg = self._sorted_known[float(threshold)]
            prefix, acc = [], 0
            for (lo, hi, *_r) in self.segs:
                prefix.append(acc); acc += hi - lo
            self.fill_step(prefix, below if ties_above else below + equal, one, zero)
            return
        if self._pending_sort:
            # binarizing the sorted genome = a step at cells - #(v > threshold): no sort needed
            done = C.c_int()
            check(self.lib.gdsp_sorted_binarize(self.ctx, self.layout, self._p(self._sig), float(threshold), int(ties_above),
                                                float(one), float(zero), C.byref(done)))
            if done.value:
                self._pending_sort = False
                return
        self.pointwise([self.op_binarize(threshold, ties_above, one, zero)])

    def addconst(self, value):
        if value != 0.0:                              # add.c:734
            self.pointwise([self.op_addconst(value)])

    def abs(self):
        self.pointwise([self.op_abs()])

    def clip(self, mn=None, mx=None):
        self.pointwise([self.op_clip(mn, mx)])

    def erase(self, mn=None, mx=None, keep_inside=False, zero=0.0):
        self.pointwise([self.op_erase(mn, mx, keep_inside, zero)])

    def minmax(self, stride=1, mn=-DBL_MAX, mx=DBL_MAX):
        a, b, n = C.c_double(), C.c_double(), C.c_uint64()
        check(self.lib.gdsp_minmax(self.ctx, self.layout, self._p(self.sig), int(stride), float(mn), float(mx),
                                   C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, int(n.value)

    def count_non_integer(self, limit=2.0 ** 52):
        n = C.c_uint64()
        check(self.lib.gdsp_count_non_integer(self.ctx, self.layout, self._p(self.sig), float(limit), C.byref(n)))
        return int(n.value)

    def invert(self, mid=None):
        if mid is None:                               # add.c:907-926
            lo, hi, _ = self.minmax()
            mid = (lo + hi) / 2.0
        self.pointwise([self.op_invert(mid)])

    def interval_table(self, seg, start, end, val=None):
        return IntervalTable(self, seg, start, end, val)

    def minover(self, table, infinity=DBL_MAX):
        """op_min_in_interval_apply (minmax.c:197-419): keep the trough of every interval, `infinity` elsewhere"""
        check(self.lib.gdsp_ivl_arg_extrema(self.ctx, self.layout, self._p(self.sig), table.handle, 0))
        self.pointwise([(capi.PW_IVL_KEEP_AT, infinity, 0, 0, 0, table)])

    def maxover(self, table, zero=0.0):
        """op_max_in_interval_apply (minmax.c:600-822): keep the peak of every interval, `zero` elsewhere"""
        check(self.lib.gdsp_ivl_arg_extrema(self.ctx, self.layout, self._p(self.sig), table.handle, 1))
        self.pointwise([(capi.PW_IVL_KEEP_AT, zero, 0, 0, 0, table)])

    def map_values(self, vin, vout):
        """op_map_apply (map.c:194-385) for strictly ascending breakpoints"""
        vin = np.ascontiguousarray(vin, np.float64); vout = np.ascontiguousarray(vout, np.float64)
        dp = C.POINTER(C.c_double)
        check(self.lib.gdsp_map_values(self.ctx, self.layout, self._p(self.sig), vin.ctypes.data_as(dp),
                                       vout.ctypes.data_as(dp), int(vin.size)))

    # ------------------------------------------------------------------ percentile.c
    def percentile(self, lo, hi=None, step=1.0, window=1, mn=-DBL_MAX, mx=DBL_MAX, destructive=True):
        """op_percentile_apply (percentile.c:392-751): sets self.variables['percentile<p>'].
        destructive=True also reproduces the reference's post-state for the
        all-qualifying case (globally sorted genome)."""
        hi = lo if hi is None else hi
        plo, phi = int(1000 * lo + .5), int(1000 * hi + .5)
        pstep = int(1000 * max(step, .001) + .5)
        out = {}
        if (plo, phi) in ((0, 0), (100000, 100000), (0, 100000)):      # percentile.c:434-530
            a, b, n = self.minmax(window, mn, mx)
            if n == 0:
                return out
            if plo == 0:
                out[percentile_name(0)] = a
            if phi == 100000:
                out[percentile_name(100000)] = b
            self.variables.update(out)
            return out
        ps = list(range(plo, phi + 1, pstep))
        pm = (C.c_uint32 * len(ps))(*ps)
        vals = (C.c_double * len(ps))()
        below, equal = (C.c_uint64 * len(ps))(), (C.c_uint64 * len(ps))()
        n, nnan = C.c_uint64(), C.c_uint64()
        check(self.lib.gdsp_percentiles_ranked(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), self.buffer_cells,
                                               int(window), float(mn), float(mx), pm, len(ps), vals, C.byref(n),
                                               below, equal, C.byref(nnan)))
        self.num_samples = int(n.value)
        if self.num_samples == 0:
            return out
        for i, p in enumerate(ps):
            out[percentile_name(p)] = vals[i]
        self.variables.update(out)
        self._sorted_known = {}
        if destructive:
            self._pending_sort = True                 # materialised by the next reader of self.sig
            if self.num_samples == self.cells and int(nnan.value) == 0:
                # every cell took part: #(v < value) and #(v == value) locate binarize's step in the sorted genome
                self._sorted_known = {float(vals[i]): (int(below[i]), int(equal[i]), bool(np.signbit(vals[i]))) for i in range(len(ps))}
        return out

    def fill_step(self, prefix, step, one=1.0, zero=0.0):
        """gdsp_fill_step: cell at global sorted position q (prefix[k] + offset inside piece k) = one if q >= step"""
        arr = (C.c_uint64 * self.nseg)(*[int(x) for x in prefix])
        check(self.lib.gdsp_fill_step(self.ctx, self.layout, self._p(self._sig), arr, int(step), float(one), float(zero)))
        self._pending_sort = False

    def percentile_collect(self, window=1, mn=-DBL_MAX, mx=DBL_MAX):
        """the collect pass of op_percentile_apply (percentile.c:547-580) for --window/--min/--max:
        qualifying samples to the front of the concatenated genome, the displaced values shuffled behind
        them exactly as the reference's swaps leave them -> number of qualifying samples"""
        wb = self.lib.gdsp_percentile_collect_work_bytes(self.buffer_cells)
        work = self.work(wb)
        n = C.c_uint64()
        check(self.lib.gdsp_percentile_collect(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), self.buffer_cells,
                                               self._p(work), int(window), float(mn), float(mx), C.byref(n)))
        self._swap()
        return int(n.value)

    def text_roundtrip(self):
        """percentile --preserve's write_all/read_all round trip (10 decimals), genodsp.c:1717-1775"""
        check(self.lib.gdsp_text_roundtrip(self.ctx, self.layout, self._p(self.sig), 10))

    def sort_genome(self):
        """the reference's post-percentile state for the all-qualifying case: genome globally sorted
        in chromsSorted order (percentile.c:611-651)"""
        flag = C.c_int()
        check(self.lib.gdsp_sort_genome(self.ctx, self.layout, self._p(self.sig), self._p(self.tmp), self.buffer_cells,
                                        C.byref(flag)))
        if flag.value:
            self._swap()

    def piece_sort(self, k):
        """sort owned piece k alone (the per-chromosome sorts of percentile.c:611-621); every other cell
        of the signal keeps its value"""
        lay = self._ensure_piece_layouts()[k]
        flag = C.c_int()
        check(self.lib.gdsp_sort_genome(self.ctx, lay, self._p(self.sig), self._p(self.tmp), self.buffer_cells, C.byref(flag)))
        if flag.value:
            lo, hi = self.segs[k][0], self.segs[k][1]
            self.sig[lo:hi].copy_(self.tmp[lo:hi])

    def merge_exchange(self, c, d):
        """one bubble step of the percentile operator (combine_sorted_vectors, percentile.c:820-864) on
        segments c and d, each already sorted: c keeps the smallest cells of both; returns how many moved"""
        (clo, chi), (dlo, dhi) = self.segs[c][:2], self.segs[d][:2]
        moved = C.c_uint64()
        check(self.lib.gdsp_merge_exchange(self.ctx, self._p(self.sig), self._p(self.tmp), int(clo), int(chi - clo),
                                           int(dlo), int(dhi - dlo), C.byref(moved)))
        return int(moved.value)

    # ------------------------------------------------------------------ clump.c
    def clump(self, average=0.0, length=100, relative_length=0.0, above=True, one=1.0, zero=0.0):
        wb = self.lib.gdsp_clump_work_bytes(self.buffer_cells)
        work = self.work(wb)
        check(self.lib.gdsp_clump(self.ctx, self.layout, self._p(self.sig), self.buffer_cells, self._p(work),
                                  float(average), int(length), float(relative_length), int(above),
                                  float(one), float(zero)))

    def anticlump(self, average=0.0, length=100, relative_length=0.0, one=1.0, zero=0.0):
        self.clump(average, length, relative_length, False, one, zero)

    def clump_piece(self, k, average=0.0, length=100, relative_length=0.0, above=True, one=1.0, zero=0.0):
        """clump on owned piece k alone (it must be a whole chromosome)"""
        lay = self._ensure_piece_layouts()[k]
        wb = self.lib.gdsp_clump_work_bytes(self.buffer_cells)
        work = self.work(wb)
        check(self.lib.gdsp_clump(self.ctx, lay, self._p(self.sig), self.buffer_cells, self._p(work),
                                  float(average), int(length), float(relative_length), int(above),
                                  float(one), float(zero)))

    # ------------------------------------------------------------------ output
    def runs_device(self, bufs, collapse=True, show_uncovered=0):
        """gdsp_runs into caller-provided device tensors (start i32, end i32, value f64 of equal length):
        -> (number of runs, per-segment first-run offsets); nothing is copied to the host"""
        s, e, v = bufs
        n = C.c_uint64()
        first = (C.c_uint64 * (self.nseg + 1))()
        check(self.lib.gdsp_runs(self.ctx, self.layout, self._p(self.sig), int(collapse), int(show_uncovered),
                                 self._p(s), self._p(e), self._p(v), int(s.numel()), C.byref(n), first))
        return int(n.value), list(first)

    def runs(self, collapse=True, show_uncovered=0, cap=None):
        """report_intervals run detection (genodsp.c:1589-1678) -> {chrom: (start, end, val)} numpy
        arrays, chromosome coordinates, 0-based half-open."""
        t = self.torch
        cap = int(cap) if cap else max(1024, self.cells // 16)
        while True:
            s = t.empty(cap, dtype=t.int32, device=self.device)
            e = t.empty(cap, dtype=t.int32, device=self.device)
            v = t.empty(cap, dtype=t.float64, device=self.device)
            n = C.c_uint64()
            first = (C.c_uint64 * (self.nseg + 1))()
            st = self.lib.gdsp_runs(self.ctx, self.layout, self._p(self.sig), int(collapse), int(show_uncovered),
                                    self._p(s), self._p(e), self._p(v), cap, C.byref(n), first)
            if st == capi.ERR_CAPACITY:
                cap = int(n.value) + 16
                continue
            check(st)
            break
        n = int(n.value)
        sh = s[:n].cpu().numpy().view(np.uint32); eh = e[:n].cpu().numpy().view(np.uint32); vh = v[:n].cpu().numpy()
        out = {}
        for k in range(self.nseg):
            name = self.chroms[self.seg_chrom[k]][0]
            a, b = int(first[k]), int(first[k + 1])
            prev = out.get(name)
            cur = (sh[a:b], eh[a:b], vh[a:b])
            out[name] = cur if prev is None else tuple(np.concatenate([p, c]) for p, c in zip(prev, cur))
        return out


def format_runs(genome, name, start, end, val, precision, add_start=0, add_end=0, with_value=True):
    """gdsp_format_runs on numpy run arrays -> bytes (None if the device declined: NaN/inf/huge values)"""
    t = genome.torch
    n = int(len(start))
    ds = t.from_numpy(np.ascontiguousarray(start, np.uint32).view(np.int32)).to(genome.device)
    de = t.from_numpy(np.ascontiguousarray(end, np.uint32).view(np.int32)).to(genome.device)
    dv = t.from_numpy(np.ascontiguousarray(val, np.float64)).to(genome.device)
    cap = int(genome.lib.gdsp_format_runs_max_bytes(n, name.encode())) + 16
    text = t.empty(cap, dtype=t.uint8, device=genome.device)
    nbytes, unsupported = C.c_uint64(), C.c_int()
    check(genome.lib.gdsp_format_runs(genome.ctx, genome._p(ds), genome._p(de), genome._p(dv), n, name.encode(),
                                      int(add_start), int(add_end), int(with_value), int(precision), genome._p(text), cap,
                                      C.byref(nbytes), C.byref(unsupported)))
    if unsupported.value:
        return None
    return bytes(text[:int(nbytes.value)].cpu().numpy().tobytes())


class IntervalTable:
    """sorted, disjoint interval table on the device (gdsp_ivl_table)"""

    def __init__(self, genome, seg, start, end, val=None):
        self.genome = genome
        seg = np.ascontiguousarray(seg, np.uint32); start = np.ascontiguousarray(start, np.uint32)
        end = np.ascontiguousarray(end, np.uint32)
        vp = None
        if val is not None:
            val = np.ascontiguousarray(val, np.float64); vp = val.ctypes.data_as(C.POINTER(C.c_double))
        h = C.c_void_p()
        u32p = C.POINTER(C.c_uint32)
        check(genome.lib.gdsp_ivl_table_create(genome.ctx, genome.layout, seg.ctypes.data_as(u32p),
                                               start.ctypes.data_as(u32p), end.ctypes.data_as(u32p), vp,
                                               seg.size, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.genome.lib.gdsp_ivl_table_destroy(self.handle); self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
