"""ctypes bindings for the two CHECKERS used by the tests (never by the product):

* ``oracle/libgdsp_oracle.so``      -- our plain-C restatement (oracle/gdsp_oracle.c)
* ``oracle/_ref/libgenodsp_ref.so`` -- the unmodified reference, built from
  /root/reference by oracle/Makefile together with oracle/ref_shim.c
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgdsp_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgenodsp_ref.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "genodsp")

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _u32p(a):
    return a.ctypes.data_as(c_u32p)


def build_checkers():
    subprocess.run(["make", "-C", ORACLE_DIR, "all"], check=True, capture_output=True)


class Oracle:
    """Array-level calls into the plain-C restatement."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        self.lib = L = C.CDLL(ORACLE_SO)
        d, u32, u64, i = C.c_double, C.c_uint32, C.c_uint64, C.c_int
        sig = {
            "gdo_accumulate": [c_dp, u32, c_u32p, c_u32p, c_dp, u64, i, i, d],
            "gdo_block_sum": [c_dp, u32, u32, d, i, d],
            "gdo_sliding_sum": [c_dp, u32, u32, d],
            "gdo_hann_taps": [c_dp, u32],
            "gdo_smooth": [c_dp, u32, u32],
            "gdo_cumulative": [c_dp, u32],
            "gdo_text_roundtrip10": [c_dp, u32],
            "gdo_over_intervals": [c_dp, u32, c_u32p, c_u32p, u32, C.c_int, C.c_double],
            "gdo_with_intervals": [c_dp, u32, c_u32p, c_u32p, c_dp, u32, C.c_int],
            "gdo_map": [c_dp, u32, c_dp, c_dp, u32],
            "gdo_local_extrema": [c_dp, u32, u32, i, d],
            "gdo_best_extrema": [c_dp, u32, u32, i],
            "gdo_close": [c_dp, u32, d, d, d, d],
            "gdo_open": [c_dp, u32, d, d, d, d],
            "gdo_dilate": [c_dp, u32, u32, u32, d, d, d],
            "gdo_erode": [c_dp, u32, u32, u32, d, d, d],
            "gdo_binarize": [c_dp, u32, d, i, d, d],
            "gdo_addconst": [c_dp, u32, d],
            "gdo_abs": [c_dp, u32],
            "gdo_clip": [c_dp, u32, i, d, i, d],
            "gdo_erase": [c_dp, u32, i, d, i, d, i, d],
            "gdo_invert": [c_dp, u32, d],
            "gdo_minmax": [c_dp, u32, c_dp, c_dp],
            "gdo_logical_prep": [c_dp, u32],
            "gdo_add_intervals": [c_dp, u32, c_u32p, c_u32p, c_dp, u64, d],
            "gdo_mask_intervals": [c_dp, u32, c_u32p, c_u32p, u64, d],
            "gdo_or_intervals": [c_dp, u32, c_u32p, c_u32p, c_dp, u64],
            "gdo_sorted_intervals": [c_dp, u32, c_u32p, c_u32p, c_dp, u64, i, d],
            "gdo_clump": [c_dp, u32, d, u32, i, d, d],
            "gdo_sort": [c_dp, u64],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = None
        L.gdo_percentile_collect.argtypes = [c_dp, u32, u32, d, d, c_dp]
        L.gdo_percentile_collect.restype = u64
        L.gdo_percentile_rank.argtypes = [u64, u32]
        L.gdo_percentile_rank.restype = u64
        L.gdo_runs.argtypes = [c_dp, u32, i, i, c_u32p, c_u32p, c_dp, u64]
        L.gdo_runs.restype = u64

    # each wrapper works on a float64 numpy array in place and returns it
    def _v(self, v):
        assert v.dtype == np.float64 and v.flags.c_contiguous
        return _dp(v), v.size

    def accumulate(self, v, s, e, val=None, overlap=0, clear=False, missing=0.0):
        s = np.ascontiguousarray(s, np.uint32)
        e = np.ascontiguousarray(e, np.uint32)
        vp = None if val is None else _dp(np.ascontiguousarray(val, np.float64))
        if val is not None:
            val = np.ascontiguousarray(val, np.float64)
            vp = _dp(val)
        self.lib.gdo_accumulate(*self._v(v), _u32p(s), _u32p(e), vp, s.size, overlap, int(clear), missing)
        return v

    def block_sum(self, v, W, denom=1.0, actual=False, zero=0.0):
        self.lib.gdo_block_sum(*self._v(v), W, denom, int(actual), zero); return v

    def sliding_sum(self, v, W, denom=1.0):
        self.lib.gdo_sliding_sum(*self._v(v), W, denom); return v

    def hann_taps(self, W):
        w = np.empty(W, np.float64)
        self.lib.gdo_hann_taps(_dp(w), W); return w

    def smooth(self, v, W):
        self.lib.gdo_smooth(*self._v(v), W); return v

    def cumulative(self, v):
        self.lib.gdo_cumulative(*self._v(v)); return v

    def over_intervals(self, v, s, e, want_max, fill):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32)
        self.lib.gdo_over_intervals(*self._v(v), s.ctypes.data_as(c_u32p), e.ctypes.data_as(c_u32p), s.size, int(want_max), fill)
        return v

    def with_intervals(self, v, s, e, val, want_max):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32); val = np.ascontiguousarray(val, np.float64)
        self.lib.gdo_with_intervals(*self._v(v), s.ctypes.data_as(c_u32p), e.ctypes.data_as(c_u32p), val.ctypes.data_as(c_dp), s.size, int(want_max))
        return v

    def map_values(self, v, vin, vout):
        vin = np.ascontiguousarray(vin, np.float64); vout = np.ascontiguousarray(vout, np.float64)
        self.lib.gdo_map(*self._v(v), vin.ctypes.data_as(c_dp), vout.ctypes.data_as(c_dp), vin.size)
        return v

    def text_roundtrip10(self, v):
        self.lib.gdo_text_roundtrip10(*self._v(v)); return v

    def local_extrema(self, v, N, want_max=True, fill=0.0):
        self.lib.gdo_local_extrema(*self._v(v), N, int(want_max), fill); return v

    def best_extrema(self, v, W, want_max=True):
        self.lib.gdo_best_extrema(*self._v(v), W, int(want_max)); return v

    def close(self, v, L, T=0.0, one=1.0, zero=0.0):
        self.lib.gdo_close(*self._v(v), float(L), T, one, zero); return v

    def open(self, v, L, T=0.0, one=1.0, zero=0.0):
        self.lib.gdo_open(*self._v(v), float(L), T, one, zero); return v

    def dilate(self, v, left, right, T=0.0, one=1.0, zero=0.0):
        self.lib.gdo_dilate(*self._v(v), left, right, T, one, zero); return v

    def erode(self, v, left, right, T=0.0, one=1.0, zero=0.0):
        self.lib.gdo_erode(*self._v(v), left, right, T, one, zero); return v

    def binarize(self, v, T=0.0, ties_above=False, one=1.0, zero=0.0):
        self.lib.gdo_binarize(*self._v(v), T, int(ties_above), one, zero); return v

    def addconst(self, v, c):
        self.lib.gdo_addconst(*self._v(v), c); return v

    def abs(self, v):
        self.lib.gdo_abs(*self._v(v)); return v

    def clip(self, v, mn=None, mx=None):
        self.lib.gdo_clip(*self._v(v), int(mn is not None), mn or 0.0, int(mx is not None), mx or 0.0); return v

    def erase(self, v, mn=None, mx=None, keep_inside=False, zero=0.0):
        self.lib.gdo_erase(*self._v(v), int(mn is not None), mn or 0.0, int(mx is not None), mx or 0.0,
                           int(keep_inside), zero); return v

    def invert(self, v, mid):
        self.lib.gdo_invert(*self._v(v), mid); return v

    def logical_prep(self, v):
        self.lib.gdo_logical_prep(*self._v(v)); return v

    def add_intervals(self, v, s, e, val, sign=1.0):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32)
        val = np.ascontiguousarray(val, np.float64)
        self.lib.gdo_add_intervals(*self._v(v), _u32p(s), _u32p(e), _dp(val), s.size, sign); return v

    def mask_intervals(self, v, s, e, mask_val=0.0):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32)
        self.lib.gdo_mask_intervals(*self._v(v), _u32p(s), _u32p(e), s.size, mask_val); return v

    def or_intervals(self, v, s, e, val):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32)
        val = np.ascontiguousarray(val, np.float64)
        self.lib.gdo_or_intervals(*self._v(v), _u32p(s), _u32p(e), _dp(val), s.size); return v

    def sorted_intervals(self, v, s, e, val, kind, aux=0.0):
        s = np.ascontiguousarray(s, np.uint32); e = np.ascontiguousarray(e, np.uint32)
        val = np.ascontiguousarray(val, np.float64)
        self.lib.gdo_sorted_intervals(*self._v(v), _u32p(s), _u32p(e), _dp(val), s.size, kind, aux); return v

    def clump(self, v, T, min_length, above=True, one=1.0, zero=0.0):
        self.lib.gdo_clump(*self._v(v), T, min_length, int(above), one, zero); return v

    def percentile_collect(self, v, W=1, mn=-np.finfo(np.float64).max, mx=np.finfo(np.float64).max):
        out = np.empty(v.size, np.float64)
        c = self.lib.gdo_percentile_collect(_dp(v), v.size, W, mn, mx, _dp(out))
        return out[:c]

    def percentile_rank(self, n, p_milli):
        return int(self.lib.gdo_percentile_rank(n, p_milli))

    def sort(self, v):
        self.lib.gdo_sort(_dp(v), v.size); return v

    def runs(self, v, collapse=True, show_uncovered=0):
        cap = v.size
        rs = np.empty(cap, np.uint32); re = np.empty(cap, np.uint32); rv = np.empty(cap, np.float64)
        r = self.lib.gdo_runs(_dp(v), v.size, int(collapse), show_uncovered, _u32p(rs), _u32p(re), _dp(rv), cap)
        return rs[:r].copy(), re[:r].copy(), rv[:r].copy()


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


class RefGenome:
    """The reference itself, in-process (oracle/ref_shim.c).  One instance at a
    time: the reference keeps its state in C globals."""

    def __init__(self, chroms):
        """chroms: list of (name, length) in file order."""
        self.lib = L = C.CDLL(REF_SO)
        L.refshim_add_chrom.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32]
        L.refshim_vector.argtypes = [C.c_char_p]
        L.refshim_vector.restype = c_dp
        L.refshim_apply.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.refshim_apply.restype = C.c_double
        L.refshim_read_intervals.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.refshim_report_intervals.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.refshim_get_global.argtypes = [C.c_char_p, c_dp]
        L.refshim_set_global.argtypes = [C.c_char_p, C.c_double]
        L.refshim_sorted_name.restype = C.c_char_p
        L.refshim_sorted_name.argtypes = [C.c_int]
        L.refshim_reset()
        self.chroms = list(chroms)
        for name, length in self.chroms:
            assert L.refshim_add_chrom(name.encode(), 0, length)
        L.refshim_finalize()
        self.vec = {}
        for name, length in self.chroms:
            p = L.refshim_vector(name.encode())
            self.vec[name] = np.ctypeslib.as_array(p, shape=(length,))

    def sorted_names(self):
        return [self.lib.refshim_sorted_name(i).decode() for i in range(len(self.chroms))]

    def apply(self, *words):
        """apply("smooth", "--window=101") -> seconds spent in the reference's apply"""
        argv = (C.c_char_p * len(words))(*[C.create_string_buffer(w.encode()).raw.rstrip(b"\0") for w in words])
        # the reference's parsers write into argv strings: give them mutable buffers
        bufs = [C.create_string_buffer(w.encode()) for w in words]
        argv = (C.c_char_p * len(words))(*[C.cast(b, C.c_char_p) for b in bufs])
        t = self.lib.refshim_apply(len(words), argv)
        assert t >= 0, "unknown operator %r" % (words[0],)
        return t

    def read_intervals(self, path, val_col=3, origin_one=False, overlap=0, clear=False, missing=0.0):
        assert self.lib.refshim_read_intervals(path.encode(), val_col, int(origin_one), overlap, int(clear), missing)

    def report(self, path, precision=0, no_values=False, collapse=True, show_uncovered=0, origin_one=False):
        assert self.lib.refshim_report_intervals(path.encode(), precision, int(no_values), int(collapse),
                                                 show_uncovered, int(origin_one))

    def get_global(self, name):
        v = C.c_double()
        ok = self.lib.refshim_get_global(name.encode(), C.byref(v))
        return v.value if ok else None

    def set_global(self, name, v):
        self.lib.refshim_set_global(name.encode(), v)

    def close(self):
        self.vec = {}
        self.lib.refshim_reset()


def run_ref_cli(args, stdin_path=None, cwd=None):
    """Run the unmodified reference binary; returns (returncode, stdout, stderr) as bytes."""
    fin = open(stdin_path, "rb") if stdin_path else subprocess.DEVNULL
    try:
        p = subprocess.run([REF_BIN] + list(args), stdin=fin, capture_output=True, cwd=cwd)
    finally:
        if stdin_path:
            fin.close()
    return p.returncode, p.stdout, p.stderr
