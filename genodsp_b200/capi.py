"""ctypes binding of the C-ABI in include/gdsp_b200.h (libgdsp_b200.so).

There is no CPU path: importing this module without the built CUDA library,
or creating a context without a B200-class device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GDSP_LIB_PATH: a differently built copy of the library (kernel experiments: scripts/build_variant.sh)
LIB_PATH = os.environ.get("GDSP_LIB_PATH") or os.path.join(_HERE, "lib", "libgdsp_b200.so")


class GdspError(RuntimeError):
    pass


class Seg(C.Structure):
    _fields_ = [("lo", C.c_uint64), ("hi", C.c_uint64), ("dlo", C.c_uint64), ("dhi", C.c_uint64),
                ("pos0", C.c_uint32), ("chrom_len", C.c_uint32)]


class Halo(C.Structure):
    _fields_ = [("peer", C.c_int32), ("send_lo", C.c_uint64), ("send_hi", C.c_uint64), ("recv_lo", C.c_uint64), ("recv_hi", C.c_uint64)]


class PwOp(C.Structure):
    _fields_ = [("code", C.c_int32), ("flags", C.c_uint32), ("a", C.c_double), ("b", C.c_double),
                ("c", C.c_double), ("table", C.c_void_p)]


# gdsp_pw_code
PW_BINARIZE_GT, PW_BINARIZE_GE, PW_ADDCONST, PW_ABS, PW_CLIP_MIN, PW_CLIP_MAX, PW_CLIP_BOTH, PW_ERASE, \
    PW_INVERT, PW_NONZERO_TO_ONE, PW_IVL_ADD, PW_IVL_SUB, PW_IVL_MUL, PW_IVL_DIV, PW_IVL_SET, \
    PW_IVL_SET_OUTSIDE, PW_IVL_ASSIGN, PW_IVL_MIN, PW_IVL_MAX, PW_IVL_KEEP_AT, PW_IVL_ACCUM_CLEAR = range(1, 22)
PW_ERASE_HAVE_MIN, PW_ERASE_HAVE_MAX, PW_ERASE_KEEP_INSIDE = 1, 2, 4
ACC_I32, ACC_F64 = 0, 1
MORPH_CLOSE, MORPH_OPEN, MORPH_DILATE, MORPH_ERODE = 0, 1, 2, 3
ERR_CAPACITY = -4
ALIGN = 64

_vp, _u32, _u64, _i, _d, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double, C.c_size_t
_u32p, _u64p, _dp = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_double)

# name -> (restype, argtypes); every function include/gdsp_b200.h declares
SIGNATURES = {
    "gdsp_ctx_create": (_i, [_i, _vp, C.POINTER(_vp)]),
    "gdsp_ctx_destroy": (None, [_vp]),
    "gdsp_ctx_set_stream": (_i, [_vp, _vp]),
    "gdsp_sync": (_i, [_vp]),
    "gdsp_ctx_set_exact_order": (_i, [_vp, _i]),
    "gdsp_ctx_get_exact_order": (_i, [_vp]),
    "gdsp_ctx_set_smooth_direct": (_i, [_vp, _i]),
    "gdsp_malloc_host": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "gdsp_free_host": (_i, [_vp]),
    "gdsp_last_error": (C.c_char_p, []),
    "gdsp_launch_count": (C.c_uint64, []),
    "gdsp_version": (C.c_char_p, []),
    "gdsp_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)]),
    "gdsp_malloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "gdsp_free": (_i, [_vp, _vp]),
    "gdsp_host_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "gdsp_host_free": (_i, [_vp, _vp]),
    "gdsp_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "gdsp_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "gdsp_d2d": (_i, [_vp, _vp, _vp, _sz]),
    "gdsp_timer_start": (_i, [_vp]),
    "gdsp_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "gdsp_layout_pack": (_i, [_u32p, _i, C.POINTER(Seg), _u64p]),
    "gdsp_layout_create": (_i, [_vp, C.POINTER(Seg), _i, C.POINTER(_vp)]),
    "gdsp_layout_destroy": (None, [_vp]),
    "gdsp_layout_nseg": (_i, [_vp]),
    "gdsp_layout_segs": (C.POINTER(Seg), [_vp]),
    "gdsp_layout_cells": (_u64, [_vp]),
    "gdsp_fill": (_i, [_vp, _vp, _vp, _d]),
    "gdsp_accumulate_work_bytes": (_sz, [_vp, _u64, _i]),
    "gdsp_accumulate_dev": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _i, _i]),
    "gdsp_accumulate_host": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp, _u64, _i, _i]),
    "gdsp_block_sum": (_i, [_vp, _vp, _vp, _u32, _i, _d, _i, _d]),
    "gdsp_sliding_sum": (_i, [_vp, _vp, _vp, _vp, _u32, _d]),
    "gdsp_smooth": (_i, [_vp, _vp, _vp, _vp, _u32, _dp]),
    "gdsp_cumulative_sum": (_i, [_vp, _vp, _vp, _vp]),
    "gdsp_local_extrema": (_i, [_vp, _vp, _vp, _vp, _u32, _i, _d]),
    "gdsp_best_extrema": (_i, [_vp, _vp, _vp, _vp, _u32, _i]),
    "gdsp_morph_work_bytes": (_sz, [_u64]),
    "gdsp_morphology": (_i, [_vp, _vp, _vp, _u64, _vp, _i, _d, _u32, _u32, _d, _d, _d]),
    "gdsp_ivl_table_create": (_i, [_vp, _vp, _u32p, _u32p, _u32p, _dp, _u64, C.POINTER(_vp)]),
    "gdsp_ivl_table_destroy": (None, [_vp]),
    "gdsp_pointwise": (_i, [_vp, _vp, _vp, _vp, C.POINTER(PwOp), _i]),
    "gdsp_minmax": (_i, [_vp, _vp, _vp, _u32, _d, _d, _dp, _dp, _u64p]),
    "gdsp_count_non_integer": (_i, [_vp, _vp, _vp, _d, _u64p]),
    "gdsp_percentiles": (_i, [_vp, _vp, _vp, _vp, _u64, _u32, _d, _d, _u32p, _i, _dp, _u64p]),
    "gdsp_percentiles_ranked": (_i, [_vp, _vp, _vp, _vp, _u64, _u32, _d, _d, _u32p, _i, _dp, _u64p, _u64p, _u64p, _u64p]),
    "gdsp_pct_count_nan": (_i, [_vp, _vp, _vp, _u32, _d, _d, _u64p, _i, C.POINTER(C.c_uint8), _u64p, _vp, _u64, _u64p, _u64p]),
    "gdsp_percentile_collect_work_bytes": (_sz, [_u64]),
    "gdsp_percentile_collect": (_i, [_vp, _vp, _vp, _vp, _u64, _vp, _u32, _d, _d, _u64p]),
    "gdsp_equal_range": (_i, [_vp, _vp, _u64, _d, _u64p, _u64p]),
    "gdsp_sort_genome": (_i, [_vp, _vp, _vp, _vp, _u64, C.POINTER(_i)]),
    "gdsp_merge_exchange": (_i, [_vp, _vp, _vp, _u64, _u64, _u64, _u64, _u64p]),
    "gdsp_sorted_binarize": (_i, [_vp, _vp, _vp, _d, _i, _d, _d, C.POINTER(_i)]),
    "gdsp_fill_step": (_i, [_vp, _vp, _vp, _u64p, _u64, _d, _d]),
    "gdsp_ivl_arg_extrema": (_i, [_vp, _vp, _vp, _vp, _i]),
    "gdsp_map_values": (_i, [_vp, _vp, _vp, _dp, _dp, _i]),
    "gdsp_format_runs_max_bytes": (C.c_size_t, [_u64, C.c_char_p]),
    "gdsp_format_runs": (_i, [_vp, _vp, _vp, _vp, _u64, C.c_char_p, _u32, _u32, _i, _i, _vp, _u64, _u64p, C.POINTER(_i)]),
    "gdsp_text_roundtrip": (_i, [_vp, _vp, _vp, _i]),
    "gdsp_pct_sample": (_i, [_vp, _vp, _vp, _u32, _d, _d, _u64, _u64, _u32, _u64, _vp, _u32p, _u64p]),
    "gdsp_sort_array": (_i, [_vp, _vp, _vp, _u64, C.POINTER(_i)]),
    "gdsp_pct_count": (_i, [_vp, _vp, _vp, _u32, _d, _d, _u64p, _i, C.POINTER(C.c_uint8), _u64p, _vp, _u64, _u64p]),
    "gdsp_clump_work_bytes": (_sz, [_u64]),
    "gdsp_clump": (_i, [_vp, _vp, _vp, _u64, _vp, _d, _u32, _d, _i, _d, _d]),
    "gdsp_clump_slab_create": (_i, [_vp, _vp, _u64, _vp, _d, _u32, _d, _i, _d, _d, C.POINTER(_vp)]),
    "gdsp_clump_slab_reduce": (_i, [_vp, _vp, _dp]),
    "gdsp_clump_slab_mark": (_i, [_vp, _vp, _dp, _dp]),
    "gdsp_clump_slab_trim": (_i, [_vp, _vp, _dp, C.POINTER(_i)]),
    "gdsp_clump_slab_emit": (_i, [_vp, _vp, C.POINTER(C.c_ubyte), C.POINTER(_i)]),
    "gdsp_clump_slab_destroy": (None, [_vp]),
    "gdsp_comm_unique_id": (_i, [C.POINTER(C.c_ubyte)]),
    "gdsp_comm_create": (_i, [_vp, C.POINTER(C.c_ubyte), _i, _i, C.POINTER(_vp)]),
    "gdsp_comm_create_all": (_i, [C.POINTER(_vp), _i, C.POINTER(_vp)]),
    "gdsp_comm_destroy": (None, [_vp]),
    "gdsp_comm_rank": (_i, [_vp]),
    "gdsp_comm_size": (_i, [_vp]),
    "gdsp_comm_exchange_halos": (_i, [_vp, _vp, C.POINTER(Halo), _i]),
    "gdsp_comm_exchange_halos_all": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.POINTER(Halo)), C.POINTER(_i), _i]),
    "gdsp_comm_allreduce_sum_u64": (_i, [_vp, _u64p, _i]),
    "gdsp_comm_allgather_f64": (_i, [_vp, _dp, _i, _dp]),
    "gdsp_comm_allgather_dev": (_i, [_vp, _vp, _u64, _vp]),
    "gdsp_comm_broadcast_f64": (_i, [_vp, _dp, _i, _i]),
    "gdsp_runs": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _u64, _u64p, _u64p]),
}

_lib = None


def load():
    """Load libgdsp_b200.so (built by `make -C genodsp_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdspError("%s is missing: build it with `make -C genodsp_b200/csrc` "
                        "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise GdspError("gdsp error %d: %s" % (status, load().gdsp_last_error().decode()))
