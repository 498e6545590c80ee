/* gdsp_b200.h -- C-ABI of the B200 (sm_100a) per-base operator library.
 *
 * This is the drop-in boundary for genodsp's hot path: every entry point
 * replaces the per-base loop of one reference function (cited as file:line into
 * rsharris/genodsp 0.0.10).  Plain C types only: device pointers are `double*`
 * etc. obtained from gdsp_malloc() or from any CUDA allocator (cudaMalloc,
 * torch) on the same device; no C++/torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative gdsp_status on failure;
 *     gdsp_last_error() returns the message of the calling thread's last failure
 *   - all work is issued on the context's stream and is asynchronous unless
 *     the function returns a host value (then it synchronises that stream)
 *   - there is NO CPU fallback: without a usable CUDA device every call fails
 *
 * Signal layout ("track"): the chromosomes of a genome live in ONE device
 * buffer of doubles; each chromosome (or the slab piece of it that this GPU
 * owns) is a segment described by gdsp_seg.  Kernels compute the cells
 * [lo,hi) of every segment and may read [dlo,dhi) -- on a single GPU the two
 * ranges coincide; on a slab-sharded run [dlo,dhi) additionally covers halo
 * cells received from the neighbouring GPU.  Anything outside [dlo,dhi) is
 * treated exactly as the reference treats positions beyond the ends of a
 * chromosome vector.
 */
#ifndef GDSP_B200_H
#define GDSP_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gdsp_ctx    gdsp_ctx;     /* device + stream + workspace          */
typedef struct gdsp_layout gdsp_layout;  /* device copy of a segment table       */

typedef enum gdsp_status
	{
	GDSP_OK            =  0,
	GDSP_ERR_CUDA      = -1,   /* a CUDA runtime call or kernel failed          */
	GDSP_ERR_ARG       = -2,   /* invalid argument                              */
	GDSP_ERR_NOMEM     = -3,   /* device or host allocation failed              */
	GDSP_ERR_CAPACITY  = -4,   /* caller-provided output buffer too small       */
	GDSP_ERR_NODEVICE  = -5    /* no CUDA device / not an sm_100 class device   */
	} gdsp_status;

/* one segment = one chromosome vector (reference: spec, genodsp_interface.h:37-49)
 * or the part of it owned by this GPU.  All ranges are element indices into the
 * signal buffer; lo must be a multiple of GDSP_ALIGN. */
typedef struct gdsp_seg
	{
	uint64_t lo, hi;        /* owned cells                                       */
	uint64_t dlo, dhi;      /* readable cells of the same chromosome (>= owned)  */
	uint32_t pos0;          /* chromosome coordinate (0-based index into the
	                           reference's valVector) of cell lo                 */
	uint32_t chrom_len;     /* spec.length of the whole chromosome               */
	} gdsp_seg;

#define GDSP_ALIGN 64       /* segment starts are multiples of 64 cells (512 B)  */

/* ---- context, memory, layout ------------------------------------------- */

/* `stream` is a cudaStream_t: NULL is the legacy default stream (what
 * torch.cuda.current_stream() reports as 0); GDSP_STREAM_PRIVATE asks the
 * library to create and own a non-blocking stream. */
#define GDSP_STREAM_PRIVATE ((void*) (intptr_t) -1)
int  gdsp_ctx_create   (int device, void* stream, gdsp_ctx** out);
void gdsp_ctx_destroy  (gdsp_ctx* ctx);
int  gdsp_ctx_set_stream (gdsp_ctx* ctx, void* stream);
int  gdsp_sync         (gdsp_ctx* ctx);
/* Exact-order mode (off by default; `--exact-order` in the CLI).  While on, gdsp_sliding_sum,
 * gdsp_cumulative_sum and gdsp_clump evaluate their running sums in the reference's own sequential order
 * (sum.c:438-455, sum.c:786-790, clump.c:600): bit-identical to the reference on every input, general reals,
 * inf and NaN included, one dependent FP64 add per cell and chromosome (seconds, not milliseconds, on a
 * human genome).  Whole chromosomes only. */
int  gdsp_ctx_set_exact_order (gdsp_ctx* ctx, int on);
int  gdsp_ctx_get_exact_order (const gdsp_ctx* ctx);
/* gdsp_smooth picks between two bit-identical kernels: the direct FIR (2W FP64 instructions per base) and, for
 * the bit-symmetric windows the reference builds (sum.c:641), one that computes the product of a tap pair once
 * (3(W-1)/2 + 2 per base).  `on` forces the direct FIR -- for A/B timing and parity tests of both. */
int  gdsp_ctx_set_smooth_direct (gdsp_ctx* ctx, int on);
/* page-locked host memory (device<->host copies from it run at full PCIe speed) */
int  gdsp_malloc_host (size_t bytes, void** out);
int  gdsp_free_host   (void* p);
const char* gdsp_last_error (void);
/* number of kernels this library has launched in the calling process (all contexts) */
uint64_t gdsp_launch_count (void);
const char* gdsp_version    (void);
int  gdsp_device_info  (gdsp_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                        size_t* free_bytes, size_t* total_bytes);

int  gdsp_malloc       (gdsp_ctx* ctx, size_t bytes, void** dptr);
int  gdsp_free         (gdsp_ctx* ctx, void* dptr);
int  gdsp_host_alloc   (gdsp_ctx* ctx, size_t bytes, void** hptr);   /* pinned */
int  gdsp_host_free    (gdsp_ctx* ctx, void* hptr);
int  gdsp_h2d          (gdsp_ctx* ctx, void* dst, const void* src, size_t bytes);
int  gdsp_d2h          (gdsp_ctx* ctx, void* dst, const void* src, size_t bytes);
int  gdsp_d2d          (gdsp_ctx* ctx, void* dst, const void* src, size_t bytes);

/* timing on the context's stream (CUDA events) */
int  gdsp_timer_start  (gdsp_ctx* ctx);
int  gdsp_timer_stop   (gdsp_ctx* ctx, float* ms);      /* synchronises */

/* Pack `nseg` whole chromosomes of the given lengths back to back (each start
 * rounded up to GDSP_ALIGN).  segs_out (optional, nseg entries) receives the
 * table, *total_cells the buffer size the caller must allocate (already padded
 * so that vector loads past a segment end stay inside the buffer).
 * Reference: the per-chromosome callocs of main, genodsp.c:865-878. */
int  gdsp_layout_pack   (const uint32_t* chrom_len, int nseg,
                         gdsp_seg* segs_out, uint64_t* total_cells);
int  gdsp_layout_create (gdsp_ctx* ctx, const gdsp_seg* segs, int nseg, gdsp_layout** out);
void gdsp_layout_destroy(gdsp_layout* lay);
int  gdsp_layout_nseg   (const gdsp_layout* lay);
const gdsp_seg* gdsp_layout_segs (const gdsp_layout* lay);
uint64_t gdsp_layout_cells (const gdsp_layout* lay);      /* sum of (hi-lo)      */

/* fill the owned cells of every segment (pads untouched).
 * Reference: zero fill genodsp.c:876-877; clear-to-missing genodsp.c:1218-1233 */
int  gdsp_fill (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig, double value);

/* ---- interval accumulation (input stage) ---------------------------------
 * Reference: read_intervals accumulate loops, genodsp.c:1307-1330
 * (overlapOp sum).  Intervals are SoA arrays: segment index into the layout,
 * chromosome start/end (0-based half-open, already origin-shifted, clipped and
 * validated by the host reader), optional value (NULL => 1.0 each, --novalue).
 * `*_host` variants take HOST arrays and stream them through pinned staging
 * buffers; `*_dev` variants take device arrays.
 *
 *   mode GDSP_ACC_I32: int32 difference array + segmented scan; exact when all
 *        values are integers and every partial sum fits in int32
 *   mode GDSP_ACC_F64: fp64 difference array; exact when every partial sum is
 *        exactly representable (integer / dyadic values), else within rounding
 *
 * sig is overwritten when add_to_existing==0, else the depth is added to it.
 * `work` is a caller-provided device buffer of gdsp_accumulate_work_bytes(). */
#define GDSP_ACC_I32 0
#define GDSP_ACC_F64 1
size_t gdsp_accumulate_work_bytes (const gdsp_layout* lay, uint64_t buffer_cells, int mode);
int  gdsp_accumulate_dev  (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                           uint64_t buffer_cells, void* work,
                           const uint32_t* d_seg, const uint32_t* d_start,
                           const uint32_t* d_end, const double* d_val,
                           uint64_t n_intervals, int mode, int add_to_existing);
int  gdsp_accumulate_host (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                           uint64_t buffer_cells, void* work,
                           const uint32_t* h_seg, const uint32_t* h_start,
                           const uint32_t* h_end, const double* h_val,
                           uint64_t n_intervals, int mode, int add_to_existing);

/* ---- windowed sums -------------------------------------------------------- */
/* op_window_sum_apply, sum.c:211-252 (in place) */
int  gdsp_block_sum   (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                       uint32_t window, int window_is_chromosome,
                       double denom, int denom_is_actual, double zero_val);
/* op_sliding_sum_apply, sum.c:420-463 (out of place: in -> out) */
int  gdsp_sliding_sum (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in,
                       double* out, uint32_t window, double denom);
/* op_smooth_apply, sum.c:616-676; taps = `window` host doubles computed by the
 * caller with the host libm exactly as sum.c:634-645 does (out of place) */
int  gdsp_smooth      (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in,
                       double* out, uint32_t window, const double* h_taps);
/* op_cumulative_sum_apply, sum.c:776-792 (in place allowed: out may equal in) */
int  gdsp_cumulative_sum (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in, double* out);

/* ---- sliding extrema ------------------------------------------------------ */
/* op_local_maxima_apply minmax.c:1183-1227 (want_max=1, fill=zeroVal) and
 * op_local_minima_apply minmax.c:981-1022 (want_max=0, fill=infinityVal) */
int  gdsp_local_extrema (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in,
                         double* out, uint32_t neighborhood, int want_max, double fill);
/* op_best_local_max_apply minmax.c:1616-1721 / op_best_local_min_apply :1369-1474 */
int  gdsp_best_extrema  (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in,
                         double* out, uint32_t window, int want_max);

/* ---- run-length morphology (in place) --------------------------------------
 * op_close_apply morphology.c:231-319, op_open_apply :529-605,
 * op_dilate_apply :882-1072, op_erode_apply :1331-1454.
 * `work` = device buffer of gdsp_morph_work_bytes(buffer_cells). */
#define GDSP_MORPH_CLOSE  0
#define GDSP_MORPH_OPEN   1
#define GDSP_MORPH_DILATE 2
#define GDSP_MORPH_ERODE  3
size_t gdsp_morph_work_bytes (uint64_t buffer_cells);
int  gdsp_morphology (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                      uint64_t buffer_cells, void* work, int kind,
                      double length, uint32_t left, uint32_t right,
                      double threshold, double one_val, double zero_val);

/* ---- fused pointwise programs ---------------------------------------------
 * One kernel applies up to GDSP_MAX_POINTWISE consecutive pointwise operators
 * per cell, so a chain costs one read + one write of the signal. */
typedef enum gdsp_pw_code
	{
	GDSP_PW_BINARIZE_GT = 1,  /* a=threshold b=one c=zero   logical.c:259-262 */
	GDSP_PW_BINARIZE_GE,      /*                            logical.c:253-256 */
	GDSP_PW_ADDCONST,         /* a=constant                 add.c:736-739     */
	GDSP_PW_ABS,              /*                            add.c:1046-1047   */
	GDSP_PW_CLIP_MIN,         /* a=min                      mask.c:896-899    */
	GDSP_PW_CLIP_MAX,         /* a=max                      mask.c:901-904    */
	GDSP_PW_CLIP_BOTH,        /* a=min b=max                mask.c:906-912    */
	GDSP_PW_ERASE,            /* a=min b=max c=zero, flags  mask.c:1185-1227  */
	GDSP_PW_INVERT,           /* a=2*mid                    add.c:936         */
	GDSP_PW_NONZERO_TO_ONE,   /* or/and first pass          logical.c:466-473 */
	GDSP_PW_IVL_ADD,          /* interval table, v+=val     add.c:280-281     */
	GDSP_PW_IVL_SUB,          /*                 v-=val     add.c:573-574     */
	GDSP_PW_IVL_MUL,          /* inside v*=val, gap a(=0)   multiply.c:193-393*/
	GDSP_PW_IVL_DIV,          /* inside v/=val, gap +-a     multiply.c:586-787*/
	GDSP_PW_IVL_SET,          /* inside v=a                 mask.c:295-296,
	                                                        logical.c:or      */
	GDSP_PW_IVL_SET_OUTSIDE,  /* gap v=a                    mask.c:483-668,
	                                                        logical.c:and     */
	GDSP_PW_IVL_ASSIGN,       /* inside v=val (input --overlap=min/max after the
	                             host reduced the overlaps, genodsp.c:1307-1322) */
	GDSP_PW_IVL_MIN,          /* inside v=min(v,val)        minmax.c:1979-1982 (minwith; the host
	                             reduces overlapping intervals to their minimum first) */
	GDSP_PW_IVL_MAX,          /* inside v=max(v,val)        minmax.c:2265-2268 (maxwith) */
	GDSP_PW_IVL_KEEP_AT,      /* v=a everywhere except at the one cell of every interval that
	                             gdsp_ivl_arg_extrema chose  minmax.c:345-348, :748-751 (minover/maxover) */
	GDSP_PW_IVL_ACCUM_CLEAR   /* inside v = (v==a) ? val : v+val   the `clear` accumulate of
	                             read_intervals, genodsp.c:1325-1329 (a = the missing value) */
	} gdsp_pw_code;

#define GDSP_PW_ERASE_HAVE_MIN    1u
#define GDSP_PW_ERASE_HAVE_MAX    2u
#define GDSP_PW_ERASE_KEEP_INSIDE 4u

/* sorted, non-overlapping interval table in buffer coordinates (cell indices
 * of the signal buffer), on the device; built by gdsp_ivl_table_create */
typedef struct gdsp_ivl_table gdsp_ivl_table;

typedef struct gdsp_pw_op
	{
	int32_t  code;            /* gdsp_pw_code                                   */
	uint32_t flags;
	double   a, b, c;
	const gdsp_ivl_table* table;   /* for GDSP_PW_IVL_* only                    */
	} gdsp_pw_op;

#define GDSP_MAX_POINTWISE 16

/* Build a device interval table from host SoA intervals (segment index,
 * chromosome start/end, value).  Intervals must be sorted by (layout order,
 * start) and pairwise disjoint -- the host reader establishes that (it is what
 * multiply.c:308-309 enforces, and what the host-side union/fold produces for
 * add/mask/or). */
int  gdsp_ivl_table_create (gdsp_ctx* ctx, const gdsp_layout* lay,
                            const uint32_t* h_seg, const uint32_t* h_start,
                            const uint32_t* h_end, const double* h_val,
                            uint64_t n, gdsp_ivl_table** out);
void gdsp_ivl_table_destroy (gdsp_ivl_table* t);

int  gdsp_pointwise (gdsp_ctx* ctx, const gdsp_layout* lay, const double* in,
                     double* out, const gdsp_pw_op* ops, int nops);

/* min and max over all owned cells (invert's auto mid add.c:907-926;
 * percentile 0 / 100 percentile.c:434-530 with stride/min/max filter) */
int  gdsp_minmax (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                  uint32_t stride, double min_allowed, double max_allowed,
                  double* h_min, double* h_max, uint64_t* h_count);

/* how many owned cells are NOT integers with |v| <= limit (NaN, infinities count).  The host's
 * add / subtract <file> (add.c:280-281 adds interval after interval, cell by cell) may fold
 * overlapping integer-valued intervals into one difference array only when this is 0. */
int  gdsp_count_non_integer (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                             double limit, uint64_t* h_count);

/* percentile --preserve: what write_all_chromosomes + read_all_chromosomes
 * (genodsp.c:1717-1775) leave in the vectors: every value v != 0 becomes
 * strtod(printf("%.10f", v)) (inf -> DBL_MAX), every zero +0.0.  Exact integer
 * arithmetic on the device, no text.  decimals must be 10.  In place. */
int  gdsp_text_roundtrip (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig, int decimals);

/* minover / maxover, minmax.c:322-343 / :725-746: for every interval of the table find
 * the cell holding the minimum (maximum); ties go to the cell farthest from both ends of
 * its interval, then to the earliest.  The chosen buffer cell index replaces the table's
 * value column (device side); apply GDSP_PW_IVL_KEEP_AT with the same table next. */
int  gdsp_ivl_arg_extrema (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                           gdsp_ivl_table* table, int want_max);

/* map, map.c:263-357: piecewise-linear function through n breakpoints (h_in strictly
 * ascending); values at or beyond the ends take the end outputs.  In place. */
int  gdsp_map_values (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                      const double* h_in, const double* h_out, int n);

/* ---- percentile -------------------------------------------------------------
 * op_percentile_apply, percentile.c:392-751.  Order statistics of the samples
 * v[ix], ix = 0,stride,2*stride,.. of every chromosome with
 * min_allowed <= v <= max_allowed.  p_milli[] are percentiles in thousandths
 * of a percent; the rank of each is (u32)((u64)n*p/100000.0) as in
 * percentile.c:588,686 (p = 100000 -> the largest sample, :690-707).
 * Non-destructive and exact (selection, no sorting of the genome).
 * `tmp` is a scratch buffer of buffer_cells doubles (>= 1024). */
int  gdsp_percentiles (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                       double* tmp, uint64_t buffer_cells, uint32_t stride,
                       double min_allowed, double max_allowed,
                       const uint32_t* h_p_milli, int np, double* h_values,
                       uint64_t* h_num_samples);
/* The same selection, also reporting for every percentile how many samples have a key BELOW the
 * returned value and how many EQUAL it (exact; h_below/h_equal: np entries) and the number of NaN
 * samples.  With every cell qualifying these give `binarize` on the reference's sorted post-state
 * without a counting pass: the step of gdsp_fill_step is below (+ equal unless ties go above). */
int  gdsp_percentiles_ranked (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                              double* tmp, uint64_t buffer_cells, uint32_t stride,
                              double min_allowed, double max_allowed,
                              const uint32_t* h_p_milli, int np, double* h_values,
                              uint64_t* h_num_samples, uint64_t* h_below, uint64_t* h_equal, uint64_t* h_nan);
/* Building blocks of gdsp_percentiles for slab-sharded runs (one rank = one
 * slab): sample the qualifying cells whose order-preserving key lies in
 * [key_lo,key_hi] (unsorted, into d_out, capacity m); sort a plain device
 * array; count the cells in the regions delimited by ascending key bounds
 * (region 2k = keys below bound k and above bound k-1, region 2k+1 = equal to
 * bound k; 2*nb+1 regions) while compacting the cells of the open regions whose
 * h_compact[k] is set.  The host combines the per-rank results with
 * all-gather / all-reduce (genodsp_b200/slab.py).  The key of a double is its
 * bit pattern with the sign bit flipped for non-negatives and all bits flipped
 * for negatives. */
int  gdsp_pct_sample  (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig, uint32_t stride,
                       double min_allowed, double max_allowed, uint64_t key_lo, uint64_t key_hi,
                       uint32_t m, uint64_t seed, double* d_out, uint32_t* h_count, uint64_t* h_slots);
int  gdsp_sort_array  (gdsp_ctx* ctx, double* d_a, double* d_b, uint64_t n, int* h_result_in_b);
/* The collect pass of op_percentile_apply (percentile.c:547-580) for --window / --min / --max: the j-th
 * qualifying sample (every stride-th chromosome coordinate with min <= v <= max, chromsSorted order) is
 * swapped with position j of the concatenated genome.  Out of place (sig -> out, whole chromosomes
 * only); *h_n = the number of qualifying samples: positions [0, n) of `out` then hold them in scan order,
 * ready for the per-chromosome sorts and bubble passes (gdsp_sort_genome on the front part), the rest is
 * the reference's shuffle of the non-qualifying values.  `work` = gdsp_percentile_collect_work_bytes(). */
size_t gdsp_percentile_collect_work_bytes (uint64_t buffer_cells);
int  gdsp_percentile_collect (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig, double* out,
                              uint64_t buffer_cells, void* work, uint32_t stride,
                              double min_allowed, double max_allowed, uint64_t* h_n);
/* positions [*h_lo, *h_hi) of the cells equal to `value` (same key) in an array sorted by gdsp_sort_array */
int  gdsp_equal_range (gdsp_ctx* ctx, const double* d_sorted, uint64_t n, double value, uint64_t* h_lo, uint64_t* h_hi);
int  gdsp_pct_count   (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig, uint32_t stride,
                       double min_allowed, double max_allowed, const uint64_t* h_bound_keys, int nb,
                       const uint8_t* h_compact, uint64_t* h_counts, double* d_cand, uint64_t cap,
                       uint64_t* h_ncand);
/* gdsp_pct_count that also returns the number of qualifying NaN cells (they sit at the ends of the key order) */
int  gdsp_pct_count_nan (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig, uint32_t stride,
                         double min_allowed, double max_allowed, const uint64_t* h_bound_keys, int nb,
                         const uint8_t* h_compact, uint64_t* h_counts, double* d_cand, uint64_t cap,
                         uint64_t* h_ncand, uint64_t* h_nan);
/* The reference's percentile is destructive; with every cell qualifying and
 * the last requested rank in the last two chromosomes the genome ends up
 * globally sorted in layout order (percentile.c:611-651).
 * `tmp` is a second buffer of buffer_cells doubles; the sorted genome lands
 * in sig (*h_result_in_tmp = 0) or in tmp (= 1): the caller swaps its buffers. */
int  gdsp_sort_genome (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                       double* tmp, uint64_t buffer_cells, int* h_result_in_tmp);
/* One step of the reference's bubble passes, combine_sorted_vectors (percentile.c:820-864): the ranges
 * [c_lo, c_lo+c_len) and [d_lo, d_lo+d_len) of sig are each sorted ascending (gdsp_sort_genome's order);
 * afterwards the first holds the c_len smallest cells of both, the second the rest, both sorted -- the
 * bytes a joint sort of the two would leave.  Done as a split search and two merges through tmp (same
 * offsets), 16 B per cell.  *h_moved = how many cells changed sides (0: nothing was written). */
int  gdsp_merge_exchange (gdsp_ctx* ctx, double* sig, double* tmp, uint64_t c_lo, uint64_t c_len,
                          uint64_t d_lo, uint64_t d_len, uint64_t* h_moved);
/* op_binarize_apply (logical.c:216-268) applied to the post-percentile state above WITHOUT sorting:
 * the binarized sorted genome is a step function at cells - #(v > threshold) (>= with ties above),
 * so one counting pass and one fill produce the same bytes.  *h_done = 0 (signal untouched) when
 * the signal holds NaNs: sort (gdsp_sort_genome) and binarize (gdsp_pointwise) instead. */
int  gdsp_sorted_binarize (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig, double threshold,
                           int ties_above, double one, double zero, int* h_done);

/* The fill half of gdsp_sorted_binarize for a slab-sharded genome: cell at position q of the
 * concatenated chromsSorted genome becomes `one` if q >= step, else `zero`; h_prefix[s] (nseg host
 * entries) is the position of segment s's first owned cell.  The caller obtains `step` by summing the
 * per-rank region counts of gdsp_pct_count (NCCL all-reduce). */
int  gdsp_fill_step (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig, const uint64_t* h_prefix,
                     uint64_t step, double one, double zero);

/* ---- text output -----------------------------------------------------------
 * The fprintf loop of report_intervals (genodsp.c:1606-1678) on the device: run r
 * becomes the line  chrom TAB start[r]+add_start TAB end[r]+add_end [TAB "%.*f"
 * of val[r]] NEWLINE  in d_text (device memory, capacity cap bytes); *h_bytes is
 * the length of the text.  The value is printed exactly as glibc does (round
 * half even on the binary value, "-0.000" for negative values that round to
 * zero).  *h_unsupported = 1 (and nothing written) when some value is NaN,
 * infinite or >= 2^63, the precision is above 17 or the name longer than 255:
 * the caller then formats those runs itself.  GDSP_ERR_CAPACITY if cap is too
 * small (*h_bytes = needed).  gdsp_format_runs_max_bytes bounds the text of n runs. */
size_t gdsp_format_runs_max_bytes (uint64_t n, const char* chrom);
int  gdsp_format_runs (gdsp_ctx* ctx, const uint32_t* d_start, const uint32_t* d_end,
                       const double* d_val, uint64_t n, const char* chrom,
                       uint32_t add_start, uint32_t add_end, int with_value, int precision,
                       char* d_text, uint64_t cap, uint64_t* h_bytes, int* h_unsupported);

/* ---- clump ---------------------------------------------------------------- */
/* clump_search, clump.c:494-736 (above=1 clump, 0 anticlump); in place.
 * `work` = device buffer of gdsp_clump_work_bytes(buffer_cells). */
size_t gdsp_clump_work_bytes (uint64_t buffer_cells);
int  gdsp_clump (gdsp_ctx* ctx, const gdsp_layout* lay, double* sig,
                 uint64_t buffer_cells, void* work, double average,
                 uint32_t min_length, double relative_length, int above,
                 double one_val, double zero_val);

/* clump / anticlump on a SLAB-SHARDED genome (one GPU owns a contiguous piece of a chromosome): the
 * chromosome-wide dependencies of clump_search (prefix sums and minima going right, the suffix maximum
 * of valid ends going left, the run trimming both ways) travel as per-piece carries; the signal never
 * moves.  Four phases, each taking/returning a few host numbers per segment which the caller combines
 * across ranks (NCCL all-gather; genodsp_b200/slab.py:slab_clump_carries shows the folds):
 *   reduce  -> h_agg[5*s..]   = {head sum, head min prefix, tail sum, tail min prefix, all-negative flag}
 *              (tail = the piece's last 4096-cell tile when the chromosome continues to the right)
 *   mark    <- h_carry_in[2*s..] = {P, M} before the first cell of the segment's halo tile (M includes P[-1]=0)
 *           -> h_sufmax[s]    = maximum valid prefix sum over the owned cells (-inf if none)
 *   trim    <- h_sufmax_in[s] = maximum of h_sufmax over the pieces to the right (-inf if none)
 *           -> h_gp[s]        = generate/propagate pair of the owned tiles (bits 0-1 upwards, 2-3 downwards)
 *   emit    <- h_cin[s] bit0/1 = a trimmed run reaches the piece from the left/right; h_allneg[s] = the
 *              whole chromosome has no qualifying cell (clump.c:545-565); writes one/zero in place.
 * Requirements: segments with pos0 > 0 have >= 4096 valid halo cells on the left (dlo <= lo-4096), cuts are
 * multiples of 4096 in chromosome coordinates, minimum length <= 4096.  `work` as for gdsp_clump. */
typedef struct gdsp_clump_slab gdsp_clump_slab;
int  gdsp_clump_slab_create (gdsp_ctx* ctx, const gdsp_layout* lay, uint64_t buffer_cells, void* work,
                             double average, uint32_t min_length, double relative_length, int above,
                             double one_val, double zero_val, gdsp_clump_slab** out);
int  gdsp_clump_slab_reduce (gdsp_clump_slab* cs, const double* sig, double* h_agg);
int  gdsp_clump_slab_mark   (gdsp_clump_slab* cs, const double* sig, const double* h_carry_in, double* h_sufmax);
int  gdsp_clump_slab_trim   (gdsp_clump_slab* cs, const double* sig, const double* h_sufmax_in, int* h_gp);
int  gdsp_clump_slab_emit   (gdsp_clump_slab* cs, double* sig, const unsigned char* h_cin, const int* h_allneg);
void gdsp_clump_slab_destroy (gdsp_clump_slab* cs);

/* ---- multi-GPU plumbing: NCCL under the C boundary ---------------------------------
 * SURVEY 8(e) / BASELINE north_star: slab halos travel by NCCL send/recv, region counts and variables by NCCL
 * all-reduce / broadcast.  A communicator belongs to one context (one GPU, one stream); the exchanges are
 * enqueued on that stream.  One process per GPU: rank 0 calls gdsp_comm_unique_id, the 128 bytes reach the other
 * ranks through the launcher (torchrun's store, MPI, a file), every rank calls gdsp_comm_create.  One process for
 * all GPUs of a box (a multi-GPU C host): gdsp_comm_create_all (ncclCommInitAll) and the *_all exchange. */
typedef struct gdsp_comm gdsp_comm;
#define GDSP_COMM_ID_BYTES 128
typedef struct gdsp_halo
	{
	int32_t  peer;               /* rank holding the neighbouring slab                         */
	uint64_t send_lo, send_hi;   /* owned cells next to the cut, sent to the peer               */
	uint64_t recv_lo, recv_hi;   /* halo cells filled with the peer's cells                     */
	} gdsp_halo;
int  gdsp_comm_unique_id  (unsigned char* id128);
int  gdsp_comm_create     (gdsp_ctx* ctx, const unsigned char* id128, int nranks, int rank, gdsp_comm** out);
int  gdsp_comm_create_all (gdsp_ctx** ctxs, int n, gdsp_comm** out /* n entries */);
void gdsp_comm_destroy    (gdsp_comm* comm);
int  gdsp_comm_rank       (const gdsp_comm* comm);
int  gdsp_comm_size       (const gdsp_comm* comm);
int  gdsp_comm_exchange_halos     (gdsp_comm* comm, double* sig, const gdsp_halo* plan, int nplan);
int  gdsp_comm_exchange_halos_all (gdsp_comm** comms, double** sigs, const gdsp_halo* const* plans, const int* nplans, int n);
int  gdsp_comm_allreduce_sum_u64  (gdsp_comm* comm, uint64_t* h_values, int n);               /* host in/out  */
int  gdsp_comm_allgather_f64      (gdsp_comm* comm, const double* h_in, int n, double* h_out); /* host; size*n out */
int  gdsp_comm_allgather_dev      (gdsp_comm* comm, const double* d_in, uint64_t n, double* d_out); /* device, async */
int  gdsp_comm_broadcast_f64      (gdsp_comm* comm, double* h_values, int n, int root);        /* host in/out  */

/* ---- run-length output ------------------------------------------------------
 * report_intervals, genodsp.c:1561-1691: maximal runs of raw-equal values
 * (every cell its own run when collapse==0); runs of value 0 are dropped
 * unless show_uncovered==1.  Runs are produced in layout order; seg_first[s]
 * (nseg+1 entries, host) receives the index of segment s's first run.
 * Returns GDSP_ERR_CAPACITY (and the needed count in *n_runs) when cap is too
 * small.  starts/ends are chromosome coordinates (0-based, end exclusive). */
int  gdsp_runs (gdsp_ctx* ctx, const gdsp_layout* lay, const double* sig,
                int collapse, int show_uncovered,
                uint32_t* d_start, uint32_t* d_end, double* d_val, uint64_t cap,
                uint64_t* h_n_runs, uint64_t* h_seg_first);

#ifdef __cplusplus
}
#endif
#endif /* GDSP_B200_H */
