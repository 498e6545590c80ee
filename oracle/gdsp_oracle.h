/* gdsp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the per-base algorithms of rsharris/genodsp 0.0.10
 * (file:line citations are into /root/reference).  It is the checker for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Nothing under genodsp_b200/ links,
 * imports or calls it, and the product fails loudly without its CUDA library.
 *
 * Parity pinning: the reference ships no tests or golden vectors.
 * This restatement is pinned (a) against the reference itself, compiled
 * unmodified into oracle/_ref/ and driven in-process by tests/test_oracle_vs_ref.py,
 * and (b) against tests/golden/ fixtures generated from that binary by
 * tests/golden/make_golden.py.
 */
#ifndef GDSP_ORACLE_H
#define GDSP_ORACLE_H
#include <stdint.h>

/* interval accumulation, read_intervals genodsp.c:1307-1330 */
void gdo_accumulate (double* v, uint32_t n, const uint32_t* s, const uint32_t* e,
                     const double* val, uint64_t m, int overlapOp, int clear,
                     double missing);
/* sum.c */
void gdo_block_sum   (double* v, uint32_t n, uint32_t W, double denom, int actualDenom, double zero);
void gdo_sliding_sum (double* v, uint32_t n, uint32_t W, double denom);
void gdo_hann_taps   (double* w, uint32_t W);
void gdo_smooth      (double* v, uint32_t n, uint32_t W);
void gdo_cumulative  (double* v, uint32_t n);
/* minover / maxover on one chromosome (minmax.c:322-348, :725-751): s/e sorted, disjoint */
void gdo_over_intervals (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, uint32_t m, int want_max, double fill);
/* minwith / maxwith (minmax.c:1979-1982, :2265-2268): any order, overlaps allowed */
void gdo_with_intervals (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, const double* val, uint32_t m, int want_max);
/* map (map.c:263-357), breakpoints sorted ascending by input */
void gdo_map (double* v, uint32_t n, const double* in, const double* out, uint32_t len);
/* percentile --preserve: write_all_chromosomes + read_all_chromosomes (genodsp.c:1717-1775) */
void gdo_text_roundtrip10 (double* v, uint32_t n);
/* minmax.c */
void gdo_local_extrema (double* v, uint32_t n, uint32_t N, int wantMax, double fill);
void gdo_best_extrema  (double* v, uint32_t n, uint32_t W, int wantMax);
/* morphology.c */
void gdo_close  (double* v, uint32_t n, double L, double T, double one, double zero);
void gdo_open   (double* v, uint32_t n, double L, double T, double one, double zero);
void gdo_dilate (double* v, uint32_t n, uint32_t left, uint32_t right, double T, double one, double zero);
void gdo_erode  (double* v, uint32_t n, uint32_t left, uint32_t right, double T, double one, double zero);
/* pointwise: logical.c, add.c, mask.c */
void gdo_binarize (double* v, uint32_t n, double T, int tiesAbove, double one, double zero);
void gdo_addconst (double* v, uint32_t n, double c);
void gdo_abs      (double* v, uint32_t n);
void gdo_clip     (double* v, uint32_t n, int haveMin, double mn, int haveMax, double mx);
void gdo_erase    (double* v, uint32_t n, int haveMin, double mn, int haveMax, double mx, int keepInside, double zero);
void gdo_invert   (double* v, uint32_t n, double mid);
void gdo_minmax   (const double* v, uint32_t n, double* mn, double* mx);
void gdo_logical_prep (double* v, uint32_t n);
/* interval-driven ops on one chromosome (intervals already origin-shifted) */
void gdo_add_intervals      (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, const double* val, uint64_t m, double sign);
void gdo_mask_intervals     (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, uint64_t m, double maskVal);
void gdo_or_intervals       (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, const double* val, uint64_t m);
/* sorted, non-overlapping interval ops; kind: 0 multiply 1 divide 2 masknot 3 and */
void gdo_sorted_intervals   (double* v, uint32_t n, const uint32_t* s, const uint32_t* e, const double* val, uint64_t m, int kind, double aux);
/* clump.c */
void gdo_clump (double* v, uint32_t n, double T, uint32_t minLength, int above, double one, double zero);
/* percentile.c: samples of one chromosome appended to out; returns count */
uint64_t gdo_percentile_collect (const double* v, uint32_t n, uint32_t W, double mn, double mx, double* out);
uint64_t gdo_percentile_rank (uint64_t numValues, uint32_t pMilli);
void gdo_sort (double* v, uint64_t n);
/* report_intervals genodsp.c:1561-1691 for one chromosome: emits runs;
 * returns number of runs written (or needed if cap too small) */
uint64_t gdo_runs (const double* v, uint32_t n, int collapse, int showUncovered,
                   uint32_t* rs, uint32_t* re, double* rv, uint64_t cap);
#endif
