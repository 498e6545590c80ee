/* gd_device.h -- host-side view of the genome on the GPU (internal to the host
 * layer; operators use it to reach the kernels of include/gdsp_b200.h). */
#ifndef gd_device_H
#define gd_device_H

#include "genodsp_interface.h"
#include "gdsp_b200.h"

typedef struct gdev
	{
	gdsp_ctx*     ctx;
	gdsp_layout*  genome;        /* every chromosome, chromsSorted order             */
	gdsp_layout** single;        /* [nchrom] one-segment layouts, same buffer cells  */
	gdsp_seg*     segs;          /* [nchrom] segment table (sorted order)            */
	int           nchrom;
	u64           cells;         /* buffer size in cells                             */
	double*       sig;           /* current signal                                   */
	double*       tmp;           /* ping-pong partner / scratch                      */
	void*         work;          /* grow-only byte scratch                           */
	size_t        work_bytes;
	u32           maxLength;
	int           pendingSorted; /* the signal is notionally the globally sorted genome a percentile leaves
	                                behind (percentile.c:611-651) but has not been sorted: set only when the
	                                next operator is binarize, which does not need the sort
	                                (gdsp_sorted_binarize); cleared by gd_materialise_sorted            */
	/* what the selection pass of that percentile learned (every cell qualifying, no NaN): for each reported
	 * value the number of cells below it and equal to it -- binarize then needs a fill, not a count */
	int           knownN;
	double        knownVal[1024];
	u64           knownBelow[1024], knownEqual[1024];
	} gdev;

extern gdev gd;

void  gd_device_open   (void);           /* after sort_chromosomes_by_length: host bookkeeping, then the CUDA
                                            part on a helper thread                                     */
void  gd_device_wait   (void);           /* join that thread; call before the first device access     */
void  gd_device_close  (void);
void  gd_check         (int status, const char* who);   /* fatal on error               */
void  gd_materialise_sorted (const char* who);   /* gd_ops_percentile.c: sort now if pendingSorted */
void  gd_swap          (void);           /* sig <-> tmp, refresh every spec->valVector  */
void* gd_work          (size_t bytes);
int   gd_sorted_index  (spec* chromSpec);
/* layout to use for an apply call: whole genome when v == NULL, else the
 * chromosome whose vector is v */
const gdsp_layout* gd_layout_for (valtype* v, int* sortedIx);
/* out-of-place operators: after computing tmp from sig over `lay`, make the
 * result current (pointer swap for the genome, copy-back for one chromosome) */
void  gd_commit_tmp    (valtype* v, int sortedIx);

/* ---- interval files as structure-of-arrays ------------------------------- */

typedef struct ivlist
	{
	u64     n, cap;
	u32*    seg;                 /* chromsSorted index                               */
	u32*    start;               /* vector coordinates (origin shifted, spec.start
	                                removed, clipped), half open                      */
	u32*    end;
	double* val;
	} ivlist;

/* exact (file-order) application of valued intervals, gd_device.c; false = too deep, nothing done */
#define GD_EXACT_ADD   0
#define GD_EXACT_SUB   1
#define GD_EXACT_CLEAR 2
int  gd_apply_intervals_exact (ivlist* l, int mode, valtype missing, const char* opName);
int  gd_fold_intervals_minmax (ivlist* l, int wantMax, valtype missing, ivlist* out);

void ivlist_init    (ivlist* l);
void ivlist_push    (ivlist* l, u32 seg, u32 start, u32 end, double val);
void ivlist_free    (ivlist* l);
/* stable sort by (seg,start); merge overlapping/abutting intervals (values dropped) */
void ivlist_sort    (ivlist* l);
void ivlist_union   (ivlist* l);

/* Read an interval file the way the reference's file-driven operators do
 * (add.c:231-282 and friends): unknown chromosomes skipped, end beyond the
 * chromosome fatal (when spec.start==0) or clipped, optional val==0 filter.
 * requireSorted additionally enforces multiply.c's "sorted, non-overlapping,
 * chromosome-contiguous" rule with its error texts. */
void ivlist_read_file (ivlist* l, const char* opName, const char* filename,
                       int valCol, int originOne, int skipZeroVal, int requireSorted);

#endif
