/* gd_ops.h -- operator registry of the host layer (the reference's dspTable,
 * genodsp.c:117-174) plus the two properties the executor uses to schedule
 * built-in operators on the GPU: "pointwise" (fusable into one launch) and
 * "genome capable" (one launch covers every chromosome). */
#ifndef gd_ops_H
#define gd_ops_H

#include "genodsp_interface.h"
#include "gd_device.h"

dspprototypes(op_window_sum)
dspprototypes(op_sliding_sum)
dspprototypes(op_smooth)
dspprototypes(op_cumulative_sum)
dspprototypes(op_clump)
dspprototypes(op_skimp)
dspprototypes(op_percentile)
dspprototypes(op_add)
dspprototypes(op_subtract)
dspprototypes(op_add_constant)
dspprototypes(op_invert)
dspprototypes(op_multiply)
dspprototypes(op_divide)
dspprototypes(op_absolute_value)
dspprototypes(op_mask)
dspprototypes(op_mask_not)
dspprototypes(op_clip)
dspprototypes(op_erase)
dspprototypes(op_binarize)
dspprototypes(op_or)
dspprototypes(op_and)
dspprototypes(op_max_in_interval)
dspprototypes(op_min_in_interval)
dspprototypes(op_local_minima)
dspprototypes(op_local_maxima)
dspprototypes(op_best_local_min)
dspprototypes(op_best_local_max)
dspprototypes(op_min_with)
dspprototypes(op_max_with)
dspprototypes(op_close)
dspprototypes(op_open)
dspprototypes(op_dilate)
dspprototypes(op_erode)
dspprototypes(op_map)
dspprototypes(op_input)
dspprototypes(op_output)
dspprototypes(op_show_variables)

extern dspinfo dspTable[];
extern const u32 dspTableLen;

/* resources a pointwise descriptor holds until its launch has been issued */
typedef struct gd_pw_resources { gdsp_ivl_table* table; } gd_pw_resources;

/* every built-in operator handles v==NULL ("whole genome") in its apply */
int  gd_is_genome_capable    (dspop* op);
int  gd_is_pointwise         (dspop* op);
/* fills out[0] (resolving named variables / reading interval files in pipeline
 * order); returns 1, or 0 when the operator turns out to be a no-op */
int  gd_pointwise_descriptor (dspop* op, gdsp_pw_op* out, gd_pw_resources* res);
void gd_pw_release           (gd_pw_resources* res);
/* resolve (and announce on stderr) the named variables an operator refers to, now */
void gd_resolve_variables    (dspop* op);
void gd_resolve_pointwise    (dspop* op);      /* gd_ops_pointwise.c */
void gd_resolve_morph        (dspop* op);      /* gd_ops_morph.c     */

/* helpers shared between files */
void gd_set_outside   (ivlist* unionList, valtype value);                /* v = value outside the intervals */
void gd_input_minmax  (ivlist* l, int overlapOp, int clear, valtype missingVal);
void gd_paint_extreme (ivlist* l, int wantMax, ivlist* out);
void gd_run_pointwise_now (dspop* op);                                   /* apply path of a pointwise operator */
void op_short_line (char* name, int nameWidth, FILE* f, char* indent, const char* text);
dspop* gd_pipeline_head (void);
int    gd_output_inhibited (void);

/* common parse helpers */
int   arg_is_window   (char* arg);             /* --window= | W= | --W= */
int   parse_window_arg (char* name, char* arg, char* argVal, int minimum, int forceOdd, const char* what);
void* op_alloc        (char* name, size_t bytes);
void  bad_arg         (char* name, char* arg);

#endif
