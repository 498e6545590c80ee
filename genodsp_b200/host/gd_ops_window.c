/* gd_ops_window.c -- windowed operators: sum, slidingsum, smooth, cumulativesum
 * (reference sum.c) and localmin, localmax, bestmin, bestmax (reference minmax.c).
 * Argument grammar, defaults and warnings follow the reference parsers; the
 * per-base work is one kernel launch over the packed genome. */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "gd_ops.h"

/* ======================================================================= sum */

typedef struct dspop_sum
	{
	dspop   common;
	u32     windowSize;
	valtype denominator;
	valtype zeroVal;
	int     windowIsChromosome, useActualDenom, denomIsWindowSize;
	int     exactOrder;          /* --exact-order (not in the reference): the running sum in the reference's sequential order */
	} dspop_sum;

static int arg_is_denom (char* arg)
	{
	return strcmp_prefix (arg, "--denom=") == 0 || strcmp_prefix (arg, "--denominator=") == 0
	    || strcmp_prefix (arg, "D=") == 0 || strcmp_prefix (arg, "--D=") == 0;
	}
static int arg_is_zero (char* arg)
	{ return strcmp_prefix (arg, "--zero=") == 0 || strcmp_prefix (arg, "Z=") == 0 || strcmp_prefix (arg, "--Z=") == 0; }

void op_window_sum_short (char* name, int nameWidth, FILE* f, char* indent)
	{ op_short_line (name, nameWidth, f, indent, "sum over non-overlapping windows"); }

void op_window_sum_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace the signal by its sum over consecutive, non-overlapping windows\n", indent);
	fprintf (f, "%scounted from the start of each chromosome. The sum lands in the window's\n", indent);
	fprintf (f, "%sfirst position; the other positions of the window become the zero value.\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --window=chromosome      one window spanning the whole chromosome\n", indent);
	fprintf (f, "%s  --window=<length>        (W=) size of window\n", indent);
	fprintf (f, "%s  --denom=<value>          (D=) divide each sum by this; \"window\" or \"W\" means\n", indent);
	fprintf (f, "%s                           the window size, \"actual\" the number of positions\n", indent);
	fprintf (f, "%s                           really in the (possibly truncated) window\n", indent);
	fprintf (f, "%s                           (default: no denominator)\n", indent);
	fprintf (f, "%s  --zero=<value>           (Z=) value written to the rest of each window\n", indent);
	fprintf (f, "%s                           (default is 0.0)\n", indent);
	}

dspop* op_window_sum_parse (char* name, int argc, char** argv)
	{
	dspop_sum* op = (dspop_sum*) op_alloc (name, sizeof (dspop_sum));
	op->common.atRandom = false;
	op->windowSize  = (u32) get_named_global ("windowSize", 100);
	op->denominator = 1.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp (arg, "--window=chromosome") == 0) op->windowIsChromosome = true;
		else if (arg_is_window (arg))
			{ op->windowSize = (u32) parse_window_arg (name, arg, argVal, 3, false, "window size");  op->windowIsChromosome = false; }
		else if (arg_is_denom (arg))
			{
			op->denominator = 1.0;  op->useActualDenom = false;  op->denomIsWindowSize = false;
			if (strcmp (argVal, "actual") == 0) op->useActualDenom = true;
			else if (strcmp (argVal, "window") == 0 || strcmp (argVal, "W") == 0) op->denomIsWindowSize = true;
			else
				{
				op->denominator = string_to_valtype (argVal);
				if (op->denominator == 0) chastise ("[%s] denominator can't be zero (\"%s\")\n", name, arg);
				}
			}
		else if (arg_is_zero (arg)) op->zeroVal = string_to_valtype (argVal);
		else bad_arg (name, arg);
		}
	if (!op->windowIsChromosome && op->windowSize < 3)
		{
		fprintf (stderr, "[%s] WARNING: raising window size from %d to %d\n", name, op->windowSize, 3);
		op->windowSize = 3;
		}
	return (dspop*) op;
	}

void op_window_sum_free (dspop* op) { free (op); }

void op_window_sum_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_sum* op = (dspop_sum*) _op;
	valtype denom = op->denomIsWindowSize ? (valtype) op->windowSize : op->denominator;
	/* with --window=chromosome and --denom=window the reference divides by the chromosome length,
	 * which is the "actual" block length */
	int actual = op->useActualDenom || (op->windowIsChromosome && op->denomIsWindowSize);
	gd_check (gdsp_block_sum (gd.ctx, gd_layout_for (v, NULL), gd.sig, op->windowSize, op->windowIsChromosome,
	                          denom, actual, op->zeroVal), _op->name);
	}

/* ================================================================ slidingsum */

void op_sliding_sum_short (char* name, int nameWidth, FILE* f, char* indent)
	{ op_short_line (name, nameWidth, f, indent, "continuous sum over overlapping windows"); }

void op_sliding_sum_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace every position by the sum of the signal over the window centred on\n", indent);
	fprintf (f, "%sit. Positions beyond the ends of a chromosome count as zero.\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --window=<length>        (W=) size of window\n", indent);
	fprintf (f, "%s  --denom=<value>          (D=) divide each sum by this (default: none)\n", indent);
	fprintf (f, "%s  --exact-order            (this implementation only) add and subtract in the\n", indent);
	fprintf (f, "%s                           sequential order of the original program: bit-identical\n", indent);
	fprintf (f, "%s                           on any real-valued signal, much slower\n", indent);
	}

dspop* op_sliding_sum_parse (char* name, int argc, char** argv)
	{
	dspop_sum* op = (dspop_sum*) op_alloc (name, sizeof (dspop_sum));
	op->common.atRandom = false;
	op->windowSize  = (u32) get_named_global ("windowSize", 100);
	op->denominator = 1.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (arg_is_window (arg)) op->windowSize = (u32) parse_window_arg (name, arg, argVal, 3, false, "window size");
		else if (arg_is_denom (arg))
			{
			/* "window" takes the window size known at this point of the argument list (sum.c:360) */
			if (strcmp (argVal, "window") == 0 || strcmp (argVal, "W") == 0) op->denominator = op->windowSize;
			else
				{
				op->denominator = string_to_valtype (argVal);
				if (op->denominator == 0) chastise ("[%s] denominator can't be zero (\"%s\")\n", name, arg);
				}
			}
		else if (strcmp (arg, "--exact-order") == 0) op->exactOrder = true;
		else bad_arg (name, arg);
		}
	if (op->windowSize < 3)
		{
		fprintf (stderr, "[%s] WARNING: raising window size from %d to %d\n", name, op->windowSize, 3);
		op->windowSize = 3;
		}
	return (dspop*) op;
	}

void op_sliding_sum_free (dspop* op) { free (op); }

void op_sliding_sum_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_sum* op = (dspop_sum*) _op;
	int ix;
	const gdsp_layout* lay = gd_layout_for (v, &ix);
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 1), _op->name);
	gd_check (gdsp_sliding_sum (gd.ctx, lay, gd.sig, gd.tmp, op->windowSize, op->denominator), _op->name);
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 0), _op->name);
	gd_commit_tmp (v, ix);
	}

/* ==================================================================== smooth */

#define maxSmoothWindow ((50*1000)+1)

typedef struct dspop_smooth { dspop common;  u32 windowSize; } dspop_smooth;

void op_smooth_short (char* name, int nameWidth, FILE* f, char* indent)
	{ op_short_line (name, nameWidth, f, indent, "apply a smoothing filter (Hann window)"); }

void op_smooth_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sSmooth the signal with a Hann window: every position becomes the weighted\n", indent);
	fprintf (f, "%ssum over the window centred on it. Positions beyond the ends of a\n", indent);
	fprintf (f, "%schromosome count as zero.\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --window=<length>        (W=) size of window\n", indent);
	fprintf (f, "%s                           (an even size is increased by 1)\n", indent);
	}

dspop* op_smooth_parse (char* name, int argc, char** argv)
	{
	dspop_smooth* op = (dspop_smooth*) op_alloc (name, sizeof (dspop_smooth));
	op->common.atRandom = false;
	op->windowSize = (u32) get_named_global ("windowSize", 101);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (arg_is_window (arg))
			{
			int w = string_to_unitized_int (argVal, true);
			if (w > maxSmoothWindow) chastise ("[%s] window size exceeds %u (\"%s\")\n", name, maxSmoothWindow, arg);
			op->windowSize = (u32) parse_window_arg (name, arg, argVal, 3, true, "window size");
			}
		else bad_arg (name, arg);
		}
	if ((op->windowSize & 1) == 0)
		{
		fprintf (stderr, "[%s] WARNING: raising window size from %d to %d\n", name, op->windowSize, op->windowSize + 1);
		op->windowSize++;
		}
	return (dspop*) op;
	}

void op_smooth_free (dspop* op) { free (op); }

void op_smooth_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_smooth* op = (dspop_smooth*) _op;
	u32 W = op->windowSize, h = (W - 1) / 2;
	/* Hann taps on the host with the same libm cos and the same arithmetic order as sum.c:634-645 */
	valtype* w = (valtype*) malloc ((size_t) W * sizeof (valtype));
	for (u32 k = 0; k <= h; k++)
		{
		double x = (k + 1) / (double) (W + 1);
		w[k] = w[W-1-k] = (1 - cos (2 * M_PI * x)) / 2;
		}
	valtype sum = 0.0;
	for (u32 k = 0; k < W; k++) sum += w[k];
	for (u32 k = 0; k < W; k++) w[k] /= sum;
	int ix;
	const gdsp_layout* lay = gd_layout_for (v, &ix);
	gd_check (gdsp_smooth (gd.ctx, lay, gd.sig, gd.tmp, W, w), _op->name);
	free (w);
	gd_commit_tmp (v, ix);
	}

/* ============================================================= cumulativesum */

void op_cumulative_sum_short (char* name, int nameWidth, FILE* f, char* indent)
	{ op_short_line (name, nameWidth, f, indent, "compute the cumulative sum of the current set of interval values"); }

void op_cumulative_sum_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace the signal by its running (cumulative) sum, restarting at the\n", indent);
	fprintf (f, "%sbeginning of every chromosome.\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s\n", indent, name);
	}

typedef struct dspop_cumsum { dspop common;  int exactOrder; } dspop_cumsum;

dspop* op_cumulative_sum_parse (char* name, int argc, char** argv)
	{
	dspop_cumsum* op = (dspop_cumsum*) op_alloc (name, sizeof (dspop_cumsum));
	op->common.atRandom = false;
	for (; argc > 0; argv++, argc--)
		{
		if (strcmp (argv[0], "--exact-order") == 0) op->exactOrder = true;   /* not in the reference: sequential summation order */
		else bad_arg (name, argv[0]);
		}
	return (dspop*) op;
	}

void op_cumulative_sum_free (dspop* op) { free (op); }

void op_cumulative_sum_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_cumsum* op = (dspop_cumsum*) _op;
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 1), _op->name);
	gd_check (gdsp_cumulative_sum (gd.ctx, gd_layout_for (v, NULL), gd.sig, gd.sig), _op->name);
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 0), _op->name);
	}

/* ========================================================= localmin / localmax */

typedef struct dspop_local { dspop common;  u32 neighborhood;  valtype fill; } dspop_local;

static dspop* local_parse (char* name, int argc, char** argv, int wantMax)
	{
	dspop_local* op = (dspop_local*) op_alloc (name, sizeof (dspop_local));
	op->common.atRandom = false;
	op->neighborhood = 3;                               /* not inherited from --window */
	op->fill = wantMax ? 0.0 : valtypeMax;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp_prefix (arg, "--neighborhood=") == 0 || strcmp_prefix (arg, "N=") == 0 || strcmp_prefix (arg, "--N=") == 0)
			op->neighborhood = (u32) parse_window_arg (name, arg, argVal, 3, true, "neighborhood");
		else if (wantMax && arg_is_zero (arg)) op->fill = string_to_valtype (argVal);
		else if (!wantMax && strcmp_prefix (arg, "--infinity=") == 0) op->fill = string_to_valtype (argVal);
		else bad_arg (name, arg);
		}
	return (dspop*) op;
	}

static void local_usage (char* name, FILE* f, char* indent, int wantMax)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sKeep the local %s of the signal: a position keeps its value when no other\n", indent, wantMax ? "maxima" : "minima");
	fprintf (f, "%sposition of its neighborhood is %s; every other position is\n", indent, wantMax ? "larger" : "smaller");
	fprintf (f, "%sreplaced by %s.\n", indent, wantMax ? "zero" : "infinity");
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --neighborhood=<length>  (N=) size of neighborhood (default is 3)\n", indent);
	if (wantMax) fprintf (f, "%s  --zero=<value>           (Z=) value for positions that are not maxima\n", indent);
	else         fprintf (f, "%s  --infinity=<value>       value for positions that are not minima\n", indent);
	}

static void local_apply (dspop* _op, valtype* v, int wantMax)
	{
	dspop_local* op = (dspop_local*) _op;
	int ix;
	const gdsp_layout* lay = gd_layout_for (v, &ix);
	gd_check (gdsp_local_extrema (gd.ctx, lay, gd.sig, gd.tmp, op->neighborhood, wantMax, op->fill), _op->name);
	gd_commit_tmp (v, ix);
	}

void   op_local_minima_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find local minima"); }
void   op_local_minima_usage (char* name, FILE* f, char* indent) { local_usage (name, f, indent, false); }
dspop* op_local_minima_parse (char* name, int argc, char** argv) { return local_parse (name, argc, argv, false); }
void   op_local_minima_free  (dspop* op) { free (op); }
void   op_local_minima_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { local_apply (op, v, false); }

void   op_local_maxima_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find local maxima"); }
void   op_local_maxima_usage (char* name, FILE* f, char* indent) { local_usage (name, f, indent, true); }
dspop* op_local_maxima_parse (char* name, int argc, char** argv) { return local_parse (name, argc, argv, true); }
void   op_local_maxima_free  (dspop* op) { free (op); }
void   op_local_maxima_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { local_apply (op, v, true); }

/* =========================================================== bestmin / bestmax */

typedef struct dspop_best { dspop common;  u32 windowSize;  int debug; } dspop_best;

static dspop* best_parse (char* name, int argc, char** argv)
	{
	dspop_best* op = (dspop_best*) op_alloc (name, sizeof (dspop_best));
	op->common.atRandom = false;
	op->windowSize = (u32) get_named_global ("windowSize", 100);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (arg_is_window (arg)) op->windowSize = (u32) parse_window_arg (name, arg, argVal, 3, false, "window size");
		else if (strcmp (arg, "--debug") == 0) op->debug = true;
		else bad_arg (name, arg);
		}
	if (op->windowSize < 3)
		{
		fprintf (stderr, "[%s] WARNING: raising window size from %d to %d\n", name, op->windowSize, 3);
		op->windowSize = 3;
		}
	return (dspop*) op;
	}

static void best_usage (char* name, FILE* f, char* indent, int wantMax)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace every position by the %s of the signal over the window centred\n", indent, wantMax ? "maximum" : "minimum");
	fprintf (f, "%son it (the window is clipped at the ends of a chromosome).\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --window=<length>        (W=) size of window\n", indent);
	}

static void best_apply (dspop* _op, valtype* v, int wantMax)
	{
	dspop_best* op = (dspop_best*) _op;
	int ix;
	const gdsp_layout* lay = gd_layout_for (v, &ix);
	gd_check (gdsp_best_extrema (gd.ctx, lay, gd.sig, gd.tmp, op->windowSize, wantMax), _op->name);
	gd_commit_tmp (v, ix);
	}

void   op_best_local_min_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find the minimum value in the window around each position"); }
void   op_best_local_min_usage (char* name, FILE* f, char* indent) { best_usage (name, f, indent, false); }
dspop* op_best_local_min_parse (char* name, int argc, char** argv) { return best_parse (name, argc, argv); }
void   op_best_local_min_free  (dspop* op) { free (op); }
void   op_best_local_min_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { best_apply (op, v, false); }

void   op_best_local_max_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find the maximum value in the window around each position"); }
void   op_best_local_max_usage (char* name, FILE* f, char* indent) { best_usage (name, f, indent, true); }
dspop* op_best_local_max_parse (char* name, int argc, char** argv) { return best_parse (name, argc, argv); }
void   op_best_local_max_free  (dspop* op) { free (op); }
void   op_best_local_max_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { best_apply (op, v, true); }
