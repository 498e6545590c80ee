/* gd_ops_io.c -- input, output, variables (reference opio.c, variables.c) and the
 * operators of minmax.c / map.c that are not part of this build's hot path. */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "gd_ops.h"

/* ==================================================================== input */

typedef struct dspop_input
	{
	dspop common;
	char* filename;
	int   valColumn;
	int   missingVal;              /* an int in the reference too (opio.c:35) */
	int   overlapOp, originOne, destroyFile;
	} dspop_input;

void op_input_short (char* name, int w, FILE* f, char* indent)
	{ op_short_line (name, w, f, indent, "read a new set of interval values from a file, replacing the current set"); }

void op_input_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace the signal by the intervals of a file.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s <filename> [options]\n", indent, name);
	fprintf (f, "%s  --value=<col>            column of the file that holds the value (default 4)\n", indent);
	fprintf (f, "%s  --novalue                the file has no value column (every value is 1)\n", indent);
	fprintf (f, "%s  --missing=<value>        value of positions no interval covers (default 0)\n", indent);
	fprintf (f, "%s  --overlap=sum            where intervals overlap, add them (default)\n", indent);
	fprintf (f, "%s  --overlap=min            ... keep the minimum\n", indent);
	fprintf (f, "%s  --overlap=max            ... keep the maximum\n", indent);
	fprintf (f, "%s  --origin=one             the file's intervals are origin-one, closed\n", indent);
	fprintf (f, "%s  --origin=zero            the file's intervals are origin-zero, half-open\n", indent);
	fprintf (f, "%s  --destroy                delete the file after reading it\n", indent);
	}

dspop* op_input_parse (char* name, int argc, char** argv)
	{
	dspop_input* op = (dspop_input*) op_alloc (name, sizeof (dspop_input));
	op->common.atRandom = true;
	op->valColumn = (int) get_named_global ("valColumn", 4-1);
	op->overlapOp = ri_overlapSum;
	op->originOne = (int) get_named_global ("originOne", false);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp (arg, "--novalue") == 0 || strcmp (arg, "--novalues") == 0 || strcmp (arg, "--value=none") == 0)
			op->valColumn = -1;
		else if (strcmp_prefix (arg, "--value=") == 0)
			{
			int col = string_to_int (argVal) - 1;
			if (col == -1) chastise ("[%s] value column can't be 0 (\"%s\")\n", name, arg);
			if (col < 0)   chastise ("[%s] value column can't be negative (\"%s\")\n", name, arg);
			if (col < 3)   chastise ("[%s] value column can't be 1, 2 or 3 (\"%s\")\n", name, arg);
			op->valColumn = col;
			}
		else if (strcmp_prefix (arg, "--missing=") == 0) op->missingVal = (int) string_to_valtype (argVal);
		else if (strcmp (arg, "--overlap=sum") == 0) op->overlapOp = ri_overlapSum;
		else if (strcmp (arg, "--overlap=minimum") == 0 || strcmp (arg, "--overlap=min") == 0) op->overlapOp = ri_overlapMin;
		else if (strcmp (arg, "--overlap=maximum") == 0 || strcmp (arg, "--overlap=max") == 0) op->overlapOp = ri_overlapMax;
		else if (strcmp (arg, "--origin=one") == 0 || strcmp (arg, "--origin=1") == 0)  op->originOne = true;
		else if (strcmp (arg, "--origin=zero") == 0 || strcmp (arg, "--origin=0") == 0) op->originOne = false;
		else if (strcmp (arg, "--destroy") == 0) op->destroyFile = true;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (op->filename == NULL) op->filename = copy_string (arg);
		else bad_arg (name, arg);
		}
	if (op->filename == NULL) { fprintf (stderr, "[%s] no filename was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}

void op_input_free (dspop* _op) { dspop_input* op = (dspop_input*) _op;  free (op->filename);  free (op); }

void op_input_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), arg_dont_complain(valtype* v))
	{
	dspop_input* op = (dspop_input*) _op;
	FILE* f = fopen (op->filename, "rt");
	if (f == NULL) { fprintf (stderr, "[%s] can't open \"%s\" for reading\n", _op->name, op->filename);  exit (EXIT_FAILURE); }
	read_intervals (f, op->valColumn, op->originOne, op->overlapOp, true, (valtype) op->missingVal);
	fclose (f);
	if (op->destroyFile) remove (op->filename);
	}

/* --overlap=min/max (genodsp.c:1307-1322): every covered position ends up with the smallest
 * (largest) value of the intervals covering it, uncovered positions with missingVal.  The host
 * reduces the overlaps to disjoint pieces (intervals visited from the best value on, each piece
 * painted once through a "next unpainted piece" union-find); the GPU fills and assigns. */
static ivlist* mmList;
static int     mmWantMax;
static int mm_by_value (const void* a, const void* b)
	{
	u64 i = *(const u64*) a, j = *(const u64*) b;
	double x = mmList->val[i], y = mmList->val[j];
	if (x != y) return mmWantMax ? ((x > y) ? -1 : 1) : ((x < y) ? -1 : 1);
	return (i < j) ? -1 : (i > j);
	}
static int u32_cmp (const void* a, const void* b)
	{ u32 x = *(const u32*) a, y = *(const u32*) b;  return (x > y) - (x < y); }

void gd_input_minmax (ivlist* l, int overlapOp, arg_dont_complain(int clear), valtype missingVal)
	{
	gd_check (gdsp_fill (gd.ctx, gd.genome, gd.sig, missingVal), "input");
	if (l->n == 0) return;
	ivlist out;  ivlist_init (&out);
	u64* order = (u64*) malloc (l->n * sizeof (u64));
	for (u64 k = 0; k < l->n; k++) order[k] = k;
	mmList = l;  mmWantMax = (overlapOp == ri_overlapMax);
	qsort (order, l->n, sizeof (u64), mm_by_value);

	for (int seg = 0; seg < gd.nchrom; seg++)
		{
		/* breakpoints of this chromosome */
		u64 cnt = 0;
		for (u64 k = 0; k < l->n; k++) if ((int) l->seg[k] == seg) cnt++;
		if (cnt == 0) continue;
		u32* bp = (u32*) malloc (2 * cnt * sizeof (u32));
		u64 nb = 0;
		for (u64 k = 0; k < l->n; k++) if ((int) l->seg[k] == seg) { bp[nb++] = l->start[k];  bp[nb++] = l->end[k]; }
		qsort (bp, nb, sizeof (u32), u32_cmp);
		u64 u = 0;
		for (u64 k = 0; k < nb; k++) if (u == 0 || bp[k] != bp[u-1]) bp[u++] = bp[k];
		nb = u;                                              /* pieces: [bp[p], bp[p+1]) for p < nb-1 */
		u64 npiece = nb - 1;
		u64* next = (u64*) malloc ((npiece + 1) * sizeof (u64));   /* next unpainted piece at or after p */
		double* pv = (double*) malloc (npiece * sizeof (double));
		char* painted = (char*) calloc (npiece, 1);
		for (u64 p = 0; p <= npiece; p++) next[p] = p;
		for (u64 q = 0; q < l->n; q++)
			{
			u64 k = order[q];
			if ((int) l->seg[k] != seg) continue;
			/* first piece of the interval: lower bound of start in bp */
			u64 lo = 0, hi = nb;
			while (lo < hi) { u64 mid = (lo + hi) / 2;  if (bp[mid] < l->start[k]) lo = mid + 1; else hi = mid; }
			u64 p = lo;
			while (true)
				{
				/* find (with path halving) the next unpainted piece */
				u64 r = p;
				while (next[r] != r) { next[r] = next[next[r]];  r = next[r]; }
				p = r;
				if (p >= npiece || bp[p] >= l->end[k]) break;
				painted[p] = 1;  pv[p] = l->val[k];
				next[p] = p + 1;
				p = p + 1;
				}
			}
		for (u64 p = 0; p < npiece; p++)
			if (painted[p]) ivlist_push (&out, (u32) seg, bp[p], bp[p+1], pv[p]);
		free (bp);  free (next);  free (pv);  free (painted);
		}
	free (order);
	gdsp_ivl_table* t;
	gd_check (gdsp_ivl_table_create (gd.ctx, gd.genome, out.seg, out.start, out.end, out.val, out.n, &t), "input");
	gdsp_pw_op p;  memset (&p, 0, sizeof (p));
	p.code = GDSP_PW_IVL_ASSIGN;  p.table = t;
	gd_check (gdsp_pointwise (gd.ctx, gd.genome, gd.sig, gd.sig, &p, 1), "input");
	gdsp_ivl_table_destroy (t);
	ivlist_free (&out);
	}

/* =================================================================== output */

typedef struct dspop_output
	{
	dspop common;
	char* filename;
	int   valPrecision, noOutputValues, collapseRuns, showUncovered, originOne;
	} dspop_output;

void op_output_short (char* name, int w, FILE* f, char* indent)
	{ op_short_line (name, w, f, indent, "write the current set of interval values to a file"); }

void op_output_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sWrite the signal, as it stands at this point of the pipeline, to a file.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s <filename> [options]\n", indent, name);
	fprintf (f, "%s  --precision=<number>     number of digits to round values to\n", indent);
	fprintf (f, "%s  --nooutputvalue          don't write values, only intervals\n", indent);
	fprintf (f, "%s  --nocollapse             don't collapse runs of identical values\n", indent);
	fprintf (f, "%s  --uncovered:hide         don't write intervals that have no coverage\n", indent);
	fprintf (f, "%s  --uncovered:show         write intervals that have no coverage\n", indent);
	fprintf (f, "%s  --uncovered:NA           mark uncovered intervals as NA\n", indent);
	fprintf (f, "%s  --origin=one             intervals are written origin-one, closed\n", indent);
	fprintf (f, "%s  --origin=zero            intervals are written origin-zero, half-open\n", indent);
	fprintf (f, "%s(defaults are the settings given before this operator on the command line)\n", indent);
	}

dspop* op_output_parse (char* name, int argc, char** argv)
	{
	dspop_output* op = (dspop_output*) op_alloc (name, sizeof (dspop_output));
	op->common.atRandom = true;
	op->valPrecision   = (int) get_named_global ("valPrecision",   0);
	op->noOutputValues = (int) get_named_global ("noOutputValues", false);
	op->collapseRuns   = (int) get_named_global ("collapseRuns",   true);
	op->showUncovered  = (int) get_named_global ("showUncovered",  uncovered_hide);
	op->originOne      = (int) get_named_global ("originOne",      false);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp_prefix (arg, "--precision=") == 0)
			{
			op->valPrecision = string_to_int (argVal);
			if (op->valPrecision < 0) chastise ("[%s] precision can't be negative (\"%s\")\n", name, arg);
			}
		else if (strcmp (arg, "--nooutputvalue") == 0 || strcmp (arg, "--nooutputvalues") == 0) op->noOutputValues = true;
		else if (strcmp (arg, "--nocollapse") == 0) op->collapseRuns = false;
		else if (strcmp (arg, "--uncovered:hide") == 0 || strcmp (arg, "--hide:uncovered") == 0) op->showUncovered = uncovered_hide;
		else if (strcmp (arg, "--uncovered:show") == 0 || strcmp (arg, "--show:uncovered") == 0) op->showUncovered = uncovered_show;
		else if (strcmp (arg, "--uncovered:NA") == 0 || strcmp (arg, "--uncovered:mark") == 0
		      || strcmp (arg, "--mark:uncovered") == 0 || strcmp (arg, "--markgaps") == 0) op->showUncovered = uncovered_NA;
		else if (strcmp (arg, "--origin=one") == 0 || strcmp (arg, "--origin=1") == 0)  op->originOne = true;
		else if (strcmp (arg, "--origin=zero") == 0 || strcmp (arg, "--origin=0") == 0) op->originOne = false;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (op->filename == NULL) op->filename = copy_string (arg);
		else bad_arg (name, arg);
		}
	if (op->filename == NULL) { fprintf (stderr, "[%s] no filename was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}

void op_output_free (dspop* _op) { dspop_output* op = (dspop_output*) _op;  free (op->filename);  free (op); }

void op_output_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), arg_dont_complain(valtype* v))
	{
	dspop_output* op = (dspop_output*) _op;
	FILE* f = fopen (op->filename, "wt");
	if (f == NULL) { fprintf (stderr, "[%s] can't open \"%s\" for writing\n", _op->name, op->filename);  exit (EXIT_FAILURE); }
	report_intervals (f, op->valPrecision, op->noOutputValues, op->collapseRuns, op->showUncovered, op->originOne);
	fclose (f);
	}

/* ================================================================ variables */

void op_show_variables_short (char* name, int w, FILE* f, char* indent)
	{ op_short_line (name, w, f, indent, "show the values of all named variables"); }
void op_show_variables_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sPrint every named variable and its value to stderr.\n%s\n%susage: %s\n", indent, indent, indent, name);
	}
dspop* op_show_variables_parse (char* name, int argc, char** argv)
	{
	dspop* op = (dspop*) op_alloc (name, sizeof (dspop));
	op->atRandom = true;
	if (argc > 0) bad_arg (name, argv[0]);
	return op;
	}
void op_show_variables_free (dspop* op) { free (op); }
void op_show_variables_apply (arg_dont_complain(dspop* op), arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), arg_dont_complain(valtype* v))
	{
	fprintf (stderr, "variables:\n");
	report_named_globals (stderr, "  ");
	}

/* ==== minover, maxover, minwith, maxwith, map: not on this build's hot path ==== */

typedef struct dspop_later { dspop common;  char* filename; } dspop_later;

static dspop* later_parse (char* name, int argc, char** argv)
	{
	dspop_later* op = (dspop_later*) op_alloc (name, sizeof (dspop_later));
	op->common.atRandom = true;
	for (; argc > 0; argv++, argc--)
		if (strcmp_prefix (argv[0], "--") != 0 && op->filename == NULL) op->filename = copy_string (argv[0]);
	if (op->filename == NULL) { fprintf (stderr, "[%s] no filename was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}
static void later_free (dspop* _op) { dspop_later* op = (dspop_later*) _op;  free (op->filename);  free (op); }
static void later_apply (dspop* op)
	{
	fprintf (stderr, "[%s] this operator has no GPU implementation in this build (and there is no CPU fallback)\n", op->name);
	exit (EXIT_FAILURE);
	}
static void later_usage (char* name, FILE* f, char* indent, const char* what)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%s%s\n%s(accepted on the command line; not implemented on the GPU in this build)\n%s\n", indent, what, indent, indent);
	fprintf (f, "%susage: %s <filename> [options]\n", indent, name);
	}

#define LATER_GROUP(fn, text, what) \
void   fn##_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, text); } \
void   fn##_usage (char* name, FILE* f, char* indent) { later_usage (name, f, indent, what); } \
dspop* fn##_parse (char* name, int argc, char** argv) { return later_parse (name, argc, argv); } \
void   fn##_free  (dspop* op) { later_free (op); } \
void   fn##_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), arg_dont_complain(valtype* v)) { later_apply (op); }

LATER_GROUP (op_max_in_interval, "find the maximum value in each of a set of intervals read from a file", "Keep, in each interval of a file, only the position holding the maximum.")
LATER_GROUP (op_min_in_interval, "find the minimum value in each of a set of intervals read from a file", "Keep, in each interval of a file, only the position holding the minimum.")
LATER_GROUP (op_min_with, "take the minimum of the current set of interval values and values read from a file", "Position-wise minimum of the signal and the intervals of a file.")
LATER_GROUP (op_max_with, "take the maximum of the current set of interval values and values read from a file", "Position-wise maximum of the signal and the intervals of a file.")
LATER_GROUP (op_map, "map the current set of interval values to new values", "Map values through a piecewise-linear table read from a file.")
