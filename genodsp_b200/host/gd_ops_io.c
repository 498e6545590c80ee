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

/* reduce overlapping intervals to disjoint pieces that hold, for every covered position, the smallest
 * (largest) value among the intervals covering it; pieces come out in layout order */
void gd_paint_extreme (ivlist* l, int wantMax, ivlist* outp)
	{
	if (l->n == 0) return;
	u64* order = (u64*) malloc (l->n * sizeof (u64));
	for (u64 k = 0; k < l->n; k++) order[k] = k;
	mmList = l;  mmWantMax = wantMax;
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
			if (painted[p]) ivlist_push (outp, (u32) seg, bp[p], bp[p+1], pv[p]);
		free (bp);  free (next);  free (pv);  free (painted);
		}
	free (order);
	}

void gd_input_minmax (ivlist* l, int overlapOp, arg_dont_complain(int clear), valtype missingVal)
	{
	gd_check (gdsp_fill (gd.ctx, gd.genome, gd.sig, missingVal), "input");
	if (l->n == 0) return;
	ivlist out;  ivlist_init (&out);
	/* the plain extreme is the reference's result unless a value equals missingVal (the running value
	 * then reads as "not yet covered" and the next interval overwrites it) or is NaN: those inputs go
	 * through the file-order fold of every piece (gd_device.c) */
	int needFold = false;
	for (u64 k = 0; k < l->n && !needFold; k++) needFold = (l->val[k] == missingVal) || (l->val[k] != l->val[k]);
	if (!needFold || !gd_fold_intervals_minmax (l, overlapOp == ri_overlapMax, missingVal, &out))
		{
		if (needFold) fprintf (stderr, "[input] WARNING: intervals too deep for the file-order fold; a value equal to --missing counts as a value\n");
		gd_paint_extreme (l, overlapOp == ri_overlapMax, &out);
		}
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


/* ======================================================================== map
 * op_map_* (map.c:44-385): a piecewise-linear function read from a two-column text file (blank and
 * '#' lines skipped, any order, sorted by input value with qsort as map.c:461).  The reference applies
 * it chromosome by chromosome and re-reads the file for every chromosome; so does this operator when
 * it is handed one chromosome, and once when it is handed the whole genome.  Breakpoints that share
 * an input value make the reference's result depend on its piece cache: refused. */

typedef struct dspop_map { dspop common;  char* filename;  int destroyFile, debug, applied; } dspop_map;

void op_map_short (char* name, int w, FILE* f, char* indent)
	{ op_short_line (name, w, f, indent, "map values according to a piecewise-linear function"); }

void op_map_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sMap every value through a piecewise-linear function read from a file: two\n", indent);
	fprintf (f, "%snumbers per line, an input value and the output it maps to. Values between two\n", indent);
	fprintf (f, "%slisted inputs are interpolated linearly; values beyond the smallest or largest\n", indent);
	fprintf (f, "%slisted input take that end's output.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s <filename> [options]\n", indent, name);
	fprintf (f, "%s  --destroy                delete the file after reading it\n", indent);
	}

dspop* op_map_parse (char* name, int argc, char** argv)
	{
	dspop_map* op = (dspop_map*) op_alloc (name, sizeof (dspop_map));
	op->common.atRandom = false;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		if (strcmp (arg, "--destroy") == 0) op->destroyFile = true;
		else if (strcmp (arg, "--debug") == 0) op->debug = true;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (op->filename == NULL) op->filename = copy_string (arg);
		else bad_arg (name, arg);
		}
	if (op->filename == NULL) { fprintf (stderr, "[%s] no filename was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}

void op_map_free (dspop* _op) { dspop_map* op = (dspop_map*) _op;  free (op->filename);  free (op); }

typedef struct mapel { double vIn, vOut; } mapel;
static int mapel_ascending (const void* a, const void* b)
	{ const mapel* x = (const mapel*) a;  const mapel* y = (const mapel*) b;  return (x->vIn > y->vIn) - (x->vIn < y->vIn); }

static u32 mapLineNumber = 0;          /* the reference's counter is a function static: it keeps counting across files */

static int read_value_pair (FILE* f, double* v1, double* v2)
	{
	char line[1000];
	static int missingEol = false;
	while (fgets (line, sizeof (line), f) != NULL)
		{
		mapLineNumber++;
		if (missingEol)
			{ fprintf (stderr, "problem at line %u, line is longer than internal buffer\n", mapLineNumber - 1);  exit (EXIT_FAILURE); }
		size_t len = strlen (line);
		if (len != 0) missingEol = (line[len-1] != '\n');
		char* scan = skip_whitespace (line);
		if (*scan == 0 || *scan == '#') continue;
		if (line[0] == ' ')
			{ fprintf (stderr, "problem at line %u, line contains no first value\n", mapLineNumber);  exit (EXIT_FAILURE); }
		char* field = line;
		char* mark = skip_darkspace (field);
		scan = skip_whitespace (mark);
		if (*mark != 0) *mark = 0;
		*v1 = string_to_valtype (field);
		if (*scan == 0)
			{ fprintf (stderr, "problem at line %u, line contains no second value\n", mapLineNumber);  exit (EXIT_FAILURE); }
		field = scan;
		mark = skip_darkspace (field);
		if (*mark != 0) *mark = 0;
		*v2 = string_to_valtype (field);
		return true;
		}
	return false;
	}

void op_map_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_map* op = (dspop_map*) _op;
	FILE* f = fopen (op->filename, "rt");
	if (f == NULL) { fprintf (stderr, "[%s] can't open \"%s\" for reading\n", _op->name, op->filename);  exit (EXIT_FAILURE); }
	u32 n = 0, cap = 64;
	mapel* m = (mapel*) malloc (cap * sizeof (mapel));
	double a, b;
	/* the reference reads the file twice (count, then fill: map.c:429-455); line numbers in messages follow */
	while (read_value_pair (f, &a, &b)) n++;
	rewind (f);
	if (n > cap) { cap = n;  m = (mapel*) realloc (m, cap * sizeof (mapel)); }
	n = 0;
	while (read_value_pair (f, &a, &b)) { m[n].vIn = a;  m[n].vOut = b;  n++; }
	fclose (f);
	if (op->destroyFile) remove (op->filename);
	if (n == 0) { fprintf (stderr, "[%s] problem with mapping file \"%s\"\n", _op->name, op->filename);  exit (EXIT_FAILURE); }
	qsort (m, n, sizeof (mapel), mapel_ascending);
	double* in = (double*) malloc (n * sizeof (double));  double* out = (double*) malloc (n * sizeof (double));
	for (u32 k = 0; k < n; k++)
		{
		if (k > 0 && m[k].vIn == m[k-1].vIn)
			{
			fprintf (stderr, "[%s] \"%s\" lists the input value " valtypeFmt " more than once; the reference's result then\n"
			                 "depends on the order in which it meets the data, which this build does not reproduce\n",
			         _op->name, op->filename, m[k].vIn);
			exit (EXIT_FAILURE);
			}
		in[k] = m[k].vIn;  out[k] = m[k].vOut;
		}
	const gdsp_layout* lay = gd_layout_for (v, NULL);
	gd_check (gdsp_map_values (gd.ctx, lay, gd.sig, in, out, (int) n), _op->name);
	free (m);  free (in);  free (out);
	if (v == NULL && op->destroyFile && gd.nchrom > 1)
		{
		/* handed the whole genome at once; the reference would now fail to re-open the file it just
		 * deleted for the second chromosome (map.c:211-212) */
		fprintf (stderr, "[%s] can't open \"%s\" for reading\n", _op->name, op->filename);
		exit (EXIT_FAILURE);
		}
	}
