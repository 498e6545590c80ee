/* gd_ops_pointwise.c -- operators that act on each position independently:
 *   binarize (logical.c), addconst, abs, invert (add.c), clip, erase (mask.c)
 *   and the interval-file operators add, subtract (add.c), multiply, divide
 *   (multiply.c), mask, masknot (mask.c), or, and (logical.c).
 * Each of them describes itself as ONE gdsp_pw_op; the executor strings the
 * descriptors of consecutive operators together and launches one fused kernel
 * (gd_core.c run_pipeline), so a chain costs one pass over the genome. */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "gd_ops.h"

enum { PK_BINARIZE, PK_ADDCONST, PK_ABS, PK_INVERT, PK_CLIP, PK_ERASE,
       PK_ADD, PK_SUBTRACT, PK_MULTIPLY, PK_DIVIDE, PK_MASK, PK_MASKNOT, PK_OR, PK_AND,
       PK_MINOVER, PK_MAXOVER, PK_MINWITH, PK_MAXWITH };

typedef struct dspop_pw
	{
	dspop   common;
	int     kind;
	/* scalar parameters */
	valtype a, b, c;
	char   *varA, *varB;            /* names of variables still to be resolved       */
	int     haveA, haveB;
	int     flag;                   /* binarize: ties above; erase: keep inside      */
	/* interval-file operators */
	char*   filename;
	int     valColumn, originOne, destroyFile, debug;
	} dspop_pw;

static int  nPw = 0;
static struct { opfunc_apply apply;  int kind; } pwKinds[16];

static void pw_register (opfunc_apply apply, int kind)
	{
	for (int i = 0; i < nPw; i++) if (pwKinds[i].apply == apply) return;
	pwKinds[nPw].apply = apply;  pwKinds[nPw].kind = kind;  nPw++;
	}

int gd_is_pw_family (dspop* op);
int gd_is_pw_family (dspop* op)
	{
	return op->funcApply == op_binarize_apply || op->funcApply == op_clip_apply || op->funcApply == op_erase_apply;
	}

int gd_is_pointwise (dspop* op)
	{
	for (int i = 0; i < nPw; i++)
		if (pwKinds[i].apply == op->funcApply)
			{
			/* invert without an explicit middle needs the min/max of the signal as it is when the
			 * operator runs, so it cannot be folded into a chain with its predecessors */
			if (pwKinds[i].kind == PK_INVERT && !((dspop_pw*) op)->haveA) return false;
			return true;
			}
	return false;
	}

static int is_opt (char* arg, const char* a, const char* b, const char* c)
	{ return strcmp_prefix (arg, a) == 0 || strcmp_prefix (arg, b) == 0 || (c != NULL && strcmp_prefix (arg, c) == 0); }

static void resolve (dspop* _op, char** var, valtype* dst, const char* what, const char* whatFail)
	{
	if (*var == NULL) return;
	if (!named_global_exists (*var, dst))
		{
		fprintf (stderr, "[%s] attempt to use %s as %s failed (no such variable)\n", _op->name, *var, whatFail);
		exit (EXIT_FAILURE);
		}
	fprintf (stderr, "[%s] using %s = " valtypeFmt " as %s\n", _op->name, *var, *dst, what);
	free (*var);  *var = NULL;
	}

/* ---- interval file -> device table ------------------------------------------ */

/* true when the list is in layout order and no two intervals share a cell (one pass, no sort: a
 * file in any other order counts as "may overlap") */
static int ivlist_sorted_disjoint (const ivlist* l)
	{
	for (u64 k = 1; k < l->n; k++)
		{
		if (l->seg[k] < l->seg[k-1]) return false;
		if (l->seg[k] == l->seg[k-1] && l->start[k] < l->end[k-1]) return false;
		}
	return true;
	}

/* The reference adds interval after interval, cell by cell: ((v+a)+b) (add.c:280-281).  Integer values
 * on an integer-valued signal give the same bits in any order and go through the accumulate kernels
 * (one difference array: v + (a+b)).  Everything else -- real values, or integer values that overlap
 * on a signal holding non-integers -- is applied layer by layer in file order. */
static void add_file_now (dspop_pw* op, double sign)
	{
	ivlist l;
	ivlist_init (&l);
	ivlist_read_file (&l, op->common.name, op->filename, op->valColumn, op->originOne, true, false);
	int allInt = true;
	double sumAbs = 0.0;
	for (u64 k = 0; k < l.n; k++)
		{
		if (l.val[k] != floor (l.val[k]) || fabs (l.val[k]) > 1e6) allInt = false;
		sumAbs += fabs (l.val[k]);
		}
	int mode = (allInt && sumAbs < 2.0e9) ? GDSP_ACC_I32 : GDSP_ACC_F64;
	if (mode == GDSP_ACC_I32 && !ivlist_sorted_disjoint (&l))
		{
		u64 nonInt = 0;
		gd_check (gdsp_count_non_integer (gd.ctx, gd.genome, gd.sig, 4503599627370496.0 /* 2^52 */, &nonInt), op->common.name);
		if (nonInt != 0) mode = GDSP_ACC_F64;
		}
	/* the reference's additions (subtractions) cell by cell in file order */
	if (mode == GDSP_ACC_F64 && gd_apply_intervals_exact (&l, sign < 0 ? GD_EXACT_SUB : GD_EXACT_ADD, 0.0, op->common.name))
		{
		ivlist_free (&l);
		if (op->destroyFile) remove (op->filename);
		return;
		}
	if (sign < 0) for (u64 k = 0; k < l.n; k++) l.val[k] = -l.val[k];
	void* work = gd_work (gdsp_accumulate_work_bytes (gd.genome, gd.cells, mode));
	gd_check (gdsp_accumulate_host (gd.ctx, gd.genome, gd.sig, gd.cells, work, l.seg, l.start, l.end, l.val, l.n, mode, 1),
	          op->common.name);
	ivlist_free (&l);
	if (op->destroyFile) remove (op->filename);
	}

static gdsp_ivl_table* table_from_file (dspop_pw* op, int valCol, int skipZero, int requireSorted, int makeUnion)
	{
	ivlist l;
	ivlist_init (&l);
	ivlist_read_file (&l, op->common.name, op->filename, valCol, op->originOne, skipZero, requireSorted);
	if (makeUnion) ivlist_union (&l); else ivlist_sort (&l);
	gdsp_ivl_table* t;
	gd_check (gdsp_ivl_table_create (gd.ctx, gd.genome, l.seg, l.start, l.end, l.val, l.n, &t), op->common.name);
	ivlist_free (&l);
	return t;
	}

void gd_set_outside (ivlist* u, valtype value)
	{
	gdsp_ivl_table* t;
	gd_check (gdsp_ivl_table_create (gd.ctx, gd.genome, u->seg, u->start, u->end, u->val, u->n, &t), "input");
	gdsp_pw_op p;  memset (&p, 0, sizeof (p));
	p.code = GDSP_PW_IVL_SET_OUTSIDE;  p.a = value;  p.table = t;
	gd_check (gdsp_pointwise (gd.ctx, gd.genome, gd.sig, gd.sig, &p, 1), "input");
	gdsp_ivl_table_destroy (t);
	}

void gd_pw_release (gd_pw_resources* res)
	{ if (res->table != NULL) { gdsp_ivl_table_destroy (res->table);  res->table = NULL; } }

void gd_resolve_pointwise (dspop* _op)
	{
	dspop_pw* op = (dspop_pw*) _op;
	switch (op->kind)
		{
		case PK_BINARIZE: resolve (_op, &op->varA, &op->a, "threshold", "threshold");  break;
		case PK_CLIP: case PK_ERASE:
			resolve (_op, &op->varA, &op->a, "minimum limit", "minimum");
			resolve (_op, &op->varB, &op->b, "maximum limit", "maximum");
			break;
		default: break;
		}
	}

/* ---- descriptor ----------------------------------------------------------------- */

int gd_pointwise_descriptor (dspop* _op, gdsp_pw_op* out, gd_pw_resources* res)
	{
	dspop_pw* op = (dspop_pw*) _op;
	memset (out, 0, sizeof (*out));
	res->table = NULL;
	switch (op->kind)
		{
		case PK_BINARIZE:
			resolve (_op, &op->varA, &op->a, "threshold", "threshold");
			if (gd.pendingSorted)
				{
				/* directly after a percentile: the sorted genome it leaves is thresholded without being
				 * sorted (gd_ops_percentile.c); this binarize is then the first operator of its chain */
				int done = 0;
				for (int k = 0; k < gd.knownN; k++)
					/* +0.0 / -0.0 share a value but not a sort key: the key counts give the step only when every
					 * zero that ties with the threshold lies on the side the counts put it */
					if (gd.knownVal[k] == op->a && signbit (gd.knownVal[k]) == signbit (op->a)
					 && (op->a != 0.0 || ((signbit (op->a) != 0) == (op->flag != 0))))
						{
						/* the percentile's own selection pass counted the cells below / equal to this value */
						u64* prefix = (u64*) malloc (gd.nchrom * sizeof (u64));
						u64 acc = 0;
						for (int i = 0; i < gd.nchrom; i++) { prefix[i] = acc;  acc += gd.segs[i].hi - gd.segs[i].lo; }
						gd_check (gdsp_fill_step (gd.ctx, gd.genome, gd.sig, prefix,
						                          gd.knownBelow[k] + (op->flag ? 0 : gd.knownEqual[k]), op->b, op->c), _op->name);
						free (prefix);
						gd.pendingSorted = 0;
						return 0;
						}
				gd_check (gdsp_sorted_binarize (gd.ctx, gd.genome, gd.sig, op->a, op->flag, op->b, op->c, &done), _op->name);
				if (done) { gd.pendingSorted = 0;  return 0; }
				gd_materialise_sorted (_op->name);
				}
			out->code = op->flag ? GDSP_PW_BINARIZE_GE : GDSP_PW_BINARIZE_GT;
			out->a = op->a;  out->b = op->b;  out->c = op->c;
			return 1;
		case PK_ADDCONST:
			if (op->a == 0.0) return 0;                      /* add.c:734 */
			out->code = GDSP_PW_ADDCONST;  out->a = op->a;
			return 1;
		case PK_ABS:
			out->code = GDSP_PW_ABS;
			return 1;
		case PK_INVERT:
			{
			valtype mid = op->a;
			if (!op->haveA)
				{
				/* (min+max)/2 over the whole genome, add.c:907-926 */
				double mn, mx;  u64 cnt;
				gd_check (gdsp_minmax (gd.ctx, gd.genome, gd.sig, 1, -INFINITY, INFINITY, &mn, &mx, &cnt), _op->name);
				mid = (mn + mx) / 2.0;
				}
			out->code = GDSP_PW_INVERT;  out->a = 2 * mid;
			return 1;
			}
		case PK_CLIP:
			resolve (_op, &op->varA, &op->a, "minimum limit", "minimum");
			resolve (_op, &op->varB, &op->b, "maximum limit", "maximum");
			if (!op->haveB)      { out->code = GDSP_PW_CLIP_MIN;   out->a = op->a; }
			else if (!op->haveA) { out->code = GDSP_PW_CLIP_MAX;   out->a = op->b; }
			else                 { out->code = GDSP_PW_CLIP_BOTH;  out->a = op->a;  out->b = op->b; }
			return 1;
		case PK_ERASE:
			resolve (_op, &op->varA, &op->a, "minimum limit", "minimum");
			resolve (_op, &op->varB, &op->b, "maximum limit", "maximum");
			out->code = GDSP_PW_ERASE;  out->a = op->a;  out->b = op->b;  out->c = op->c;
			out->flags = (op->haveA ? GDSP_PW_ERASE_HAVE_MIN : 0) | (op->haveB ? GDSP_PW_ERASE_HAVE_MAX : 0)
			           | (op->flag ? GDSP_PW_ERASE_KEEP_INSIDE : 0);
			return 1;
		case PK_MULTIPLY:
			res->table = table_from_file (op, op->valColumn, true, true, false);
			out->code = GDSP_PW_IVL_MUL;  out->a = 0.0;  out->table = res->table;
			return 1;
		case PK_DIVIDE:
			res->table = table_from_file (op, op->valColumn, true, true, false);
			out->code = GDSP_PW_IVL_DIV;  out->a = op->a;  out->table = res->table;
			return 1;
		case PK_MASK:
			resolve (_op, &op->varA, &op->a, "mask value", "mask value");
			res->table = table_from_file (op, -1, false, false, true);
			out->code = GDSP_PW_IVL_SET;  out->a = op->a;  out->table = res->table;
			return 1;
		case PK_MASKNOT:
			res->table = table_from_file (op, -1, false, true, false);
			out->code = GDSP_PW_IVL_SET_OUTSIDE;  out->a = op->a;  out->table = res->table;
			return 1;
		default:
			break;
		}
	fprintf (stderr, "internal error: %s has no single pointwise descriptor\n", _op->name);
	exit (EXIT_FAILURE);
	}

/* or / and need two descriptors (non-zero -> 1, then the interval rule); add / subtract go through
 * the accumulate kernels: those four run on their own from their apply function */
static int fusable_kind (int kind)
	{ return kind != PK_OR && kind != PK_AND && kind != PK_ADD && kind != PK_SUBTRACT && kind < PK_MINOVER; }

static void pw_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), valtype* v)
	{
	dspop_pw* op = (dspop_pw*) _op;
	const gdsp_layout* lay = gd_layout_for (v, NULL);
	gdsp_pw_op prog[2];
	gd_pw_resources res;  res.table = NULL;
	int n = 0;
	memset (prog, 0, sizeof (prog));
	if (op->kind == PK_ADD || op->kind == PK_SUBTRACT) { add_file_now (op, op->kind == PK_ADD ? 1.0 : -1.0);  return; }
	if (op->kind == PK_OR || op->kind == PK_AND)
		{
		prog[0].code = GDSP_PW_NONZERO_TO_ONE;                          /* logical.c:466-473, :768-775 */
		res.table = (op->kind == PK_OR) ? table_from_file (op, op->valColumn, true, false, true)
		                                : table_from_file (op, op->valColumn, true, true, false);
		prog[1].code = (op->kind == PK_OR) ? GDSP_PW_IVL_SET : GDSP_PW_IVL_SET_OUTSIDE;
		prog[1].a = (op->kind == PK_OR) ? 1.0 : 0.0;
		prog[1].table = res.table;
		n = 2;
		}
	else if (op->kind == PK_MINOVER || op->kind == PK_MAXOVER)
		{
		/* minmax.c:197-419, :600-822: the extremum cell of every interval survives, everything else
		 * (rest of the interval, gaps, chromosomes absent from the file) takes the fill value */
		res.table = table_from_file (op, -1, true, true, false);
		gd_check (gdsp_ivl_arg_extrema (gd.ctx, lay, gd.sig, res.table, op->kind == PK_MAXOVER), _op->name);
		prog[0].code = GDSP_PW_IVL_KEEP_AT;  prog[0].a = op->a;  prog[0].table = res.table;
		n = 1;
		}
	else if (op->kind == PK_MINWITH || op->kind == PK_MAXWITH)
		{
		/* minmax.c:1979-1982, :2265-2268: intervals in any order, overlaps allowed; the host reduces
		 * them to disjoint pieces holding the smallest (largest) covering value */
		ivlist l, pieces;
		ivlist_init (&l);  ivlist_init (&pieces);
		ivlist_read_file (&l, _op->name, op->filename, op->valColumn, op->originOne, false, false);
		gd_paint_extreme (&l, op->kind == PK_MAXWITH, &pieces);
		gd_check (gdsp_ivl_table_create (gd.ctx, gd.genome, pieces.seg, pieces.start, pieces.end, pieces.val, pieces.n, &res.table), _op->name);
		ivlist_free (&l);  ivlist_free (&pieces);
		prog[0].code = (op->kind == PK_MINWITH) ? GDSP_PW_IVL_MIN : GDSP_PW_IVL_MAX;  prog[0].table = res.table;
		n = 1;
		if (op->destroyFile) remove (op->filename);
		}
	else n = gd_pointwise_descriptor (_op, &prog[0], &res);
	if (n > 0) gd_check (gdsp_pointwise (gd.ctx, lay, gd.sig, gd.sig, prog, n), _op->name);
	gd_pw_release (&res);
	}

/* distinct apply entry points so that the executor can recognise each operator's kind */
#define PW_APPLY(fn) void fn##_apply (dspop* op, char* n, u32 l, valtype* v) { pw_apply (op, n, l, v); }
PW_APPLY (op_binarize)  PW_APPLY (op_add_constant)  PW_APPLY (op_absolute_value)  PW_APPLY (op_invert)
PW_APPLY (op_clip)      PW_APPLY (op_erase)         PW_APPLY (op_add)             PW_APPLY (op_subtract)
PW_APPLY (op_multiply)  PW_APPLY (op_divide)        PW_APPLY (op_mask)            PW_APPLY (op_mask_not)
PW_APPLY (op_min_in_interval)  PW_APPLY (op_max_in_interval)  PW_APPLY (op_min_with)  PW_APPLY (op_max_with)
PW_APPLY (op_or)        PW_APPLY (op_and)

static dspop_pw* pw_new (char* name, int kind, opfunc_apply apply, int atRandom)
	{
	dspop_pw* op = (dspop_pw*) op_alloc (name, sizeof (dspop_pw));
	op->common.atRandom = atRandom;
	op->kind = kind;
	if (fusable_kind (kind)) pw_register (apply, kind);
	return op;
	}

static void pw_free (dspop* _op)
	{
	dspop_pw* op = (dspop_pw*) _op;
	free (op->varA);  free (op->varB);  free (op->filename);
	free (op);
	}

/* ================================================================= binarize */

void op_binarize_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "binarize the current set of interval values"); }
void op_binarize_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sBinarize the signal: positions above the threshold become one, all others\n", indent);
	fprintf (f, "%szero.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s [<threshold>] [options]\n", indent, name);
	fprintf (f, "%s  --threshold=<variable>   (T=) take the threshold from a named variable\n", indent);
	fprintf (f, "%s  --ties:below             values equal to the threshold become zero (default)\n", indent);
	fprintf (f, "%s  --ties:above             values equal to the threshold become one\n", indent);
	fprintf (f, "%s  --one=<value>            value for positions above the threshold (default 1.0)\n", indent);
	fprintf (f, "%s  --zero=<value>           value for the other positions (default 0.0)\n", indent);
	}
dspop* op_binarize_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_BINARIZE, op_binarize_apply, false);
	int haveThreshold = false;
	op->b = 1.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (is_opt (arg, "--threshold=", "T=", "--T="))
			{
			if (haveThreshold)
				{ fprintf (stderr, "[%s] threshold specified more than once (at \"%s\")\n", name, arg);  exit (EXIT_FAILURE); }
			op->varA = copy_string (argVal);  haveThreshold = true;            /* always a variable (logical.c:121) */
			}
		else if (strcmp (arg, "--ties:below") == 0 || strcmp (arg, "--ties=below") == 0) op->flag = false;
		else if (strcmp (arg, "--ties:above") == 0 || strcmp (arg, "--ties=above") == 0) op->flag = true;
		else if (is_opt (arg, "--one=", "O=", "--O="))  op->b = string_to_valtype (argVal);
		else if (is_opt (arg, "--zero=", "Z=", "--Z=")) op->c = string_to_valtype (argVal);
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (!haveThreshold) { op->a = string_to_valtype (arg);  haveThreshold = true; }
		else bad_arg (name, arg);
		}
	return (dspop*) op;
	}
void op_binarize_free (dspop* op) { pw_free (op); }

/* ================================================================= addconst */

void op_add_constant_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "add a constant to the current set of interval values"); }
void op_add_constant_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sAdd a constant to every position.\n%s\n%susage: %s <value>\n", indent, indent, indent, name);
	}
dspop* op_add_constant_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_ADDCONST, op_add_constant_apply, false);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (!op->haveA) { op->a = string_to_valtype (arg);  op->haveA = true; }
		else bad_arg (name, arg);
		}
	if (!op->haveA) { fprintf (stderr, "[%s] no constant was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}
void op_add_constant_free (dspop* op) { pw_free (op); }

/* ====================================================================== abs */

void op_absolute_value_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "take the absolute value of the current set of interval values"); }
void op_absolute_value_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReplace every position by its absolute value.\n%s\n%susage: %s\n", indent, indent, indent, name);
	}
dspop* op_absolute_value_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_ABS, op_absolute_value_apply, false);
	if (argc > 0) bad_arg (name, argv[0]);
	return (dspop*) op;
	}
void op_absolute_value_free (dspop* op) { pw_free (op); }

/* =================================================================== invert */

void op_invert_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "invert the current set of interval values"); }
void op_invert_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sReflect the signal about a middle value: v becomes 2*mid - v.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s [<mid>]\n", indent, name);
	fprintf (f, "%s  <mid>   zero|negate (0), one (1), 1/2|binary (0.5) or a number; by default\n", indent);
	fprintf (f, "%s          the midpoint of the minimum and maximum over the whole genome\n", indent);
	}
dspop* op_invert_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_INVERT, op_invert_apply, true);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		if (strcmp_prefix (arg, "--") == 0 || op->haveA) bad_arg (name, arg);
		if      (strcmp (arg, "zero") == 0 || strcmp (arg, "negate") == 0) op->a = 0.0;
		else if (strcmp (arg, "one") == 0)                                 op->a = 1.0;
		else if (strcmp (arg, "1/2") == 0 || strcmp (arg, "binary") == 0)  op->a = 0.5;
		else op->a = string_to_valtype (arg);
		op->haveA = true;
		}
	return (dspop*) op;
	}
void op_invert_free (dspop* op) { pw_free (op); }

/* ============================================================== clip / erase */

static void limits_parse (char* name, dspop_pw* op, int argc, char** argv, int isErase)
	{
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp_prefix (arg, "--min=") == 0 || strcmp_prefix (arg, "--minimum=") == 0)
			{
			if (!try_string_to_valtype (argVal, &op->a)) op->varA = copy_string (argVal);
			op->haveA = true;
			}
		else if (strcmp_prefix (arg, "--max=") == 0 || strcmp_prefix (arg, "--maximum=") == 0)
			{
			if (!try_string_to_valtype (argVal, &op->b)) op->varB = copy_string (argVal);
			op->haveB = true;
			}
		else if (isErase && (strcmp (arg, "--keep:outside") == 0 || strcmp (arg, "--keep=outside") == 0)) op->flag = false;
		else if (isErase && (strcmp (arg, "--keep:inside") == 0  || strcmp (arg, "--keep=inside") == 0))  op->flag = true;
		else if (isErase && is_opt (arg, "--zero=", "Z=", "--Z=")) op->c = string_to_valtype (argVal);
		else bad_arg (name, arg);
		}
	if (!op->haveA && !op->haveB) chastise ("[%s] neither min nor max was provided\n", name);
	if (op->haveA && op->haveB && op->varA == NULL && op->varB == NULL && op->a > op->b)
		chastise ("[%s] min can't be greater than max (" valtypeFmt ">" valtypeFmt ")\n", name, op->a, op->b);
	}

void op_clip_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "clip the current set of interval values"); }
void op_clip_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sLimit the signal to a range: values below the minimum are raised to it,\n", indent);
	fprintf (f, "%svalues above the maximum lowered to it.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --min=<value>            lower limit (a number or a variable name)\n", indent);
	fprintf (f, "%s  --max=<value>            upper limit (a number or a variable name)\n", indent);
	}
dspop* op_clip_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_CLIP, op_clip_apply, false);
	limits_parse (name, op, argc, argv, false);
	return (dspop*) op;
	}
void op_clip_free (dspop* op) { pw_free (op); }

void op_erase_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "erase any of the current set of interval values that are within (or outside) some range"); }
void op_erase_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sErase (set to the zero value) the positions whose value lies inside a\n", indent);
	fprintf (f, "%srange; with --keep:inside, the positions whose value lies outside it.\n%s\n", indent, indent);
	fprintf (f, "%susage: %s [options]\n", indent, name);
	fprintf (f, "%s  --min=<value>            lower end of the range (number or variable name)\n", indent);
	fprintf (f, "%s  --max=<value>            upper end of the range (number or variable name)\n", indent);
	fprintf (f, "%s  --keep:outside           erase values inside the range (default)\n", indent);
	fprintf (f, "%s  --keep:inside            erase values outside the range\n", indent);
	fprintf (f, "%s  --zero=<value>           (Z=) value written to erased positions (default 0.0)\n", indent);
	}
dspop* op_erase_parse (char* name, int argc, char** argv)
	{
	dspop_pw* op = pw_new (name, PK_ERASE, op_erase_apply, false);
	limits_parse (name, op, argc, argv, true);
	return (dspop*) op;
	}
void op_erase_free (dspop* op) { pw_free (op); }

/* ===================================================== interval-file operators */

static void file_usage (char* name, FILE* f, char* indent, const char* what, int hasValue, const char* extra)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%s%s\n%s\n", indent, what, indent);
	fprintf (f, "%susage: %s <filename> [options]\n", indent, name);
	if (hasValue)
		fprintf (f, "%s  --value=<col>            column of the file that holds the value (default 4)\n", indent);
	if (hasValue == 1)
		fprintf (f, "%s  --novalue                the file has no value column (every value is 1)\n", indent);
	fprintf (f, "%s  --origin=one             the file's intervals are origin-one, closed\n", indent);
	fprintf (f, "%s  --origin=zero            the file's intervals are origin-zero, half-open\n", indent);
	if (extra != NULL) fprintf (f, "%s%s\n", indent, extra);
	}

/* valueOpts: 0 none, 1 --value/--novalue;  destroyOpt: --destroy accepted */
static dspop* file_parse (char* name, int argc, char** argv, int kind, opfunc_apply apply,
                          int valueOpts, int destroyOpt)
	{
	dspop_pw* op = pw_new (name, kind, apply, true);
	op->valColumn = (int) get_named_global ("valColumn", 4-1);
	op->originOne = (int) get_named_global ("originOne", false);
	if (kind == PK_DIVIDE || kind == PK_MINOVER) op->a = valtypeMax;
	if (kind == PK_MAXOVER) op->a = 0.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (valueOpts == 1 && (strcmp (arg, "--novalue") == 0 || strcmp (arg, "--novalues") == 0 || strcmp (arg, "--value=none") == 0))
			op->valColumn = -1;
		else if (valueOpts && strcmp_prefix (arg, "--value=") == 0)
			{
			int col = string_to_int (argVal) - 1;
			if (col == -1) chastise ("[%s] value column can't be 0 (\"%s\")\n", name, arg);
			if (col < 0)   chastise ("[%s] value column can't be negative (\"%s\")\n", name, arg);
			if (col < 3)   chastise ("[%s] value column can't be 1, 2 or 3 (\"%s\")\n", name, arg);
			op->valColumn = col;
			}
		else if (strcmp (arg, "--origin=one") == 0 || strcmp (arg, "--origin=1") == 0)  op->originOne = true;
		else if (strcmp (arg, "--origin=zero") == 0 || strcmp (arg, "--origin=0") == 0) op->originOne = false;
		else if (destroyOpt && strcmp (arg, "--destroy") == 0) op->destroyFile = true;
		else if ((kind == PK_DIVIDE || kind == PK_MINOVER) && strcmp_prefix (arg, "--infinity=") == 0) op->a = string_to_valtype (argVal);
		else if (kind == PK_MAXOVER && is_opt (arg, "--zero=", "Z=", "--Z=")) op->a = string_to_valtype (argVal);
		else if ((kind == PK_MULTIPLY || kind == PK_DIVIDE || kind == PK_MINOVER || kind == PK_MAXOVER) && strcmp (arg, "--debug") == 0) op->debug = true;
		else if ((kind == PK_MASK || kind == PK_MASKNOT) && is_opt (arg, "--mask=", "M=", "--M="))
			{
			if (op->haveA)
				{ fprintf (stderr, "[%s] mask value specified more than once (at \"%s\")\n", name, arg);  exit (EXIT_FAILURE); }
			if (kind == PK_MASK) { if (!try_string_to_valtype (argVal, &op->a)) op->varA = copy_string (argVal); }
			else op->a = string_to_valtype (argVal);
			op->haveA = true;
			}
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (op->filename == NULL) op->filename = copy_string (arg);
		else bad_arg (name, arg);
		}
	if (op->filename == NULL) { fprintf (stderr, "[%s] no filename was provided\n", name);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}

#define FILE_GROUP(fn, kind, text, what, valueOpts, destroyOpt, extra) \
void   fn##_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, text); } \
void   fn##_usage (char* name, FILE* f, char* indent) { file_usage (name, f, indent, what, valueOpts, extra); } \
dspop* fn##_parse (char* name, int argc, char** argv) { return file_parse (name, argc, argv, kind, fn##_apply, valueOpts, destroyOpt); } \
void   fn##_free  (dspop* op) { pw_free (op); }

FILE_GROUP (op_add,      PK_ADD,      "add interval values (read from a file) to the current set of interval values",
            "Add the values of the intervals in a file to the signal (overlapping\n  intervals add up).", 1, 1,
            "  --destroy                delete the file after reading it")
FILE_GROUP (op_subtract, PK_SUBTRACT, "subtract interval values (read from a file) from the current set of interval values",
            "Subtract the values of the intervals in a file from the signal.", 1, 1,
            "  --destroy                delete the file after reading it")
FILE_GROUP (op_multiply, PK_MULTIPLY, "multiply the current set of interval values by interval values read from a file",
            "Multiply the signal by the values of the intervals in a file; positions not\n  covered by the file become zero. The file must be sorted along each chromosome\n  and its intervals must not overlap.", 1, 0, NULL)
FILE_GROUP (op_divide,   PK_DIVIDE,   "divide the current set of interval values by interval values read from a file",
            "Divide the signal by the values of the intervals in a file; positions not\n  covered by the file become +/- infinity. The file must be sorted along each\n  chromosome and its intervals must not overlap.", 1, 0,
            "  --infinity=<value>       value used for infinity (default is the largest double)")
FILE_GROUP (op_mask,     PK_MASK,     "mask the current set of interval values by intervals read from a file",
            "Set the signal to the mask value inside the intervals of a file.", 0, 0,
            "  --mask=<value>           (M=) mask value, a number or a variable name (default 0.0)")
FILE_GROUP (op_mask_not, PK_MASKNOT,  "mask the current set of interval values by the complement of intervals read from a file",
            "Set the signal to the mask value outside the intervals of a file. The file\n  must be sorted along each chromosome and its intervals must not overlap.", 0, 0,
            "  --mask=<value>           (M=) mask value (default 0.0)")
FILE_GROUP (op_or,       PK_OR,       "perform logical OR of the current set of interval values with intervals read from a file",
            "Logical OR: non-zero positions become 1, and so do the positions covered by\n  the file's intervals (intervals whose value is 0 are ignored).", 1, 0, NULL)
FILE_GROUP (op_and,      PK_AND,      "perform logical AND of the current set of interval values with intervals read from a file",
            "Logical AND: non-zero positions become 1, and positions not covered by the\n  file's intervals become 0. The file must be sorted along each chromosome and\n  its intervals must not overlap.", 1, 0, NULL)

FILE_GROUP (op_min_in_interval, PK_MINOVER, "find the minimum value in each of a set of intervals read from a file",
            "Keep, in every interval of a file, only the position holding the minimum; all\n  other positions (inside and outside the intervals) become infinity. Ties go to\n  the position nearest the centre of the interval. The file must be sorted along\n  each chromosome and its intervals must not overlap.", 0, 0,
            "  --infinity=<value>       value given to every other position (default is the largest double)")
FILE_GROUP (op_max_in_interval, PK_MAXOVER, "find the maximum value in each of a set of intervals read from a file",
            "Keep, in every interval of a file, only the position holding the maximum; all\n  other positions (inside and outside the intervals) become zero. Ties go to the\n  position nearest the centre of the interval. The file must be sorted along\n  each chromosome and its intervals must not overlap.", 0, 0,
            "  --zero=<value>           (Z=) value given to every other position (default 0.0)")
FILE_GROUP (op_min_with, PK_MINWITH, "take the minimum of the current set of interval values and values read from a file",
            "Position by position, keep the smaller of the signal and the values of the\n  intervals of a file covering that position.", 2, 1,
            "  --destroy                delete the file after reading it")
FILE_GROUP (op_max_with, PK_MAXWITH, "take the maximum of the current set of interval values and values read from a file",
            "Position by position, keep the larger of the signal and the values of the\n  intervals of a file covering that position.", 2, 1,
            "  --destroy                delete the file after reading it")
