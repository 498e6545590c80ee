/* gd_ops_morph.c -- close, open, dilate, erode (reference morphology.c) and
 * clump, anticlump (reference clump.c).  Grammar and messages follow the
 * reference parsers (morphology.c:100-225, :400-525, :700-870, :1180-1325;
 * clump.c:129-258, :360-485). */
#include <stdlib.h>
#include <string.h>
#include "gd_ops.h"

typedef struct dspop_morph
	{
	dspop   common;
	int     kind;
	valtype length;              /* closing / opening / dilation / erosion length  */
	u32     left, right;         /* dilate / erode only                             */
	int     haveThreshold;
	char*   thresholdVarName;
	valtype threshold, oneVal, zeroVal;
	int     debug;
	} dspop_morph;

static const char* morphNoun[] = { "closing", "opening", "dilation", "erosion" };

static int is_opt (char* arg, const char* a, const char* b, const char* c)
	{ return strcmp_prefix (arg, a) == 0 || strcmp_prefix (arg, b) == 0 || strcmp_prefix (arg, c) == 0; }

static u32 side_length (char* argVal, int allowMinusOne)
	{
	if (allowMinusOne && strcmp_suffix (argVal, "-1") == 0)
		{
		char* t = copy_string (argVal);
		t[strlen (t) - 2] = 0;
		u32 v = (u32) (string_to_unitized_int (t, true) - 1);
		free (t);
		return v;
		}
	return allowMinusOne ? (u32) string_to_unitized_int (argVal, true) : (u32) string_to_valtype (argVal);
	}

static dspop* morph_parse (char* name, int argc, char** argv, int kind)
	{
	dspop_morph* op = (dspop_morph*) op_alloc (name, sizeof (dspop_morph));
	int haveLength = false;
	op->common.atRandom = false;
	op->kind = kind;  op->oneVal = 1.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (is_opt (arg, "--threshold=", "T=", "--T="))
			{
			if (op->haveThreshold)
				{ fprintf (stderr, "[%s] threshold specified more than once (at \"%s\")\n", name, arg);  exit (EXIT_FAILURE); }
			if (!try_string_to_valtype (argVal, &op->threshold)) op->thresholdVarName = copy_string (argVal);
			op->haveThreshold = true;
			}
		else if (is_opt (arg, "--one=", "O=", "--O="))   op->oneVal  = string_to_valtype (argVal);
		else if (is_opt (arg, "--zero=", "Z=", "--Z="))  op->zeroVal = string_to_valtype (argVal);
		else if (kind >= GDSP_MORPH_DILATE && strcmp_prefix (arg, "--left=") == 0)
			op->left  = side_length (argVal, kind == GDSP_MORPH_DILATE);
		else if (kind >= GDSP_MORPH_DILATE && strcmp_prefix (arg, "--right=") == 0)
			op->right = side_length (argVal, kind == GDSP_MORPH_DILATE);
		else if (strcmp (arg, "--debug") == 0 && kind != GDSP_MORPH_OPEN) op->debug = true;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (!haveLength) { op->length = (u32) string_to_unitized_int (arg, true);  haveLength = true; }
		else bad_arg (name, arg);
		}
	if (kind >= GDSP_MORPH_DILATE)
		{
		if (haveLength && (op->left != 0 || op->right != 0))
			{ fprintf (stderr, "[%s] %s length was provided in more than one way\n", name, morphNoun[kind]);  exit (EXIT_FAILURE); }
		if (!haveLength && op->left == 0 && op->right == 0)
			{ fprintf (stderr, "[%s] %s length was not provided\n", name, morphNoun[kind]);  exit (EXIT_FAILURE); }
		}
	else if (!haveLength)
		{ fprintf (stderr, "[%s] %s length was not provided\n", name, morphNoun[kind]);  exit (EXIT_FAILURE); }
	return (dspop*) op;
	}

static void morph_free (dspop* _op)
	{
	dspop_morph* op = (dspop_morph*) _op;
	if (op->thresholdVarName != NULL) free (op->thresholdVarName);
	free (op);
	}

static void morph_usage (char* name, FILE* f, char* indent, int kind)
	{
	static const char* what[] = {
		"Fill short gaps: positions above the threshold become 1; a gap between two\n%ssuch stretches that is no longer than <length> is filled with 1s as well; all\n%sother positions (including the gaps touching a chromosome end) become 0.\n",
		"Remove short stretches: a stretch of positions above the threshold becomes\n%s1s only when it is longer than <length>; everything else becomes 0.\n%s\n",
		"Widen stretches: every stretch of positions above the threshold is extended\n%sby about <length>/2 on each side (or by --left / --right); positions inside the\n%swidened stretches become 1, all others 0.\n",
		"Narrow stretches: every stretch of positions above the threshold is shortened\n%sby about <length>/2 on each side (or by --left / --right); positions inside\n%sthe narrowed stretches become 1, all others 0.\n" };
	if (indent == NULL) indent = "";
	fprintf (f, "%s", indent);
	fprintf (f, what[kind], indent, indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s <length> [options]\n", indent, name);
	fprintf (f, "%s  --threshold=<value>      (T=) positions above this are \"in\"; a number or the\n", indent);
	fprintf (f, "%s                           name of a variable (default is 0.0)\n", indent);
	if (kind >= GDSP_MORPH_DILATE)
		{
		fprintf (f, "%s  --left=<length>          amount applied on the left side only\n", indent);
		fprintf (f, "%s  --right=<length>         amount applied on the right side only\n", indent);
		}
	fprintf (f, "%s  --one=<value>            (O=) value written for \"in\" (default is 1.0)\n", indent);
	fprintf (f, "%s  --zero=<value>           (Z=) value written for \"out\" (default is 0.0)\n", indent);
	}

static void morph_resolve (dspop* _op)
	{
	dspop_morph* op = (dspop_morph*) _op;
	if (op->thresholdVarName != NULL)
		{
		if (!named_global_exists (op->thresholdVarName, &op->threshold))
			{
			fprintf (stderr, "[%s] attempt to use %s as threshold failed (no such variable)\n", _op->name, op->thresholdVarName);
			exit (EXIT_FAILURE);
			}
		fprintf (stderr, "[%s] using %s = " valtypeFmt " as threshold\n", _op->name, op->thresholdVarName, op->threshold);
		free (op->thresholdVarName);  op->thresholdVarName = NULL;
		}
	}

static void morph_apply (dspop* _op, valtype* v)
	{
	dspop_morph* op = (dspop_morph*) _op;
	morph_resolve (_op);
	u32 left = op->left, right = op->right;
	if (op->kind >= GDSP_MORPH_DILATE && left == 0 && right == 0)
		{
		left  = (u32) (op->length / 2);                 /* morphology.c:924-928, :1374-1378 */
		right = (u32) (op->length - left);
		}
	void* work = gd_work (gdsp_morph_work_bytes (gd.cells));
	gd_check (gdsp_morphology (gd.ctx, gd_layout_for (v, NULL), gd.sig, gd.cells, work, op->kind, op->length,
	                           left, right, op->threshold, op->oneVal, op->zeroVal), _op->name);
	}

#define MORPH_GROUP(fn, kind, text) \
void   fn##_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, text); } \
void   fn##_usage (char* name, FILE* f, char* indent) { morph_usage (name, f, indent, kind); } \
dspop* fn##_parse (char* name, int argc, char** argv) { return morph_parse (name, argc, argv, kind); } \
void   fn##_free  (dspop* op) { morph_free (op); } \
void   fn##_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { morph_apply (op, v); }

MORPH_GROUP (op_close,  GDSP_MORPH_CLOSE,  "fill short gaps between intervals (and binarize)")
MORPH_GROUP (op_open,   GDSP_MORPH_OPEN,   "remove short intervals (and binarize)")
MORPH_GROUP (op_dilate, GDSP_MORPH_DILATE, "widen intervals (and binarize)")
MORPH_GROUP (op_erode,  GDSP_MORPH_ERODE,  "narrow intervals (and binarize)")

/* ======================================================== clump / anticlump */

typedef struct dspop_clump
	{
	dspop   common;
	char*   averageVarName;
	valtype average;
	u32     minLength;
	double  relativeLength;
	valtype oneVal, zeroVal;
	int     debug, debugDetail, progress;
	int     exactOrder;          /* --exact-order (not in the reference): prefix sums in the reference's sequential order */
	} dspop_clump;

static void rel_fail (char* name, char* arg, const char* why)
	{ fprintf (stderr, "[%s] %s (at \"%s\")\n", name, why, arg);  exit (EXIT_FAILURE); }

/* <len> | CL | CL*f | f*CL | CL/d | max(a,b)   (clump.c:360-485) */
static void parse_min_length (char* name, char* arg, char* argVal, dspop_clump* op, int maxOk)
	{
	if (maxOk && strcmp_prefix (argVal, "max(") == 0 && strcmp_suffix (argVal, ")") == 0)
		{
		char* f1 = copy_string (argVal + 4);
		f1[strlen (f1) - 1] = 0;
		char* f2 = strchr (f1, ',');
		if (f2 == NULL) rel_fail (name, arg, "can't parse relative length");
		*(f2++) = 0;
		dspop_clump a, b;
		a.minLength = b.minLength = 0;  a.relativeLength = b.relativeLength = 0.0;
		parse_min_length (name, arg, f1, &a, false);
		parse_min_length (name, arg, f2, &b, false);
		if ((a.relativeLength > 0) == (b.relativeLength > 0)) rel_fail (name, arg, "can't parse relative length");
		if (a.relativeLength > 0) { op->relativeLength = a.relativeLength;  op->minLength = b.minLength; }
		                     else { op->relativeLength = b.relativeLength;  op->minLength = a.minLength; }
		free (f1);
		return;
		}
	if (strcmp (argVal, "CL") == 0) { op->relativeLength = 1.0;  op->minLength = 0;  return; }
	if (strcmp_prefix (argVal, "CL*") == 0 || strcmp_suffix (argVal, "*CL") == 0)
		{
		if (strcmp_prefix (argVal, "CL*") == 0) op->relativeLength = string_to_double (argVal + 3);
		else
			{
			char* t = copy_string (argVal);
			t[strlen (t) - 3] = 0;
			op->relativeLength = string_to_double (t);
			free (t);
			}
		if (op->relativeLength <= 0.0) rel_fail (name, arg, "relative length has to be positive");
		if (op->relativeLength >  1.0) rel_fail (name, arg, "relative length can't be more than 1");
		op->minLength = 0;
		return;
		}
	if (strcmp_prefix (argVal, "CL/") == 0)
		{
		op->relativeLength = string_to_double (argVal + 3);
		if (op->relativeLength < 0.0) rel_fail (name, arg, "relative length has to be positive");
		if (op->relativeLength < 1.0) rel_fail (name, arg, "relative length can't be more than 1");
		op->relativeLength = 1.0 / op->relativeLength;
		op->minLength = 0;
		return;
		}
	int len = string_to_unitized_int (argVal, true);
	if (len == 0) chastise ("[%s] minimum length can't be zero (\"%s\")\n", name, arg);
	if (len < 0)  chastise ("[%s] minimum length can't be negative (\"%s\")\n", name, arg);
	op->minLength = (u32) len;
	op->relativeLength = 0.0;
	}

static dspop* clump_parse (char* name, int argc, char** argv)
	{
	dspop_clump* op = (dspop_clump*) op_alloc (name, sizeof (dspop_clump));
	int haveAverage = false;
	op->common.atRandom = false;
	op->minLength = 100;  op->oneVal = 1.0;
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (is_opt (arg, "--average=", "T=", "--T="))
			{
			if (haveAverage)
				{ fprintf (stderr, "[%s] average threshold specified more than once (at \"%s\")\n", name, arg);  exit (EXIT_FAILURE); }
			op->averageVarName = copy_string (argVal);        /* always a variable name (clump.c:171) */
			haveAverage = true;
			}
		else if (is_opt (arg, "--length=", "L=", "--L=")) parse_min_length (name, arg, argVal, op, true);
		else if (is_opt (arg, "--one=", "O=", "--O="))    op->oneVal  = string_to_valtype (argVal);
		else if (is_opt (arg, "--zero=", "Z=", "--Z="))   op->zeroVal = string_to_valtype (argVal);
		else if (strcmp (arg, "--debug") == 0)            op->debug = true;
		else if (strcmp (arg, "--debug=detail") == 0)     op->debugDetail = true;
		else if (strcmp_prefix (arg, "--progress=") == 0) op->progress = string_to_unitized_int (argVal, true);
		else if (strcmp (arg, "--exact-order") == 0)      op->exactOrder = true;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (!haveAverage) { op->average = string_to_valtype (arg);  haveAverage = true; }
		else bad_arg (name, arg);
		}
	return (dspop*) op;
	}

static void clump_usage (char* name, FILE* f, char* indent, int above)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sFind intervals whose average is %s a threshold. Positions inside such\n", indent, above ? "at least" : "at most");
	fprintf (f, "%sintervals become 1, all others 0. Intervals of at least the given length\n", indent);
	fprintf (f, "%sare located, merged where they overlap, and their ends are trimmed of\n", indent);
	fprintf (f, "%spositions on the wrong side of the threshold (so a reported run may be\n", indent);
	fprintf (f, "%sshorter than the length).\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s <average> [options]\n", indent, name);
	fprintf (f, "%s  --average=<variable>     (T=) take the threshold from a named variable\n", indent);
	fprintf (f, "%s  --length=<length>        (L=) minimum interval length (default 100); also\n", indent);
	fprintf (f, "%s                           CL, CL*<f>, <f>*CL, CL/<d> (relative to the chromosome\n", indent);
	fprintf (f, "%s                           length) or max(<length>,<relative>)\n", indent);
	fprintf (f, "%s  --exact-order            (this implementation only) prefix sums in the sequential order\n", indent);
	fprintf (f, "%s                           of the original program: bit-identical on any real-valued\n", indent);
	fprintf (f, "%s                           signal, much slower\n", indent);
	fprintf (f, "%s  --one=<value>            (O=) value for positions in clumps (default 1.0)\n", indent);
	fprintf (f, "%s  --zero=<value>           (Z=) value for other positions (default 0.0)\n", indent);
	}

static void clump_resolve (dspop* _op)
	{
	dspop_clump* op = (dspop_clump*) _op;
	if (op->averageVarName != NULL)
		{
		if (!named_global_exists (op->averageVarName, &op->average))
			{
			fprintf (stderr, "[%s] attempt to use %s as threshold failed (no such variable)\n", _op->name, op->averageVarName);
			exit (EXIT_FAILURE);
			}
		fprintf (stderr, "[%s] using %s = " valtypeFmt " as threshold\n", _op->name, op->averageVarName, op->average);
		free (op->averageVarName);  op->averageVarName = NULL;
		}
	}

static void clump_apply (dspop* _op, valtype* v, int above)
	{
	dspop_clump* op = (dspop_clump*) _op;
	clump_resolve (_op);
	void* work = gd_work (gdsp_clump_work_bytes (gd.cells));
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 1), _op->name);
	gd_check (gdsp_clump (gd.ctx, gd_layout_for (v, NULL), gd.sig, gd.cells, work, op->average, op->minLength,
	                      op->relativeLength, above, op->oneVal, op->zeroVal), _op->name);
	if (op->exactOrder) gd_check (gdsp_ctx_set_exact_order (gd.ctx, 0), _op->name);
	}

static void clump_free (dspop* _op)
	{
	dspop_clump* op = (dspop_clump*) _op;
	if (op->averageVarName != NULL) free (op->averageVarName);
	free (op);
	}

void   op_clump_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find intervals with an average above some threshold"); }
void   op_clump_usage (char* name, FILE* f, char* indent) { clump_usage (name, f, indent, true); }
dspop* op_clump_parse (char* name, int argc, char** argv) { return clump_parse (name, argc, argv); }
void   op_clump_free  (dspop* op) { clump_free (op); }
void   op_clump_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { clump_apply (op, v, true); }

void   op_skimp_short (char* name, int w, FILE* f, char* indent) { op_short_line (name, w, f, indent, "find intervals with an average below some threshold"); }
void   op_skimp_usage (char* name, FILE* f, char* indent) { clump_usage (name, f, indent, false); }
dspop* op_skimp_parse (char* name, int argc, char** argv) { return clump_parse (name, argc, argv); }
void   op_skimp_free  (dspop* op) { clump_free (op); }
void   op_skimp_apply (dspop* op, arg_dont_complain(char* n), arg_dont_complain(u32 l), valtype* v) { clump_apply (op, v, false); }

void gd_resolve_morph (dspop* op)
	{
	if (op->funcApply == op_close_apply || op->funcApply == op_open_apply
	 || op->funcApply == op_dilate_apply || op->funcApply == op_erode_apply) morph_resolve (op);
	else if (op->funcApply == op_clump_apply || op->funcApply == op_skimp_apply) clump_resolve (op);
	}
