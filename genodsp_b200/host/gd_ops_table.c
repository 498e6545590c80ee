/* gd_ops_table.c -- the operator table: same names, aliases and order as the
 * reference's dspTable (genodsp.c:117-174), so `genodsp ?` lists the same
 * operators and every alias resolves to the same operator. */
#include <string.h>
#include <stdlib.h>
#include "gd_ops.h"

dspinfo dspTable[] =
	{
	dspinforecord("sum",           op_window_sum),      dspinfoalias("window_sum"),
	dspinforecord("slidingsum",    op_sliding_sum),     dspinfoalias("sliding_sum"),
	dspinforecord("smooth",        op_smooth),
	dspinforecord("cumulativesum", op_cumulative_sum),  dspinfoalias("cumulative"), dspinfoalias("integrate"),
	dspinforecord("clump",         op_clump),
	dspinforecord("anticlump",     op_skimp),           dspinfoalias("anti_clump"), dspinfoalias("skimp"),
	dspinforecord("percentile",    op_percentile),
	dspinforecord("add",           op_add),
	dspinforecord("subtract",      op_subtract),
	dspinforecord("addconst",      op_add_constant),    dspinfoalias("add_const"),
	dspinforecord("invert",        op_invert),
	dspinforecord("multiply",      op_multiply),
	dspinforecord("divide",        op_divide),
	dspinforecord("abs",           op_absolute_value),
	dspinforecord("mask",          op_mask),
	dspinforecord("masknot",       op_mask_not),        dspinfoalias("mask_not"),
	dspinforecord("clip",          op_clip),
	dspinforecord("erase",         op_erase),
	dspinforecord("binarize",      op_binarize),
	dspinforecord("or",            op_or),
	dspinforecord("and",           op_and),
	dspinforecord("maxover",       op_max_in_interval), dspinfoalias("max_over"),
	dspinforecord("minover",       op_min_in_interval), dspinfoalias("min_over"),
	dspinforecord("localmin",      op_local_minima),    dspinfoalias("local_min"),
	dspinforecord("localmax",      op_local_maxima),    dspinfoalias("local_max"),
	dspinforecord("bestmin",       op_best_local_min),  dspinfoalias("best_min"), dspinfoalias("bestlocalmin"), dspinfoalias("best_local_min"),
	dspinforecord("bestmax",       op_best_local_max),  dspinfoalias("best_max"), dspinfoalias("bestlocalmax"), dspinfoalias("best_local_max"),
	dspinforecord("minwith",       op_min_with),        dspinfoalias("min_with"),
	dspinforecord("maxwith",       op_max_with),        dspinfoalias("max_with"),
	dspinforecord("close",         op_close),
	dspinforecord("open",          op_open),
	dspinforecord("dilate",        op_dilate),
	dspinforecord("erode",         op_erode),
	dspinforecord("map",           op_map),
	dspinforecord("input",         op_input),
	dspinforecord("output",        op_output),
	dspinforecord("variables",     op_show_variables)
	};

const u32 dspTableLen = sizeof (dspTable) / sizeof (dspinfo);

/* every operator of this table accepts v==NULL */
int gd_is_genome_capable (dspop* op)
	{
	for (u32 ix = 0; ix < dspTableLen; ix++)
		if (dspTable[ix].funcApply != NULL && dspTable[ix].funcApply == op->funcApply) return true;
	return false;
	}

void op_short_line (char* name, int nameWidth, FILE* f, char* indent, const char* text)
	{
	int fill = nameWidth - 2 - (int) strlen (name);
	if (indent == NULL) indent = "";
	if (fill > 0) fprintf (f, "%s%s:%*s", indent, name, fill + 1, " ");
	         else fprintf (f, "%s%s: ", indent, name);
	fprintf (f, "%s\n", text);
	}

int arg_is_window (char* arg)
	{ return strcmp_prefix (arg, "--window=") == 0 || strcmp_prefix (arg, "W=") == 0 || strcmp_prefix (arg, "--W=") == 0; }

/* window / neighborhood size with the reference's checks and WARNING texts
 * (sum.c:117-135, sum.c:548-573, minmax.c:913-936) */
int parse_window_arg (char* name, char* arg, char* argVal, int minimum, int forceOdd, const char* what)
	{
	int w = string_to_unitized_int (argVal, true);
	if (w == 0) chastise ("[%s] %s can't be zero (\"%s\")\n", name, what, arg);
	if (w < 0)  chastise ("[%s] %s can't be negative (\"%s\")\n", name, what, arg);
	if (w < minimum)
		{
		fprintf (stderr, "[%s] WARNING: raising %s from %d to %d\n", name, what, w, minimum);
		w = minimum;
		}
	if (forceOdd && (w & 1) == 0)
		{
		fprintf (stderr, "[%s] WARNING: raising %s from %d to %d\n", name, what, w, w + 1);
		w++;
		}
	return w;
	}

void* op_alloc (char* name, size_t bytes)
	{
	void* p = calloc (1, bytes);
	if (p == NULL)
		{
		fprintf (stderr, "[%s] failed to allocate control record (%d bytes)\n", name, (int) bytes);
		exit (EXIT_FAILURE);
		}
	return p;
	}

void bad_arg (char* name, char* arg)
	{ chastise ("[%s] Can't understand \"%s\"\n", name, arg); }

int gd_is_pw_family (dspop* op);
void gd_resolve_variables (dspop* op)
	{
	if (gd_is_pw_family (op)) gd_resolve_pointwise (op);
	else gd_resolve_morph (op);
	}
