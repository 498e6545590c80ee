/* gd_ops_percentile.c -- the percentile operator (reference percentile.c).
 * Grammar, variable names, report lines and the bash/map outputs follow
 * op_percentile_parse (:131-375), op_percentile_apply (:392-751) and
 * set_percentile_name (:756-780).  The order statistics come from
 * gdsp_percentiles (exact selection, no sort).  The reference's percentile is
 * destructive; when a later operator (or the final output) reads the signal,
 * the reference's post-state is materialised:
 *   - --window>1 / --min / --max first bring the qualifying values to the front with the
 *     reference's collect permutation (gdsp_percentile_collect, percentile.c:547-580);
 *   - the front (every chromosome when all positions qualify) is then left as the reference's
 *     sorts and bubble passes leave it (percentile.c:611-651): chromosomes 0..K (K = chromosome
 *     holding the last reported rank) become the sorted prefix.  When K is one of the last two
 *     that is one global sort (thresholded without sorting when `binarize` follows); otherwise
 *     every front chromosome is sorted on its own and each pass step (combine_sorted_vectors)
 *     is a split search and two merges on the device (gdsp_merge_exchange). */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "gd_ops.h"

#define percentileStepUnits 1000

typedef struct dspop_percentile
	{
	dspop   common;
	char*   preserveFilename;
	char*   mapFilename;
	u32     percentileLo, percentileHi, percentileStep;
	u32     windowSize;
	valtype minAllowed, maxAllowed;
	int     valPrecision, reportForBash, quiet, debug, debugShowIndex, debugStage;
	} dspop_percentile;

void op_percentile_short (char* name, int w, FILE* f, char* indent)
	{ op_short_line (name, w, f, indent, "compute percentiles of the current set of interval values"); }

void op_percentile_usage (char* name, FILE* f, char* indent)
	{
	if (indent == NULL) indent = "";
	fprintf (f, "%sCompute percentiles of the signal over the whole genome and store them in\n", indent);
	fprintf (f, "%snamed variables (percentile 99 -> \"percentile99\") for later operators.\n", indent);
	fprintf (f, "%sThe signal itself is left in an unspecified order unless --preserve is used.\n", indent);
	fprintf (f, "%s\n", indent);
	fprintf (f, "%susage: %s <percentile> [options]\n", indent, name);
	fprintf (f, "%s  <percentile>             <lo>, <lo>..<hi>[by<step>] or <lo>,<hi>\n", indent);
	fprintf (f, "%s  --step=<value>           step between reported percentiles (default 1)\n", indent);
	fprintf (f, "%s  --window=<length>        (W=) use only every <length>th position (default 1)\n", indent);
	fprintf (f, "%s  --min=<value>            ignore values below this\n", indent);
	fprintf (f, "%s  --max=<value>            ignore values above this\n", indent);
	fprintf (f, "%s  --precision=<number>     digits used when reporting the values\n", indent);
	fprintf (f, "%s  --preserve=<filename>    keep the signal unchanged (the file is scratch)\n", indent);
	fprintf (f, "%s  --map=<filename>         also write \"value percentile\" lines to a file\n", indent);
	fprintf (f, "%s  --report:bash            report as bash variable assignments on stdout\n", indent);
	fprintf (f, "%s  --quiet                  don't report the values on stderr\n", indent);
	}

static u32 pct_units (valtype v)
	{
	if (v < 0.0)   return 0;
	if (v > 100.0) return 100 * percentileStepUnits;
	return (u32) (int) (percentileStepUnits * v + .5);
	}

dspop* op_percentile_parse (char* name, int argc, char** argv)
	{
	dspop_percentile* op = (dspop_percentile*) op_alloc (name, sizeof (dspop_percentile));
	int haveRange = false;
	op->common.atRandom = true;
	op->percentileStep = percentileStepUnits;
	op->windowSize   = (u32) get_named_global ("windowSize", 1);
	op->minAllowed   = -valtypeMax;
	op->maxAllowed   =  valtypeMax;
	op->valPrecision = (int) get_named_global ("valPrecision", 0);
	for (; argc > 0; argv++, argc--)
		{
		char* arg = argv[0];
		char* argVal = strchr (arg, '=');  if (argVal != NULL) argVal++;
		if (strcmp_prefix (arg, "--step=") == 0)
			{
		set_step:;
			valtype st = string_to_valtype (argVal);
			if (st == 0) chastise ("[%s] step can't be zero (\"%s\")\n", name, arg);
			if (st < 0)  chastise ("[%s] step can't be negative (\"%s\")\n", name, arg);
			if (st < .001) st = .001;
			op->percentileStep = (u32) (percentileStepUnits * st + .5);
			}
		else if (arg_is_window (arg))
			{
			int w = string_to_unitized_int (argVal, true);
			if (w == 0) w = 1;
			if (w < 0) chastise ("[%s] window size can't be negative (\"%s\")\n", name, arg);
			op->windowSize = (u32) w;
			}
		else if (strcmp_prefix (arg, "--min=") == 0) op->minAllowed = string_to_valtype (argVal);
		else if (strcmp_prefix (arg, "--max=") == 0) op->maxAllowed = string_to_valtype (argVal);
		else if (strcmp_prefix (arg, "--precision=") == 0)
			{
			int p = string_to_int (argVal);
			if (p < 0) chastise ("[%s] precision can't be negative (\"%s\")\n", name, arg);
			op->valPrecision = p;
			}
		else if (strcmp_prefix (arg, "--preserve=") == 0)
			{
			if (op->preserveFilename != NULL && strcmp (argVal, op->preserveFilename) != 0)
				chastise ("[%s] can't specify two files for data preservation\n(\"%s\" and \"%s\")", name, op->preserveFilename, argVal);
			free (op->preserveFilename);
			op->preserveFilename = copy_string (argVal);
			}
		else if (strcmp_prefix (arg, "--map=") == 0 || strcmp_prefix (arg, "--mapping=") == 0)
			{
			if (op->mapFilename != NULL && strcmp (argVal, op->mapFilename) != 0)
				chastise ("[%s] can't specify two files for mapping\n(\"%s\" and \"%s\")", name, op->mapFilename, argVal);
			free (op->mapFilename);
			op->mapFilename = copy_string (argVal);
			}
		else if (strcmp (arg, "--report:bash") == 0 || strcmp (arg, "--bash") == 0) op->reportForBash = true;
		else if (strcmp (arg, "--quiet") == 0 || strcmp (arg, "--silent") == 0)     op->quiet = true;
		else if (strcmp (arg, "--debug") == 0)         op->debug = true;
		else if (strcmp (arg, "--debug=index") == 0)   op->debugShowIndex = true;
		else if (strcmp (arg, "--debug=collect") == 0) op->debugStage = 1;
		else if (strcmp (arg, "--debug=sort1") == 0)   op->debugStage = 2;
		else if (strcmp (arg, "--debug=sort2") == 0)   op->debugStage = 3;
		else if (strcmp_prefix (arg, "--") == 0) bad_arg (name, arg);
		else if (!haveRange && strchr (arg, ',') != NULL)
			{
			char* second = strchr (arg, ',');
			*(second++) = 0;
			valtype lo = string_to_valtype (arg), hi = string_to_valtype (second);
			if (lo > hi) { valtype t = hi;  hi = lo;  lo = t; }
			op->percentileLo = pct_units (lo);  op->percentileHi = pct_units (hi);
			haveRange = true;
			op->percentileStep = (op->percentileLo < op->percentileHi) ? op->percentileHi - op->percentileLo : 1;
			}
		else if (!haveRange)
			{
			char* second = strstr (arg, "..");
			char* by = NULL;
			if (second != NULL) { *second = 0;  second += 2;  by = strstr (second, "by");  if (by != NULL) { *by = 0;  by += 2; } }
			valtype lo = string_to_valtype (arg), hi = (second != NULL) ? string_to_valtype (second) : lo;
			if (lo > hi) { valtype t = hi;  hi = lo;  lo = t; }
			op->percentileLo = pct_units (lo);  op->percentileHi = pct_units (hi);
			haveRange = true;
			if (by != NULL) { argVal = by;  goto set_step; }
			}
		else bad_arg (name, arg);
		}
	if (!haveRange) { fprintf (stderr, "[%s] no range of percentiles was provided\n", name);  exit (EXIT_FAILURE); }
	if (op->reportForBash && op->quiet) chastise ("[%s] Can't use both --report:bash and --quiet\n", name);
	return (dspop*) op;
	}

void op_percentile_free (dspop* _op)
	{
	dspop_percentile* op = (dspop_percentile*) _op;
	free (op->preserveFilename);  free (op->mapFilename);
	free (op);
	}

static void set_percentile_name (char* varName, u32 percentile)
	{
	if (percentile % percentileStepUnits == 0)
		{ sprintf (varName, "percentile%d", percentile / percentileStepUnits);  return; }
	float pPct = percentile / ((float) percentileStepUnits);
	int precision = 1;
	for (u32 denom = percentileStepUnits / 10; denom >= 1; precision++, denom /= 10)
		if (percentile % denom == 0) { sprintf (varName, "percentile%.*f", precision, pPct);  return; }
	sprintf (varName, "percentile%f", pPct);
	}

/* does anything read the signal after this operator? */
static int signal_is_read_later (dspop* op)
	{
	if (op->next != NULL) return strcmp (op->next->name, "input") != 0;
	return !gd_output_inhibited ();
	}

void gd_materialise_sorted (const char* who)
	{
	if (!gd.pendingSorted) return;
	gd.pendingSorted = 0;
	int inTmp = 0;
	gd_check (gdsp_sort_genome (gd.ctx, gd.genome, gd.sig, gd.tmp, gd.cells, &inTmp), who);
	if (inTmp) gd_swap ();
	}

void op_percentile_apply (dspop* _op, arg_dont_complain(char* vName), arg_dont_complain(u32 vLen), arg_dont_complain(valtype* v))
	{
	dspop_percentile* op = (dspop_percentile*) _op;
	char varName[100];
	const u32 full = 100 * percentileStepUnits;

	/* percentile 0, 100 and 0..100 are plain min/max scans and leave the signal alone (:434-530) */
	if (op->mapFilename == NULL
	 && ((op->percentileLo == 0 && op->percentileHi == 0) || (op->percentileLo == full && op->percentileHi == full)
	  || (op->percentileLo == 0 && op->percentileHi == full)))
		{
		double mn, mx;  u64 cnt;
		gd_check (gdsp_minmax (gd.ctx, gd.genome, gd.sig, op->windowSize, op->minAllowed, op->maxAllowed, &mn, &mx, &cnt), _op->name);
		if (cnt == 0) goto no_values;
		if (op->percentileHi == full) { set_percentile_name (varName, op->percentileHi);  set_named_global (varName, mx); }
		if (op->percentileLo == 0)    { set_percentile_name (varName, op->percentileLo);  set_named_global (varName, mn); }
		return;
		}

	/* --preserve (percentile.c:534-535): the reference writes the vectors to the scratch file here ... */
	if (op->preserveFilename != NULL)
		{
		FILE* f = fopen (op->preserveFilename, "wt");
		if (f == NULL) { fprintf (stderr, "can't open \"%s\" for writing\n", op->preserveFilename);  exit (EXIT_FAILURE); }
		fclose (f);
		if (trackOperations) fprintf (stderr, "write_all(%s)\n", op->preserveFilename);
		}

	FILE* mapF = (op->mapFilename != NULL) ? fopen (op->mapFilename, "wt") : NULL;

	u32 np = 0;
	for (u32 p = op->percentileLo; p <= op->percentileHi; p += op->percentileStep) np++;
	u32*    pm   = (u32*) malloc (np * sizeof (u32));
	double* vals = (double*) malloc (np * sizeof (double));
	np = 0;
	for (u32 p = op->percentileLo; p <= op->percentileHi; p += op->percentileStep) pm[np++] = p;
	u64 numValues = 0, numNan = 0;
	u64* below = (u64*) malloc (np * sizeof (u64));
	u64* equal = (u64*) malloc (np * sizeof (u64));
	gd_check (gdsp_percentiles_ranked (gd.ctx, gd.genome, gd.sig, gd.tmp, gd.cells, op->windowSize, op->minAllowed,
	                                   op->maxAllowed, pm, (int) np, vals, &numValues, below, equal, &numNan), _op->name);
	gd.knownN = 0;
	if (numNan == 0 && numValues == gdsp_layout_cells (gd.genome) && np <= 1024)
		{
		gd.knownN = (int) np;
		for (u32 i = 0; i < np; i++) { gd.knownVal[i] = vals[i];  gd.knownBelow[i] = below[i];  gd.knownEqual[i] = equal[i]; }
		}
	free (below);  free (equal);
	if (numValues == 0) { free (pm);  free (vals);  if (mapF) fclose (mapF);  goto no_values; }

	u64 lastRank = 0;
	for (u32 i = 0; i < np; i++)
		{
		float pPct = pm[i] / ((float) percentileStepUnits);
		u64 rank = (pm[i] >= full) ? numValues - 1 : (u64) (u32) (((u64) (u32) numValues) * pm[i] / (100.0 * percentileStepUnits));
		lastRank = rank;
		set_percentile_name (varName, pm[i]);
		set_named_global (varName, vals[i]);
		if (op->reportForBash && pm[i] < full)
			fprintf (stdout, "%s=" valtypeFmtPrec " # bash command\n", varName, op->valPrecision, vals[i]);
		else if (!op->quiet)
			{
			fprintf (stderr, "percentile %.3f is ", pPct);
			if (op->debugShowIndex) fprintf (stderr, "[%llu] ", (unsigned long long) rank);
			fprintf (stderr, valtypeFmtPrec "\n", op->valPrecision, vals[i]);
			}
		if (mapF != NULL) fprintf (mapF, valtypeFmtPrec " %.3f\n", op->valPrecision, vals[i], pPct);
		}
	free (pm);  free (vals);
	if (mapF != NULL) fclose (mapF);

	/* the reference leaves the genome permuted; reproduce that state only if someone will look */
	if (op->preserveFilename != NULL)
		{
		/* ... and reads them back here (percentile.c:717-725): every value returns as the double nearest
		 * to its 10-decimal text, every zero as +0.0.  gdsp_text_roundtrip does that on the device. */
		if (trackOperations) fprintf (stderr, "read_all(%s)\n", op->preserveFilename);
		gd_check (gdsp_text_roundtrip (gd.ctx, gd.genome, gd.sig, 10), _op->name);
		FILE* f = fopen (op->preserveFilename, "wb");      /* the reference leaves an empty scratch file */
		if (f != NULL) fclose (f);
		return;
		}
	if (!signal_is_read_later (_op)) return;
	{
	/* The reference leaves the vectors permuted (percentile.c:547-651).  `front` = the part of the
	 * concatenated chromsSorted genome that holds the qualifying samples: every chromosome before
	 * `last` whole, the first numInLast cells of chromosome `last`. */
	const int filtered = (op->windowSize != 1 || op->minAllowed != -valtypeMax || op->maxAllowed != valtypeMax);
	int last = gd.nchrom - 1;
	u64 numInLast = gd.segs[last].hi - gd.segs[last].lo;
	if (filtered)
		{
		/* the collect pass: qualifying samples to the front, the displaced values shuffled behind them */
		u64 n = 0;
		void* work = gd_work (gdsp_percentile_collect_work_bytes (gd.cells));
		gd_check (gdsp_percentile_collect (gd.ctx, gd.genome, gd.sig, gd.tmp, gd.cells, work, op->windowSize,
		                                   op->minAllowed, op->maxAllowed, &n), _op->name);
		gd_swap ();
		u64 acc = 0;
		for (int i = 0; i < gd.nchrom; i++)
			{
			const u64 len = gd.segs[i].hi - gd.segs[i].lo;
			if (n <= acc + len) { last = i;  numInLast = n - acc;  break; }
			acc += len;
			}
		}
	gdsp_seg* front = (gdsp_seg*) malloc ((last + 1) * sizeof (gdsp_seg));
	for (int i = 0; i <= last; i++) front[i] = gd.segs[i];
	front[last].hi = front[last].lo + numInLast;
	front[last].dhi = front[last].hi;

	/* K = chromosome (sorted order) that holds the last reported rank */
	u64 acc = 0;  int K = last;
	for (int i = 0; i <= last; i++)
		{ acc += front[i].hi - front[i].lo;  if (lastRank < acc) { K = i;  break; } }
	if (K >= last - 1)
		{
		/* the front ends up globally sorted (chromosome `last` alone holds the leftovers of the passes) */
		if (!filtered)
			{
			/* `percentile P = binarize ...` thresholds it straight away: a step function that needs a count,
			 * not a sort -- leave the state pending and let binarize decide (not under --progress=operations,
			 * whose trace interleaves binarize's messages differently) */
			gd.pendingSorted = 1;
			if (!(_op->next != NULL && _op->next->funcApply == op_binarize_apply && !trackOperations))
				gd_materialise_sorted (_op->name);
			}
		else
			{
			int inTmp = 0;
			gdsp_layout* lay;
			gd_check (gdsp_layout_create (gd.ctx, front, last + 1, &lay), _op->name);
			gd_check (gdsp_sort_genome (gd.ctx, lay, gd.sig, gd.tmp, gd.cells, &inTmp), _op->name);
			if (inTmp)
				for (int q = 0; q <= last; q++)
					gd_check (gdsp_d2d (gd.ctx, gd.sig + front[q].lo, gd.tmp + front[q].lo,
					                    (front[q].hi - front[q].lo) * sizeof (double)), _op->name);
			gdsp_layout_destroy (lay);
			}
		}
	else
		{
		/* the reference's passes (percentile.c:611-651): every front chromosome is sorted on its own, then
		 * chromosome c takes the smallest len(c) values of {c, d} for every later d of the front, in order.
		 * Both sides of a step are sorted, so the step is a split search and two merges
		 * (gdsp_merge_exchange) rather than a sort of the pair. */
		int inTmp = 0;
		for (int d = 0; d <= last; d++)
			{
			gdsp_layout* lay;
			gd_check (gdsp_layout_create (gd.ctx, &front[d], 1, &lay), _op->name);
			gd_check (gdsp_sort_genome (gd.ctx, lay, gd.sig, gd.tmp, gd.cells, &inTmp), _op->name);
			if (inTmp) gd_check (gdsp_d2d (gd.ctx, gd.sig + front[d].lo, gd.tmp + front[d].lo,
			                               (front[d].hi - front[d].lo) * sizeof (double)), _op->name);
			gdsp_layout_destroy (lay);
			}
		for (int c = 0; c <= K; c++)
			for (int d = c + 1; d <= last; d++)
				gd_check (gdsp_merge_exchange (gd.ctx, gd.sig, gd.tmp, front[c].lo, front[c].hi - front[c].lo,
				                               front[d].lo, front[d].hi - front[d].lo, NULL), _op->name);
		}
	free (front);
	}
	return;

no_values:
	fprintf (stderr, "[%s] percentile can't be computed;  no input values meet the criteria\n", _op->name);
	}
