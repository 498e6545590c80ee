/* ref_shim.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * Wraps the UNMODIFIED reference translation unit genodsp.c (included by path
 * from /root/reference, never copied into this repo) so that tests can drive
 * the reference's own operators in-process on arrays:
 *   - its main() is renamed ref_main (-Dmain=ref_main on the command line),
 *   - the static helpers add_chromosome_spec (genodsp.c:1014),
 *     sort_chromosomes_by_length (genodsp.c:1113), init_scratch_vectors
 *     (genodsp.c:1895) and the dspTable[] registry (genodsp.c:117) become
 *     reachable because this file is the same translation unit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the resulting library.
 */
#include "genodsp.c"

#include <time.h>

/* drop every chromosome, operator and variable and start over */
void refshim_reset (void)
	{
	u32 ix;
	if (chromsSorted != NULL)
		{
		for (ix=0 ; chromsSorted[ix]!=NULL ; ix++)
			{
			spec* c = chromsSorted[ix];
			if (c->chrom     != NULL) free (c->chrom);
			if (c->valVector != NULL) free (c->valVector);
			free (c);
			}
		free (chromsSorted);
		chromsSorted = NULL;
		chromsOfInterest = NULL;
		free_scratch_vectors ();
		}
	else
		{
		spec* c, *n;
		for (c=chromsOfInterest ; c!=NULL ; c=n)
			{ n = c->next;  if (c->chrom != NULL) free (c->chrom);  free (c); }
		chromsOfInterest = NULL;
		}
	free_named_globals ();
	init_named_globals ();
	set_named_global ("valColumn",     (valtype) (4-1));
	set_named_global ("valPrecision",  (valtype) 0);
	set_named_global ("collapseRuns",  (valtype) true);
	set_named_global ("showUncovered", (valtype) uncovered_hide);
	set_named_global ("originOne",     (valtype) false);
	clipToLength = false;
	trackOperations = false;
	}

int refshim_add_chrom (char* name, u32 start, u32 length)
	{ return add_chromosome_spec (name, start, length); }

/* sort by length, set up scratch pool, allocate zeroed vectors
 * (what main does at genodsp.c:848-878) */
void refshim_finalize (void)
	{
	u32 maxLength = 0, ix;
	sort_chromosomes_by_length ();
	for (ix=0 ; chromsSorted[ix]!=NULL ; ix++)
		if (chromsSorted[ix]->length > maxLength) maxLength = chromsSorted[ix]->length;
	init_scratch_vectors (maxLength);
	for (ix=0 ; chromsSorted[ix]!=NULL ; ix++)
		chromsSorted[ix]->valVector = (valtype*) calloc (chromsSorted[ix]->length, sizeof(valtype));
	}

int refshim_num_chroms (void)
	{
	int n = 0;
	spec* c;
	for (c=chromsOfInterest ; c!=NULL ; c=c->next) n++;
	return n;
	}

/* name of the ix-th chromosome in sorted (descending length) order */
char* refshim_sorted_name (int ix) { return chromsSorted[ix]->chrom; }
u32   refshim_sorted_length (int ix) { return chromsSorted[ix]->length; }

valtype* refshim_vector (char* name)
	{
	spec* c = find_chromosome_spec (name);
	return (c == NULL)? NULL : c->valVector;
	}

void refshim_set_clip (int clip) { clipToLength = clip; }

/* parse "opName arg arg ..." through the reference's registry and run it the
 * way the executor loop (genodsp.c:900-936) would run a one-operator
 * pipeline; returns seconds spent inside the apply calls, <0 on lookup miss */
double refshim_apply (int argc, char** argv)
	{
	dspinfo* opInfo = NULL, *realOpInfo = NULL;
	u32      dspIx, ix;
	dspop*   op;
	struct timespec t0, t1;

	for (dspIx=0 ; dspIx<dspTableLen ; dspIx++)
		{
		if (dspTable[dspIx].funcShort != NULL) realOpInfo = &dspTable[dspIx];
		if (strcmp (argv[0],dspTable[dspIx].name) != 0) continue;
		opInfo = realOpInfo;
		break;
		}
	if (opInfo == NULL) return -1.0;

	op = (*opInfo->funcParse) (opInfo->name, argc-1, argv+1);
	op->name      = copy_string (opInfo->name);
	op->funcApply = opInfo->funcApply;
	op->funcFree  = opInfo->funcFree;
	op->next      = NULL;

	clock_gettime (CLOCK_MONOTONIC, &t0);
	if (op->atRandom)
		{
		u32 maxLength = 0;
		for (ix=0 ; chromsSorted[ix]!=NULL ; ix++)
			if (chromsSorted[ix]->length > maxLength) maxLength = chromsSorted[ix]->length;
		(*op->funcApply) (op, "*", maxLength, NULL);
		}
	else
		{
		for (ix=0 ; chromsSorted[ix]!=NULL ; ix++)
			(*op->funcApply) (op, chromsSorted[ix]->chrom,
			                  chromsSorted[ix]->length, chromsSorted[ix]->valVector);
		}
	clock_gettime (CLOCK_MONOTONIC, &t1);

	free (op->name);  op->name = NULL;
	(*op->funcFree) (op);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}

int refshim_read_intervals (char* filename, int valCol, int origin1,
                            int overlapOp, int clear, double missingVal)
	{
	FILE* f = fopen (filename, "rt");
	if (f == NULL) return 0;
	read_intervals (f, valCol, origin1, overlapOp, clear, missingVal);
	fclose (f);
	return 1;
	}

int refshim_report_intervals (char* filename, int precision, int noOutputValues,
                              int collapse, int showUncov, int origin1)
	{
	FILE* f = fopen (filename, "wt");
	if (f == NULL) return 0;
	report_intervals (f, precision, noOutputValues, collapse, showUncov, origin1);
	fclose (f);
	return 1;
	}

int refshim_get_global (char* name, double* v)
	{ return named_global_exists (name, v); }

void refshim_set_global (char* name, double v)
	{ set_named_global (name, v); }
