/* gd_device.c -- device state of the host layer: one context, the packed genome
 * buffer (replacing the per-chromosome callocs of main, genodsp.c:865-878, and
 * the scratch-vector pool, genodsp.c:1904-2037) and interval-file loading. */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include "gd_device.h"

gdev gd;

void gd_check (int status, const char* who)
	{
	if (status == GDSP_OK) return;
	fprintf (stderr, "[%s] GPU failure: %s\n", who, gdsp_last_error ());
	exit (EXIT_FAILURE);
	}

/* Opening the device (CUDA initialisation, context, the two genome-sized buffers and their zero fill) takes
 * a large part of a second and needs nothing from the input: it runs on a helper thread while the main
 * thread parses the interval text.  gd_device_open() does the host-side bookkeeping and
 * starts the thread; gd_device_wait() joins it (idempotent) and must precede the first device call. */
#include <pthread.h>
static pthread_t openThread;
static int       openPending = 0;

static void* device_open_thread (void* arg)
	{
	(void) arg;
	const int n = gd.nchrom;
	gdsp_ctx* ctx;
	gd_check (gdsp_ctx_create (0, GDSP_STREAM_PRIVATE, &ctx), "genodsp");
	gd.ctx = ctx;
	gd_check (gdsp_layout_create (gd.ctx, gd.segs, n, &gd.genome), "genodsp");
	for (int i = 0; i < n; i++)
		gd_check (gdsp_layout_create (gd.ctx, &gd.segs[i], 1, &gd.single[i]), "genodsp");
	void* p;
	gd_check (gdsp_malloc (gd.ctx, gd.cells * sizeof (double), &p), "genodsp");  gd.sig = (double*) p;
	gd_check (gdsp_malloc (gd.ctx, gd.cells * sizeof (double), &p), "genodsp");  gd.tmp = (double*) p;
	gd_check (gdsp_fill (gd.ctx, gd.genome, gd.sig, 0.0), "genodsp");
	return NULL;
	}

void gd_device_open (void)
	{
	int n = 0;
	while (chromsSorted[n] != NULL) n++;
	memset (&gd, 0, sizeof (gd));
	gd.nchrom = n;
	u32* lens = (u32*) malloc (n * sizeof (u32));
	gd.segs   = (gdsp_seg*) malloc (n * sizeof (gdsp_seg));
	gd.single = (gdsp_layout**) malloc (n * sizeof (gdsp_layout*));
	for (int i = 0; i < n; i++)
		{
		lens[i] = chromsSorted[i]->length;
		if (lens[i] > gd.maxLength) gd.maxLength = lens[i];
		}
	u64 total;
	gd_check (gdsp_layout_pack (lens, n, gd.segs, &total), "genodsp");
	gd.cells = total;
	free (lens);
	if (getenv ("GENODSP_PARSE_ONLY") != NULL) return;      /* tokenizer self-check (tests): no device needed */
	/* the CLI computes on device 0 only: unless the user chose devices, hide the others before the CUDA runtime
	 * initialises, so that an 8-GPU node does not create eight primary-context candidates */
	setenv ("CUDA_VISIBLE_DEVICES", "0", 0);
	if (getenv ("GENODSP_SYNC_OPEN") != NULL || pthread_create (&openThread, NULL, device_open_thread, NULL) != 0)
		device_open_thread (NULL);
	else
		openPending = 1;
	if (!openPending) gd_device_wait ();
	}

void gd_device_wait (void)
	{
	if (openPending) { pthread_join (openThread, NULL);  openPending = 0; }
	if (gd.sig != NULL && chromsSorted[0] != NULL && chromsSorted[0]->valVector == NULL)
		for (int i = 0; i < gd.nchrom; i++) chromsSorted[i]->valVector = gd.sig + gd.segs[i].lo;
	}

void gd_device_close (void)
	{
	gd_device_wait ();
	if (gd.ctx == NULL) return;
	gdsp_sync (gd.ctx);
	for (int i = 0; i < gd.nchrom; i++) gdsp_layout_destroy (gd.single[i]);
	gdsp_layout_destroy (gd.genome);
	gdsp_free (gd.ctx, gd.sig);  gdsp_free (gd.ctx, gd.tmp);
	if (gd.work) gdsp_free (gd.ctx, gd.work);
	gdsp_ctx_destroy (gd.ctx);
	free (gd.segs);  free (gd.single);
	memset (&gd, 0, sizeof (gd));
	}

void gd_swap (void)
	{
	double* t = gd.sig;  gd.sig = gd.tmp;  gd.tmp = t;
	for (int i = 0; i < gd.nchrom; i++) chromsSorted[i]->valVector = gd.sig + gd.segs[i].lo;
	}

void* gd_work (size_t bytes)
	{
	if (bytes > gd.work_bytes)
		{
		if (gd.work) gd_check (gdsp_free (gd.ctx, gd.work), "genodsp");
		gd.work = NULL;  gd.work_bytes = 0;
		void* p;
		gd_check (gdsp_malloc (gd.ctx, bytes, &p), "genodsp");
		gd.work = p;  gd.work_bytes = bytes;
		}
	return gd.work;
	}

int gd_sorted_index (spec* c)
	{
	for (int i = 0; i < gd.nchrom; i++) if (chromsSorted[i] == c) return i;
	return -1;
	}

const gdsp_layout* gd_layout_for (valtype* v, int* sortedIx)
	{
	if (v == NULL) { if (sortedIx) *sortedIx = -1;  return gd.genome; }
	for (int i = 0; i < gd.nchrom; i++)
		if (chromsSorted[i]->valVector == v) { if (sortedIx) *sortedIx = i;  return gd.single[i]; }
	fprintf (stderr, "internal error: apply called with a vector that is not a chromosome\n");
	exit (EXIT_FAILURE);
	}

void gd_commit_tmp (valtype* v, int sortedIx)
	{
	if (v == NULL) { gd_swap ();  return; }
	const gdsp_seg* s = &gd.segs[sortedIx];
	gd_check (gdsp_d2d (gd.ctx, gd.sig + s->lo, gd.tmp + s->lo, (s->hi - s->lo) * sizeof (double)), "genodsp");
	}

/* ---- interval lists -------------------------------------------------------- */

void ivlist_init (ivlist* l) { memset (l, 0, sizeof (*l)); }

void ivlist_free (ivlist* l)
	{
	free (l->seg);  free (l->start);  free (l->end);  free (l->val);
	memset (l, 0, sizeof (*l));
	}

void ivlist_push (ivlist* l, u32 seg, u32 start, u32 end, double val)
	{
	if (l->n == l->cap)
		{
		u64 nc = l->cap ? 2 * l->cap : (1u << 16);
		l->seg   = (u32*) realloc (l->seg,   nc * sizeof (u32));
		l->start = (u32*) realloc (l->start, nc * sizeof (u32));
		l->end   = (u32*) realloc (l->end,   nc * sizeof (u32));
		l->val   = (double*) realloc (l->val, nc * sizeof (double));
		if (!l->seg || !l->start || !l->end || !l->val)
			{ fprintf (stderr, "out of memory holding %llu intervals\n", (unsigned long long) nc);  exit (EXIT_FAILURE); }
		l->cap = nc;
		}
	l->seg[l->n] = seg;  l->start[l->n] = start;  l->end[l->n] = end;  l->val[l->n] = val;
	l->n++;
	}

static ivlist* sortTarget;
static int iv_cmp (const void* a, const void* b)
	{
	u64 i = *(const u64*) a, j = *(const u64*) b;
	if (sortTarget->seg[i]   != sortTarget->seg[j])   return (sortTarget->seg[i]   < sortTarget->seg[j])   ? -1 : 1;
	if (sortTarget->start[i] != sortTarget->start[j]) return (sortTarget->start[i] < sortTarget->start[j]) ? -1 : 1;
	return (i < j) ? -1 : (i > j);            /* stable */
	}

void ivlist_sort (ivlist* l)
	{
	if (l->n < 2) return;
	int sorted = true;
	for (u64 k = 1; k < l->n && sorted; k++)
		if (l->seg[k] < l->seg[k-1] || (l->seg[k] == l->seg[k-1] && l->start[k] < l->start[k-1])) sorted = false;
	if (sorted) return;
	u64* perm = (u64*) malloc (l->n * sizeof (u64));
	for (u64 k = 0; k < l->n; k++) perm[k] = k;
	sortTarget = l;
	qsort (perm, l->n, sizeof (u64), iv_cmp);
	u32* seg = (u32*) malloc (l->n * sizeof (u32));  u32* st = (u32*) malloc (l->n * sizeof (u32));
	u32* en  = (u32*) malloc (l->n * sizeof (u32));  double* va = (double*) malloc (l->n * sizeof (double));
	for (u64 k = 0; k < l->n; k++)
		{ seg[k] = l->seg[perm[k]];  st[k] = l->start[perm[k]];  en[k] = l->end[perm[k]];  va[k] = l->val[perm[k]]; }
	free (l->seg);  free (l->start);  free (l->end);  free (l->val);  free (perm);
	l->seg = seg;  l->start = st;  l->end = en;  l->val = va;  l->cap = l->n;
	}

void ivlist_union (ivlist* l)
	{
	ivlist_sort (l);
	u64 o = 0;
	for (u64 k = 0; k < l->n; k++)
		{
		if (l->start[k] >= l->end[k]) continue;
		if (o > 0 && l->seg[o-1] == l->seg[k] && l->start[k] <= l->end[o-1])
			{ if (l->end[k] > l->end[o-1]) l->end[o-1] = l->end[k];  continue; }
		l->seg[o] = l->seg[k];  l->start[o] = l->start[k];  l->end[o] = l->end[k];  l->val[o] = 1.0;
		o++;
		}
	l->n = o;
	}

void ivlist_read_file (ivlist* l, const char* opName, const char* filename,
                       int valCol, int originOne, int skipZeroVal, int requireSorted)
	{
	FILE* f = fopen (filename, "rt");
	if (f == NULL)
		{ fprintf (stderr, "[%s] can't open \"%s\" for reading\n", opName, filename);  exit (EXIT_FAILURE); }

	char   line[1001], prevChrom[1001];
	char*  chrom;
	u32    start, end, prevEnd = 0;
	double val;
	spec*  cs = NULL;
	int    segIx = -1;
	u32    o = originOne ? 1 : 0;

	for (int i = 0; i < gd.nchrom; i++) chromsSorted[i]->flag = false;
	prevChrom[0] = 0;
	while (read_interval (f, line, sizeof (line), valCol, &chrom, &start, &end, &val))
		{
		if (skipZeroVal && val == 0.0) continue;
		if (strcmp (chrom, prevChrom) != 0)
			{
			cs = find_chromosome_spec (chrom);
			segIx = (cs != NULL) ? gd_sorted_index (cs) : -1;
			if (cs != NULL && requireSorted)
				{
				prevEnd = 0;
				if (cs->flag)
					{
					fprintf (stderr, "[%s] in \"%s\", not all intervals on %s are together (%d..%d begins new group)\n",
					         opName, filename, chrom, start, end);
					exit (EXIT_FAILURE);
					}
				}
			safe_strncpy (prevChrom, chrom, sizeof (prevChrom) - 1);
			}
		if (cs == NULL) continue;
		if (!cs->flag)
			{
			if (trackOperations) fprintf (stderr, "%s(%s)\n", opName, chrom);
			cs->flag = true;
			}
		start -= o;
		u32 a = start, b = end;
		if (cs->start == 0)
			{
			if (end > cs->length)
				{
				fprintf (stderr, "[%s] in \"%s\", %s %d %d is beyond the end of the chromosome (L=%d)\n",
				         opName, filename, chrom, start, end, cs->length);
				exit (EXIT_FAILURE);
				}
			}
		else
			{
			if (end <= cs->start) continue;
			b = end - cs->start;
			a = (start <= cs->start) ? 0 : start - cs->start;
			if (a >= cs->length) continue;
			if (b >= cs->length) b = cs->length;
			}
		if (requireSorted)
			{
			if (a < prevEnd)
				{
				fprintf (stderr, "[%s] in \"%s\", intervals on %s are not sorted (%d..%d after %d)\n",
				         opName, filename, chrom, start, end, cs->start + prevEnd);
				exit (EXIT_FAILURE);
				}
			prevEnd = b;
			}
		ivlist_push (l, (u32) segIx, a, b, val);
		}
	fclose (f);
	/* the sorted-file operators then treat every chromosome the file never mentioned
	 * (multiply.c:343-355, :737-749, logical.c:884-896, mask.c:622-634) */
	if (requireSorted && trackOperations)
		for (int i = 0; i < gd.nchrom; i++)
			if (!chromsSorted[i]->flag) fprintf (stderr, "%s(%s,absent)\n", opName, chromsSorted[i]->chrom);
	}

/* ---- exact application of valued intervals --------------------------------------------------
 * The reference adds every interval's value to every cell it covers, one interval after the other
 * in file order (genodsp.c:1325-1329, add.c:280-281).  For integer and dyadic values any order gives
 * the same bits and the difference-array kernels are used; for other real values (bedGraph-like
 * decimals, abutting or overlapping intervals, NaN / infinity) a difference array is NOT exact --
 * v + (-a + b) is not b -- so the intervals are cut into elementary pieces, every piece keeps the
 * values of the intervals covering it IN FILE ORDER, and layer k (the k-th value of every piece) is
 * applied by one pointwise launch: each cell receives exactly the reference's sequence of additions.
 * Disjoint input (the common case) is a single launch.  Returns false (nothing done) when some piece
 * is covered by more than GD_EXACT_MAX_DEPTH intervals or the piece lists would not fit in memory;
 * the caller then falls back to the difference array (1e-12 relative). */
#define GD_EXACT_MAX_DEPTH 4096

static int ex_by_pos (const void* a, const void* b)
	{
	u32 x = *(const u32*) a, y = *(const u32*) b;
	return (x > y) - (x < y);
	}

typedef struct exseg { u32* bp;  u64 np;  u32* cnt;  u64* off;  double* vals;  u32 maxDepth; } exseg;

/* phase 1 (host only): elementary pieces of every chromosome and their value lists in file order;
 * false when some piece is too deep or the lists would not fit (ex is still to be released) */
static int ex_build (ivlist* l, exseg* ex)
	{
	/* per chromosome (segment): intervals grouped, file order kept */
	u64* segCount = (u64*) calloc ((size_t) gd.nchrom + 1, sizeof (u64));
	for (u64 k = 0; k < l->n; k++) segCount[l->seg[k] + 1]++;
	for (int s = 0; s < gd.nchrom; s++) segCount[s + 1] += segCount[s];
	u64* order = (u64*) malloc (l->n * sizeof (u64));
	{
	u64* cur = (u64*) malloc ((size_t) gd.nchrom * sizeof (u64));
	for (int s = 0; s < gd.nchrom; s++) cur[s] = segCount[s];
	for (u64 k = 0; k < l->n; k++) order[cur[l->seg[k]]++] = k;           /* stable: file order inside a segment */
	free (cur);
	}

	int ok = true;
	/* piece-value slots we are willing to hold: an interval that spans p elementary pieces takes p slots */
	u64 budget = 64 * l->n + (1u << 20);
	if (budget > 1000000000ull) budget = 1000000000ull;
	for (int s = 0; s < gd.nchrom && ok; s++)
		{
		u64 a = segCount[s], b = segCount[s + 1], m = b - a;
		if (m == 0) continue;
		u32* bp = (u32*) malloc (2 * m * sizeof (u32));
		u64 nb = 0;
		for (u64 q = a; q < b; q++) { u64 k = order[q];  if (l->start[k] < l->end[k]) { bp[nb++] = l->start[k];  bp[nb++] = l->end[k]; } }
		if (nb == 0) { free (bp);  continue; }
		qsort (bp, nb, sizeof (u32), ex_by_pos);
		u64 u = 0;
		for (u64 k = 0; k < nb; k++) if (u == 0 || bp[k] != bp[u-1]) bp[u++] = bp[k];
		nb = u;
		u64 np = nb - 1;                                                    /* pieces [bp[p], bp[p+1]) */
		u32* cnt = (u32*) calloc (np + 1, sizeof (u32));
		/* depth of every piece through a difference array over piece indices (integers: exact) */
		int64_t* d = (int64_t*) calloc (np + 2, sizeof (int64_t));
		for (u64 q = a; q < b; q++)
			{
			u64 k = order[q];
			if (l->start[k] >= l->end[k]) continue;
			u32* ps = (u32*) bsearch (&l->start[k], bp, nb, sizeof (u32), ex_by_pos);
			u32* pe = (u32*) bsearch (&l->end[k],   bp, nb, sizeof (u32), ex_by_pos);
			d[ps - bp] += 1;  d[pe - bp] -= 1;
			}
		int64_t run = 0;  u64 total = 0;  u32 maxDepth = 0;
		for (u64 p = 0; p < np; p++)
			{
			run += d[p];  cnt[p] = (u32) run;  total += (u64) run;
			if (run > GD_EXACT_MAX_DEPTH) ok = false;
			if ((u32) run > maxDepth) maxDepth = (u32) run;
			}
		free (d);
		if (total > budget) ok = false; else budget -= total;
		ex[s].bp = bp;  ex[s].np = np;  ex[s].cnt = cnt;  ex[s].maxDepth = maxDepth;
		if (!ok) break;
		u64* off = (u64*) malloc ((np + 1) * sizeof (u64));
		off[0] = 0;
		for (u64 p = 0; p < np; p++) off[p + 1] = off[p] + cnt[p];
		double* vals = (double*) malloc ((off[np] + 1) * sizeof (double));
		u32* fill = (u32*) calloc (np + 1, sizeof (u32));
		for (u64 q = a; q < b; q++)                                           /* file order */
			{
			u64 k = order[q];
			if (l->start[k] >= l->end[k]) continue;
			u64 p0 = (u64) ((u32*) bsearch (&l->start[k], bp, nb, sizeof (u32), ex_by_pos) - bp);
			u64 p1 = (u64) ((u32*) bsearch (&l->end[k],   bp, nb, sizeof (u32), ex_by_pos) - bp);
			for (u64 p = p0; p < p1; p++) vals[off[p] + fill[p]++] = l->val[k];
			}
		free (fill);
		ex[s].off = off;  ex[s].vals = vals;
		}
	free (segCount);  free (order);
	return ok;
	}

static void ex_release (exseg* ex)
	{
	for (int s = 0; s < gd.nchrom; s++) { free (ex[s].bp);  free (ex[s].cnt);  free (ex[s].off);  free (ex[s].vals); }
	free (ex);
	}

/* `input --overlap=min|max` when the plain extreme is not what the reference computes.  Its rule per
 * position, interval after interval in file order (genodsp.c:1307-1322), is
 *     v == missingVal ? v = val : (val < v ? v = val : v)          (val > v for max)
 * which is the smallest (largest) covering value UNLESS the running value comes back to missingVal --
 * an interval whose value equals missingVal, e.g. a zero-valued bedGraph row with the default
 * --missing=0 -- after which the next interval overwrites it; a NaN value sticks once it is assigned.
 * The fold of every elementary piece's ordered value list is that rule itself.  Pieces come out in
 * layout order; false (nothing appended) when the lists would not fit. */
int gd_fold_intervals_minmax (ivlist* l, int wantMax, valtype missing, ivlist* outp)
	{
	if (l->n == 0) return true;
	exseg* ex = (exseg*) calloc ((size_t) gd.nchrom, sizeof (exseg));
	int ok = ex_build (l, ex);
	if (ok)
		for (int s = 0; s < gd.nchrom; s++)
			for (u64 p = 0; p < ex[s].np; p++)
				{
				if (ex[s].cnt[p] == 0) continue;
				const double* vals = ex[s].vals + ex[s].off[p];
				double v = missing;
				for (u32 k = 0; k < ex[s].cnt[p]; k++)
					{
					double val = vals[k];
					if (v == missing) v = val;
					else if (wantMax ? (val > v) : (val < v)) v = val;
					}
				ivlist_push (outp, (u32) s, ex[s].bp[p], ex[s].bp[p + 1], v);
				}
	ex_release (ex);
	return ok;
	}

int gd_apply_intervals_exact (ivlist* l, int mode, valtype missing, const char* opName)
	{
	if (l->n == 0)
		{
		if (mode == GD_EXACT_CLEAR) gd_check (gdsp_fill (gd.ctx, gd.genome, gd.sig, missing), opName);
		return true;
		}
	exseg* ex = (exseg*) calloc ((size_t) gd.nchrom, sizeof (exseg));
	int ok = ex_build (l, ex);
	if (getenv ("GENODSP_TIMING") != NULL)
		fprintf (stderr, "[timing] exact interval application: %s\n", ok ? "yes" : "no (difference array)");

	/* phase 2 (only if every chromosome fits): layer k of chromosome s = one pointwise launch over s */
	if (ok)
		{
		if (mode == GD_EXACT_CLEAR) gd_check (gdsp_fill (gd.ctx, gd.genome, gd.sig, missing), opName);
		ivlist layer;
		ivlist_init (&layer);
		for (int s = 0; s < gd.nchrom; s++)
			for (u32 k = 0; k < ex[s].maxDepth; k++)
				{
				layer.n = 0;
				for (u64 p = 0; p < ex[s].np; p++)
					if (ex[s].cnt[p] > k) ivlist_push (&layer, 0, ex[s].bp[p], ex[s].bp[p + 1], ex[s].vals[ex[s].off[p] + k]);
				gdsp_ivl_table* t;
				gd_check (gdsp_ivl_table_create (gd.ctx, gd.single[s], layer.seg, layer.start, layer.end, layer.val, layer.n, &t), opName);
				gdsp_pw_op p;  memset (&p, 0, sizeof (p));
				p.code = (mode == GD_EXACT_CLEAR) ? GDSP_PW_IVL_ACCUM_CLEAR : (mode == GD_EXACT_SUB) ? GDSP_PW_IVL_SUB : GDSP_PW_IVL_ADD;
				p.a = missing;  p.table = t;
				gd_check (gdsp_pointwise (gd.ctx, gd.single[s], gd.sig, gd.sig, &p, 1), opName);
				gdsp_ivl_table_destroy (t);
				}
		ivlist_free (&layer);
		}
	ex_release (ex);
	return ok;
	}
