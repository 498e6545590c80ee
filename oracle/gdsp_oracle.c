/* gdsp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see gdsp_oracle.h).
 *
 * CPU restatement of genodsp 0.0.10's per-base algorithms.  Selection /
 * comparison operators are written from their closed forms;
 * operators whose floating-point result depends on evaluation order follow the
 * reference's order exactly (no FMA: build with -ffp-contract=off, as the
 * reference's x86-64 baseline build has none).
 *
 * Every function cites the reference file:line it restates.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include "gdsp_oracle.h"

typedef uint32_t u32;
typedef uint64_t u64;

/* read_intervals accumulate loops, genodsp.c:1307-1330.
 * overlapOp 0 sum, 1 min, 2 max (genodsp_interface.h:149-153). */
void gdo_accumulate (double* v, u32 n, const u32* s, const u32* e,
                     const double* val, u64 m, int overlapOp, int clear,
                     double missing)
	{
	for (u64 k = 0; k < m; k++)
		{
		double x = (val != NULL) ? val[k] : 1.0;
		u32 hi = e[k] < n ? e[k] : n;
		for (u32 i = s[k]; i < hi; i++)
			{
			if (clear && v[i] == missing) { v[i] = x; continue; }
			if      (overlapOp == 1) { if (x < v[i]) v[i] = x; }
			else if (overlapOp == 2) { if (x > v[i]) v[i] = x; }
			else                       v[i] += x;
			}
		}
	}

/* op_window_sum_apply, sum.c:211-252: blocks [kW,(k+1)W) counted from the
 * vector start; the block total (left-to-right, seeded by the first element)
 * over denom goes to the block's first slot, `zero` elsewhere. */
void gdo_block_sum (double* v, u32 n, u32 W, double denom, int actualDenom, double zero)
	{
	for (u64 b = 0; b < n; b += W)
		{
		u32 lo = (u32) b;
		u32 hi = (b + W < n) ? (u32) (b + W) : n;
		double t = v[lo];
		for (u32 i = lo + 1; i < hi; i++) t += v[i];
		v[lo] = actualDenom ? t / (double) (hi - lo) : t / denom;
		for (u32 i = lo + 1; i < hi; i++) v[i] = zero;
		}
	}

/* op_sliding_sum_apply, sum.c:420-463: running sum, add the entering element
 * first and then subtract the leaving one, output centred with
 * h=(W-1)/2, finally divided by denom. */
void gdo_sliding_sum (double* v, u32 n, u32 W, double denom)
	{
	u32 h = (W - 1) / 2;
	double* out = (double*) malloc ((size_t) n * sizeof(double));
	double run = 0.0;
	for (u64 t = 0; t < (u64) n + h; t++)
		{
		if (t < n)  run += v[t];
		if (t >= W) run -= v[t - W];
		if (t >= h) out[t - h] = run;
		}
	for (u32 i = 0; i < n; i++) v[i] = out[i] / denom;
	free (out);
	}

/* Hann taps, sum.c:634-645: built symmetrically from both ends, then each
 * divided by the left-to-right total. */
void gdo_hann_taps (double* w, u32 W)
	{
	u32 h = (W - 1) / 2;
	for (u32 k = 0; k <= h; k++)
		{
		double x = (k + 1) / (double) (W + 1);
		w[k] = w[W - 1 - k] = (1 - cos (2 * M_PI * x)) / 2;
		}
	double tot = 0.0;
	for (u32 k = 0; k < W; k++) tot += w[k];
	for (u32 k = 0; k < W; k++) w[k] /= tot;
	}

/* op_smooth_apply, sum.c:616-676: direct FIR, taps applied in ascending tap
 * order, truncated (not renormalised) at both ends; product and sum rounded
 * separately. */
void gdo_smooth (double* v, u32 n, u32 W)
	{
	u32 h = (W - 1) / 2;
	double* w   = (double*) malloc ((size_t) W * sizeof(double));
	double* out = (double*) malloc ((size_t) n * sizeof(double));
	gdo_hann_taps (w, W);
	for (u32 i = 0; i < n; i++)
		{
		u32 k0 = (i < h) ? h - i : 0;
		u32 k1 = ((u64) i + W > (u64) n + h) ? n - 1 + h - i : W - 1;
		double acc = 0.0;
		for (u32 k = k0; k <= k1; k++)
			{
			double p = w[k] * v[i - h + k];
			acc += p;
			}
		out[i] = acc;
		}
	memcpy (v, out, (size_t) n * sizeof(double));
	free (out);  free (w);
	}

/* op_cumulative_sum_apply, sum.c:776-792 */
/* percentile --preserve (percentile.c:534-535, :717-725): the vectors are written as text by
 * write_all_chromosomes (genodsp.c:1754-1775: report_intervals with precision 10, runs of zero not
 * shown) and read back by read_all_chromosomes (genodsp.c:1717-1742: read_intervals, value column 4,
 * clear to 0.0).  Per cell: a zero comes back as +0.0; anything else as what read_interval makes of
 * its "%.10f" text, i.e. string_to_double (utilities.c:334-370): "inf" -> DBL_MAX, else sscanf %lf. */
void gdo_text_roundtrip10 (double* v, u32 n)
	{
	char buf[512];
	for (u32 i = 0; i < n; i++)
		{
		if (v[i] == 0.0) { v[i] = 0.0;  continue; }
		snprintf (buf, sizeof (buf), "%.10f", v[i]);
		if      (strcmp (buf, "inf")  == 0) v[i] =  DBL_MAX;
		else if (strcmp (buf, "-inf") == 0) v[i] = -DBL_MAX;
		else { double x = 0;  char extra;  if (sscanf (buf, "%lf%c", &x, &extra) == 1) v[i] = x; }
		}
	}

/* minover / maxover, minmax.c:322-348 and :725-751.  One pass over each interval keeps the best
 * value, its index and that index's inset (distance to the nearer interval end); an equal value
 * replaces the incumbent only with a strictly larger inset.  Then the interval is filled except at
 * the winner; gaps before, between and after the intervals are filled too (minmax.c:312-320,
 * :354-363). */
void gdo_over_intervals (double* v, u32 n, const u32* s, const u32* e, u32 m, int want_max, double fill)
	{
	u32 prevEnd = 0;
	for (u32 k = 0; k < m; k++)
		{
		for (u32 ix = prevEnd; ix < s[k]; ix++) v[ix] = fill;
		double best = v[s[k]];
		u32 bestIx = s[k], bestInset = 0;
		for (u32 ix = s[k] + 1; ix < e[k]; ix++)
			{
			int worse  = want_max ? (v[ix] < best) : (v[ix] > best);
			int better = want_max ? (v[ix] > best) : (v[ix] < best);
			if (worse) continue;
			u32 inset = (ix - s[k] < e[k] - ix) ? ix - s[k] : e[k] - ix;
			if (better) { best = v[ix];  bestIx = ix;  bestInset = inset; }
			else if (inset > bestInset) { bestIx = ix;  bestInset = inset; }
			}
		for (u32 ix = s[k]; ix < e[k]; ix++) if (ix != bestIx) v[ix] = fill;
		prevEnd = e[k];
		}
	for (u32 ix = prevEnd; ix < n; ix++) v[ix] = fill;
	}

/* minwith / maxwith, minmax.c:1979-1982 and :2265-2268 */
void gdo_with_intervals (double* v, u32 n, const u32* s, const u32* e, const double* val, u32 m, int want_max)
	{
	(void) n;
	for (u32 k = 0; k < m; k++)
		for (u32 ix = s[k]; ix < e[k]; ix++)
			{
			if (want_max) { if (val[k] > v[ix]) v[ix] = val[k]; }
			else          { if (val[k] < v[ix]) v[ix] = val[k]; }
			}
	}

/* map, map.c:263-357, for breakpoints with distinct inputs (then the reference's piece cache does
 * not influence the result): clamp at the ends, exact hits return the listed output, otherwise
 * outLo + (x - inLo) * outDiff / inDiff in that evaluation order (map.c:340). */
void gdo_map (double* v, u32 n, const double* in, const double* out, u32 len)
	{
	for (u32 i = 0; i < n; i++)
		{
		double x = v[i];
		if (x <= in[0])       { v[i] = out[0];        continue; }
		if (x >= in[len - 1]) { v[i] = out[len - 1];  continue; }
		u32 lo = 0, hi = len - 1;
		while (lo + 1 < hi)
			{
			u32 mid = (lo + hi) / 2;
			if (x < in[mid]) hi = mid; else lo = mid;
			}
		double inDiff = in[lo + 1] - in[lo], outDiff = out[lo + 1] - out[lo];
		if      (x == in[lo])     v[i] = out[lo];
		else if (x == in[lo + 1]) v[i] = out[lo + 1];
		else                      v[i] = out[lo] + (x - in[lo]) * outDiff / inDiff;
		}
	}

void gdo_cumulative (double* v, u32 n)
	{
	double run = 0.0;
	for (u32 i = 0; i < n; i++) { run += v[i];  v[i] = run; }
	}

/* op_local_maxima_apply minmax.c:1183-1227 / op_local_minima_apply :981-1022:
 * keep v[i] unless some OTHER element of [i-h,i+h] (clipped to the vector) is
 * strictly greater (smaller); otherwise write `fill`. */
void gdo_local_extrema (double* v, u32 n, u32 N, int wantMax, double fill)
	{
	u32 h = (N - 1) / 2;
	double* out = (double*) malloc ((size_t) n * sizeof(double));
	for (u32 i = 0; i < n; i++)
		{
		u32 lo = (i < h) ? 0 : i - h;
		u32 hi = ((u64) i + h >= n) ? n - 1 : i + h;
		int beaten = 0;
		for (u32 j = lo; j <= hi && !beaten; j++)
			{
			if (j == i) continue;
			if (wantMax ? (v[j] > v[i]) : (v[j] < v[i])) beaten = 1;
			}
		out[i] = beaten ? fill : v[i];
		}
	memcpy (v, out, (size_t) n * sizeof(double));
	free (out);
	}

/* op_best_local_max_apply minmax.c:1616-1721 / ..._min_apply :1369-1474:
 * extremum of [i-l, i+r] clipped, l=(W-1)/2, r=(W-1)-l.  (Closed form; the
 * reference's incremental search returns the same value for NaN-free data.) */
void gdo_best_extrema (double* v, u32 n, u32 W, int wantMax)
	{
	u32 l = (W - 1) / 2, r = (W - 1) - l;
	double* out = (double*) malloc ((size_t) n * sizeof(double));
	for (u32 i = 0; i < n; i++)
		{
		u32 lo = (i < l) ? 0 : i - l;
		u32 hi = ((u64) i + r >= n) ? n - 1 : i + r;
		double b = v[lo];
		for (u32 j = lo + 1; j <= hi; j++)
			if (wantMax ? (v[j] > b) : (v[j] < b)) b = v[j];
		out[i] = b;
		}
	memcpy (v, out, (size_t) n * sizeof(double));
	free (out);
	}

/* op_close_apply, morphology.c:231-319.  Set = !(v<=T).  A gap [s,e) is filled
 * iff s!=0, e!=n and e-s <= L (L compared as a double). */
void gdo_close (double* v, u32 n, double L, double T, double one, double zero)
	{
	u32 i = 0;
	while (i < n)
		{
		if (!(v[i] <= T)) { v[i++] = one; continue; }
		u32 s = i;
		while (i < n && v[i] <= T) i++;
		int fill = (s != 0) && (i != n) && !((double) (i - s) > L);
		for (u32 j = s; j < i; j++) v[j] = fill ? one : zero;
		}
	}

/* op_open_apply, morphology.c:529-605.  Set = (v>T).  A run [s,e) survives iff
 * e-s > L. */
void gdo_open (double* v, u32 n, double L, double T, double one, double zero)
	{
	u32 i = 0;
	while (i < n)
		{
		if (!(v[i] > T)) { v[i++] = zero; continue; }
		u32 s = i;
		while (i < n && v[i] > T) i++;
		int keep = ((double) (i - s) > L);
		for (u32 j = s; j < i; j++) v[j] = keep ? one : zero;
		}
	}

/* op_dilate_apply, morphology.c:882-1072.  out[i]=one iff some set element j
 * has i-right <= j <= i+left (element 0 tests v>T, the others !(v<=T),
 * morphology.c:932-934,940). */
void gdo_dilate (double* v, u32 n, u32 left, u32 right, double T, double one, double zero)
	{
	unsigned char* in = (unsigned char*) malloc (n);
	for (u32 i = 0; i < n; i++) in[i] = (i == 0) ? (v[i] > T) : !(v[i] <= T);
	/* nearest set element at or before i / at or after i */
	int64_t last = -1;
	int64_t* prev = (int64_t*) malloc ((size_t) n * sizeof(int64_t));
	for (u32 i = 0; i < n; i++) { if (in[i]) last = i;  prev[i] = last; }
	int64_t next = -1;
	for (int64_t i = (int64_t) n - 1; i >= 0; i--)
		{
		if (in[i]) next = i;
		int hit = (prev[i] >= 0 && i - prev[i] <= (int64_t) right)
		       || (next    >= 0 && next - i <= (int64_t) left);
		v[i] = hit ? one : zero;
		}
	free (prev);  free (in);
	}

/* op_erode_apply, morphology.c:1331-1454.  Each run [s,e) of (v>T) keeps
 * [s+right, e-left); the reference's u32 underflow case (e<left, a crash
 * there) is defined as "run removed". */
void gdo_erode (double* v, u32 n, u32 left, u32 right, double T, double one, double zero)
	{
	u32 i = 0;
	while (i < n)
		{
		if (!(v[i] > T)) { v[i++] = zero; continue; }
		u32 s = i;
		while (i < n && v[i] > T) i++;
		u64 a = (u64) s + right;
		int64_t b = (int64_t) i - (int64_t) left;
		for (u32 j = s; j < i; j++)
			v[j] = ((int64_t) a < b && j >= a && (int64_t) j < b) ? one : zero;
		}
	}

/* op_binarize_apply, logical.c:216-268 */
void gdo_binarize (double* v, u32 n, double T, int tiesAbove, double one, double zero)
	{
	for (u32 i = 0; i < n; i++)
		v[i] = (tiesAbove ? (v[i] >= T) : (v[i] > T)) ? one : zero;
	}

/* op_add_constant_apply, add.c:726-741 */
void gdo_addconst (double* v, u32 n, double c)
	{
	if (c == 0.0) return;
	for (u32 i = 0; i < n; i++) v[i] += c;
	}

/* op_absolute_value_apply, add.c:1038-1049 */
void gdo_abs (double* v, u32 n)
	{ for (u32 i = 0; i < n; i++) if (v[i] < 0) v[i] = -v[i]; }

/* op_clip_apply, mask.c:850-924 */
void gdo_clip (double* v, u32 n, int haveMin, double mn, int haveMax, double mx)
	{
	for (u32 i = 0; i < n; i++)
		{
		if      (haveMin && v[i] < mn) v[i] = mn;
		else if (haveMax && v[i] > mx) v[i] = mx;
		}
	}

/* op_erase_apply, mask.c:1147-1243 */
void gdo_erase (double* v, u32 n, int haveMin, double mn, int haveMax, double mx,
                int keepInside, double zero)
	{
	for (u32 i = 0; i < n; i++)
		{
		int kill;
		if (keepInside) kill = (haveMin && v[i] < mn) || (haveMax && v[i] > mx);
		else            kill = (!haveMin || v[i] >= mn) && (!haveMax || v[i] <= mx);
		if (kill) v[i] = zero;
		}
	}

/* op_invert_apply, add.c:890-939 (application half): v = 2*mid - v */
void gdo_invert (double* v, u32 n, double mid)
	{
	double twice = 2 * mid;
	for (u32 i = 0; i < n; i++) v[i] = twice - v[i];
	}

/* running min/max used by invert (add.c:907-926) and percentile 0/100;
 * mn/mx are updated, caller seeds them */
void gdo_minmax (const double* v, u32 n, double* mn, double* mx)
	{
	for (u32 i = 0; i < n; i++)
		{
		if (v[i] < *mn) *mn = v[i];
		if (v[i] > *mx) *mx = v[i];
		}
	}

/* or/and first pass, logical.c:466-473, :768-775: non-zero becomes 1.0 */
void gdo_logical_prep (double* v, u32 n)
	{ for (u32 i = 0; i < n; i++) if (v[i] != 0.0) v[i] = 1.0; }

/* op_add_apply add.c:231-282 (sign=+1) / op_subtract_apply add.c:484-599
 * (sign=-1): intervals applied in file order, val==0 skipped */
void gdo_add_intervals (double* v, u32 n, const u32* s, const u32* e,
                        const double* val, u64 m, double sign)
	{
	for (u64 k = 0; k < m; k++)
		{
		if (val[k] == 0.0) continue;
		u32 hi = e[k] < n ? e[k] : n;
		if (sign > 0) for (u32 i = s[k]; i < hi; i++) v[i] += val[k];
		else          for (u32 i = s[k]; i < hi; i++) v[i] -= val[k];
		}
	}

/* op_mask_apply, mask.c:283-296 */
void gdo_mask_intervals (double* v, u32 n, const u32* s, const u32* e, u64 m, double maskVal)
	{
	for (u64 k = 0; k < m; k++)
		{
		u32 hi = e[k] < n ? e[k] : n;
		for (u32 i = s[k]; i < hi; i++) v[i] = maskVal;
		}
	}

/* op_or_apply, logical.c:439-560 (after gdo_logical_prep) */
void gdo_or_intervals (double* v, u32 n, const u32* s, const u32* e,
                       const double* val, u64 m)
	{
	for (u64 k = 0; k < m; k++)
		{
		if (val != NULL && val[k] == 0.0) continue;
		u32 hi = e[k] < n ? e[k] : n;
		for (u32 i = s[k]; i < hi; i++) v[i] = 1.0;
		}
	}

/* sorted non-overlapping interval operators on one chromosome:
 *   kind 0  multiply  multiply.c:193-393  inside v*=val, gaps 0.0
 *   kind 1  divide    multiply.c:586-787  inside v/=val, gaps +-aux (infinity)
 *   kind 2  masknot   mask.c:483-668      inside kept,   gaps aux (mask value)
 *   kind 3  and       logical.c:737-930   inside kept,   gaps 0.0
 * m==0 means the chromosome is absent from the file: all gap. */
void gdo_sorted_intervals (double* v, u32 n, const u32* s, const u32* e,
                           const double* val, u64 m, int kind, double aux)
	{
	u32 pos = 0;
	for (u64 k = 0; k <= m; k++)
		{
		u32 gs = pos, ge = (k < m) ? s[k] : n;
		for (u32 i = gs; i < ge; i++)
			{
			if      (kind == 0 || kind == 3) v[i] = 0.0;
			else if (kind == 1)              v[i] = (v[i] >= 0) ? aux : -aux;
			else                             v[i] = aux;
			}
		if (k == m) break;
		u32 hi = e[k] < n ? e[k] : n;
		if      (kind == 0) for (u32 i = s[k]; i < hi; i++) v[i] *= val[k];
		else if (kind == 1) for (u32 i = s[k]; i < hi; i++) v[i] /= val[k];
		pos = hi;
		}
	}

/* clump_search, clump.c:494-736.
 * d[i] = v[i]-T (T-v[i] for anticlump); all d<0 => everything zero (:545-565).
 * P = sequential prefix sums of d with P[-1]=0.  For each i let j(i) be the
 * earliest j in [-1,i] with P[j] <= P[i] (found by bisection over the running
 * prefix minimum, which is the reference's minSums/minScan bookkeeping
 * :600-625 in closed form); if i-j(i) >= minLength, [j(i)+1, i] is marked.
 * Each maximal marked run is then trimmed to its first..last element with
 * v>=T (v<=T for anticlump), :659-719. */
void gdo_clump (double* v, u32 n, double T, u32 minLength, int above, double one, double zero)
	{
	int allNeg = 1;
	for (u32 i = 0; i < n && allNeg; i++)
		{
		double d = above ? v[i] - T : T - v[i];
		if (d >= 0.0) allNeg = 0;
		}
	if (allNeg) { for (u32 i = 0; i < n; i++) v[i] = zero;  return; }

	/* M[k] = min(P[-1..k-1]) stored with a +1 shift so that M[0] is P[-1] */
	double* M = (double*) malloc (((size_t) n + 1) * sizeof(double));
	int32_t* mark = (int32_t*) calloc ((size_t) n + 1, sizeof(int32_t));
	double P = 0.0, mn = 0.0;
	M[0] = 0.0;
	for (u32 i = 0; i < n; i++)
		{
		double d = above ? v[i] - T : T - v[i];
		P += d;
		if (P < mn) mn = P;
		M[i + 1] = mn;
		/* earliest shifted index q in [0,i+1] with M[q] <= P (M is non-increasing) */
		u64 lo = 0, hi = (u64) i + 1;
		while (lo < hi)
			{
			u64 mid = (lo + hi) / 2;
			if (M[mid] <= P) hi = mid; else lo = mid + 1;
			}
		/* j = lo-1 ; interval [j+1, i] = [lo, i] has length i-j = i+1-lo */
		if ((u64) i + 1 - lo >= minLength) { mark[lo] += 1;  mark[i + 1] -= 1; }
		}
	free (M);

	int32_t depth = 0;
	u32 i = 0;
	unsigned char* in = (unsigned char*) malloc (n);
	for (u32 k = 0; k < n; k++) { depth += mark[k];  in[k] = depth > 0; }
	free (mark);
	while (i < n)
		{
		if (!in[i]) { v[i++] = zero; continue; }
		u32 s = i;
		while (i < n && in[i]) i++;
		int64_t first = -1, last = -1;
		for (u32 k = s; k < i; k++)
			if (above ? (v[k] >= T) : (v[k] <= T)) { if (first < 0) first = k;  last = k; }
		for (u32 k = s; k < i; k++)
			v[k] = (first >= 0 && (int64_t) k >= first && (int64_t) k <= last) ? one : zero;
		}
	free (in);
	}

/* percentile sample collection, percentile.c:547-580 (selection only; the
 * destructive permutation is modelled by the callers): every W-th element of
 * the chromosome, kept unless v<min or v>max */
u64 gdo_percentile_collect (const double* v, u32 n, u32 W, double mn, double mx, double* out)
	{
	u64 c = 0;
	for (u64 i = 0; i < n; i += W)
		{
		if (v[i] < mn) continue;
		if (v[i] > mx) continue;
		out[c++] = v[i];
		}
	return c;
	}

/* rank of percentile p (thousandths of a percent), percentile.c:588, :686:
 * (u32) ((u64) n * p / 100000.0) */
u64 gdo_percentile_rank (u64 numValues, u32 pMilli)
	{
	u32 nv = (u32) numValues;
	return (u32) (((u64) nv) * pMilli / (100.0 * 1000));
	}

static int dbl_up (const void* a, const void* b)
	{
	double x = *(const double*) a, y = *(const double*) b;
	return (x > y) - (x < y);
	}

/* ascending sort with the reference comparator valtype_ascending,
 * genodsp.c:2262-2270 */
void gdo_sort (double* v, u64 n) { qsort (v, n, sizeof(double), dbl_up); }

/* report_intervals state machine for one chromosome, genodsp.c:1589-1678,
 * restated as: maximal runs of raw-equal values (each base its own run when
 * !collapse); a run whose value ==0 is dropped unless showUncovered==1 (show).
 * NA lines (showUncovered==-1) are the gaps between emitted runs and are
 * derived by the caller.  Run value = value of the run's first element. */
u64 gdo_runs (const double* v, u32 n, int collapse, int showUncovered,
              u32* rs, u32* re, double* rv, u64 cap)
	{
	u64 r = 0;
	u32 i = 0;
	while (i < n)
		{
		u32 s = i;
		double x = v[i];
		i++;
		if (showUncovered != 1 && x == 0) continue;
		/* the state machine starts with val=+0.0 (:1590): a collapsed run of
		 * zeros that begins at base 0 reports that +0.0, not v[0] */
		if (s == 0 && collapse && x == 0) x = 0.0;
		if (collapse) while (i < n && v[i] == x && !(showUncovered != 1 && v[i] == 0)) i++;
		if (r < cap) { rs[r] = s;  re[r] = i;  rv[r] = x; }
		r++;
		}
	return r;
	}
