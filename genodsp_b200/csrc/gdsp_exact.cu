// gdsp_exact.cu -- opt-in "exact order" mode for the three operators whose reference loops are
// sequentially dependent rounded sums:
//     slidingsum     sum.c:438-455   sum += v[ix]; sum -= v[ix-W]   (one running sum per chromosome)
//     cumulativesum  sum.c:786-790   valSum += v[ix]
//     clump          clump.c:600     valSum += v[ix] - T             (the prefix sum the search compares)
// The default kernels associate these sums tile by tile: bit-identical to the reference whenever every
// partial sum is exact (integer / dyadic signals), within 1e-12 otherwise.  With the mode on
// (gdsp_ctx_set_exact_order, `--exact-order` in the CLI) the recurrence runs in the reference's order:
// one warp per chromosome, lanes stage 1024-cell chunks through shared memory (coalesced loads and
// stores, the next chunk already in flight in registers), lane 0 folds the chunk.  The result is the
// reference's bit for bit on EVERY input -- general reals, inf, NaN (which poison the running sum from
// there to the end of the chromosome, exactly as they do there) -- at the cost of a dependent FP64 add
// per cell: ~1-3 s for the longest human chromosome instead of milliseconds.  A correctness mode.
#include "gdsp_common.cuh"

#define EX_CHUNK 1024
#define EX_PER   (EX_CHUNK / 32)

extern "C" int gdsp_ctx_set_exact_order (gdsp_ctx* c, int on)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_ctx_set_exact_order: NULL context");
	c->exact_order = on ? 1 : 0;
	return GDSP_OK;
	}

extern "C" int gdsp_ctx_get_exact_order (const gdsp_ctx* c) { return (c != NULL && c->exact_order) ? 1 : 0; }

// ---- cumulativesum ----------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_cumsum_seq (const SegDev* __restrict__ segs, const double* __restrict__ in, double* __restrict__ out)
	{
	__shared__ double buf[EX_CHUNK];
	const SegDev sd = segs[blockIdx.x];
	const int lane = threadIdx.x;
	double nxt[EX_PER];
	#pragma unroll
	for (int k = 0; k < EX_PER; k++) { const uint64_t i = sd.lo + (uint64_t) k * 32 + lane;  nxt[k] = (i < sd.hi) ? in[i] : 0.0; }
	double run = 0.0;
	for (uint64_t c0 = sd.lo; c0 < sd.hi; c0 += EX_CHUNK)
		{
		const uint32_t n = (uint32_t) ((sd.hi - c0 < EX_CHUNK) ? (sd.hi - c0) : EX_CHUNK);
		#pragma unroll
		for (int k = 0; k < EX_PER; k++) buf[k * 32 + lane] = nxt[k];
		#pragma unroll
		for (int k = 0; k < EX_PER; k++)                               // next chunk: in flight while lane 0 folds
			{ const uint64_t i = c0 + EX_CHUNK + (uint64_t) k * 32 + lane;  nxt[k] = (i < sd.hi) ? in[i] : 0.0; }
		__syncwarp ();
		if (lane == 0)
			{
			#pragma unroll 8
			for (uint32_t k = 0; k < n; k++) { run = __dadd_rn (run, buf[k]);  buf[k] = run; }
			}
		__syncwarp ();
		#pragma unroll
		for (int k = 0; k < EX_PER; k++) { const uint32_t e = k * 32 + lane;  if (e < n) out[c0 + e] = buf[e]; }
		__syncwarp ();
		}
	}

int gdsp_cumulative_sum_exact (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out)
	{
	for (int s = 0; s < L->nseg; s++)
		GDSP_REQUIRE (L->h[s].pos0 == 0, "cumulativesum --exact-order: whole chromosomes only (not slab-sharded)");
	k_cumsum_seq<<<L->nseg, 32, 0, c->stream>>> (L->d, in, out);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---- slidingsum -------------------------------------------------------------------------------
// step ix = 0 .. vLen+hOff-1:  if (ix < vLen) sum += v[ix];  if (ix >= W) sum -= v[ix-W];
//                              if (ix >= hOff) s[ix-hOff] = sum;          out = s / denom
__global__ void __launch_bounds__(32)
k_sliding_seq (const SegDev* __restrict__ segs, const double* __restrict__ in, double* __restrict__ out,
               uint32_t W, double denom)
	{
	__shared__ double sa[EX_CHUNK], sb[EX_CHUNK];
	const SegDev sd = segs[blockIdx.x];
	const int lane = threadIdx.x;
	const uint64_t vLen = sd.hi - sd.lo, hOff = (W - 1) / 2, steps = vLen + hOff;
	double sum = 0.0;
	for (uint64_t x0 = 0; x0 < steps; x0 += EX_CHUNK)
		{
		const uint32_t n = (uint32_t) ((steps - x0 < EX_CHUNK) ? (steps - x0) : EX_CHUNK);
		#pragma unroll
		for (int k = 0; k < EX_PER; k++)
			{
			const uint64_t ix = x0 + (uint64_t) k * 32 + lane;
			sa[k * 32 + lane] = (ix < vLen) ? in[sd.lo + ix] : 0.0;
			sb[k * 32 + lane] = (ix >= W && ix - W < vLen) ? in[sd.lo + ix - W] : 0.0;
			}
		__syncwarp ();
		if (lane == 0)
			{
			if (x0 + n <= vLen && x0 >= W)                             // the steady state: both updates, no tests
				{
				#pragma unroll 8
				for (uint32_t k = 0; k < n; k++) { sum = __dadd_rn (sum, sa[k]);  sum = __dsub_rn (sum, sb[k]);  sa[k] = sum; }
				}
			else
				for (uint32_t k = 0; k < n; k++)
					{
					const uint64_t ix = x0 + k;
					if (ix < vLen) sum = __dadd_rn (sum, sa[k]);
					if (ix >= W)   sum = __dsub_rn (sum, sb[k]);
					sa[k] = sum;
					}
			}
		__syncwarp ();
		#pragma unroll
		for (int k = 0; k < EX_PER; k++)
			{
			const uint32_t e = k * 32 + lane;
			const uint64_t ix = x0 + e;
			if (e < n && ix >= hOff) out[sd.lo + ix - hOff] = __ddiv_rn (sa[e], denom);
			}
		__syncwarp ();
		}
	}

int gdsp_sliding_sum_exact (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, uint32_t W, double denom)
	{
	for (int s = 0; s < L->nseg; s++)
		GDSP_REQUIRE (L->h[s].pos0 == 0 && L->h[s].hi - L->h[s].lo == L->h[s].chrom_len,
		              "slidingsum --exact-order: whole chromosomes only (not slab-sharded)");
	k_sliding_seq<<<L->nseg, 32, 0, c->stream>>> (L->d, in, out, W, denom);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---- clump: the prefix sums P and their running minimum M, in the reference's order ------------------
// (replaces pass A of the stored-prefix path, gdsp_clump.cu; passes B and C only compare these values)
__global__ void __launch_bounds__(32)
k_clump_prefix_seq (const SegDev* __restrict__ segs, const double* __restrict__ sig, double T, int above,
                    double* __restrict__ P, double* __restrict__ M, int* __restrict__ segAllNeg)
	{
	__shared__ double buf[EX_CHUNK], bm[EX_CHUNK];
	const SegDev sd = segs[blockIdx.x];
	const int lane = threadIdx.x;
	double run = 0.0, mn = 0.0;                                        // P[-1] = 0 takes part in every prefix minimum
	bool nonNeg = false;
	for (uint64_t c0 = sd.lo; c0 < sd.hi; c0 += EX_CHUNK)
		{
		const uint32_t n = (uint32_t) ((sd.hi - c0 < EX_CHUNK) ? (sd.hi - c0) : EX_CHUNK);
		#pragma unroll
		for (int k = 0; k < EX_PER; k++)
			{
			const uint32_t e = k * 32 + lane;
			double d = 0.0;
			if (e < n)
				{
				const double v = sig[c0 + e];
				d = above ? __dsub_rn (v, T) : __dsub_rn (T, v);
				if (d >= 0.0) nonNeg = true;
				}
			buf[e] = d;
			}
		__syncwarp ();
		if (lane == 0)
			{
			#pragma unroll 8
			for (uint32_t k = 0; k < n; k++)
				{
				run = __dadd_rn (run, buf[k]);
				if (run < mn) mn = run;                                // clump.c:601-607
				buf[k] = run;  bm[k] = mn;
				}
			}
		__syncwarp ();
		#pragma unroll
		for (int k = 0; k < EX_PER; k++) { const uint32_t e = k * 32 + lane;  if (e < n) { P[c0 + e] = buf[e];  M[c0 + e] = bm[e]; } }
		__syncwarp ();
		}
	if (__any_sync (0xffffffffu, nonNeg) && lane == 0) segAllNeg[blockIdx.x] = 0;
	}

int gdsp_clump_prefix_exact (gdsp_ctx* c, gdsp_layout* L, const double* sig, double T, int above,
                             double* P, double* M, int* segAllNeg)
	{
	k_clump_prefix_seq<<<L->nseg, 32, 0, c->stream>>> (L->d, sig, T, above, P, M, segAllNeg);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
