// gdsp_minmax.cu -- sliding-window extrema.  By window width:
//   localmax/localmin up to 33 cells   k_local_direct    chained compares, no extremum formed
//   up to 63 cells                      k_extrema_small   sparse-table doubling in registers
//   64 .. 2049 cells                    k_extrema_blocks  16-cell blocks per thread + sparse table over block extrema
//   up to 6145 cells                    k_extrema         van Herk / Gil-Werman with warp-shuffle segmented scans
//   beyond                              k_extrema_wide    direct scan per output
//
// Replaces op_local_maxima_apply (minmax.c:1183-1227), op_local_minima_apply
// (minmax.c:981-1022), op_best_local_max_apply (minmax.c:1616-1721) and
// op_best_local_min_apply (minmax.c:1369-1474).
//
// A tile stages nOut + Wn - 1 cells (Wn = window width); cells outside the
// chromosome's readable range are staged as the neutral element (-inf / +inf),
// which is the reference's "clip the window to the vector" rule.  The staged
// cells are cut into van Herk blocks of exactly Wn cells;
//     g[j]  = extremum from the start of j's block to j   (forward scan)
//     hs[j] = extremum from j to the end of j's block     (backward scan)
// and the window [c, c+Wn-1] is ext(hs[c], g[c+Wn-1]).  Both scans are
// block-wide segmented scans: every thread owns a strip of E consecutive staged
// cells in registers; strip aggregates are combined with warp shuffles.
#include "gdsp_common.cuh"
#include <stdlib.h>


// fmax/fmin ignore a NaN operand, which is what the reference's `if (v[j] > best)` scans do with a NaN
// neighbour; but they compile to ~8 instructions per call.  The tiled kernels therefore replace NaN by
// the neutral element once, when a staged cell enters registers (sanitize), and use the
// 3-instruction compare-select below; only a window made ENTIRELY of NaN differs (neutral instead of
// NaN), which localmax/localmin never show because they output the untouched centre value.
template <bool WANT_MAX> __device__ __forceinline__ double ext_ieee (double a, double b)
	{ return WANT_MAX ? fmax (a, b) : fmin (a, b); }
template <bool WANT_MAX> __device__ __forceinline__ double ext (double a, double b)
	{ return WANT_MAX ? ((b > a) ? b : a) : ((b < a) ? b : a); }
__device__ __forceinline__ double sanitize (double v, double neutral) { return (v == v) ? v : neutral; }

template <int LOGE> __device__ __forceinline__ uint32_t mm_pad (uint32_t j) { return j + (j >> LOGE); }

// MODE 0: out = window extremum (bestmax/bestmin)
// MODE 1: out = in unless the window extremum beats it, then fill (localmax/localmin)
template <int LOGE, int MM_THREADS, bool WANT_MAX, int MODE>
__global__ void __launch_bounds__(MM_THREADS)
k_extrema (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
           const double* __restrict__ in, double* __restrict__ out,
           uint32_t reachL, uint32_t Wn, uint32_t tileOut, double fill)
	{
	constexpr int      E    = 1 << LOGE;
	constexpr int      MM_WARPS = MM_THREADS / 32;
	constexpr uint32_t SCAP = (uint32_t) E * MM_THREADS;
	constexpr uint32_t PADN = SCAP + (SCAP >> LOGE) + 2;
	extern __shared__ double sm[];
	double* A = sm;                 // staged cells, later g
	double* B = sm + PADN;          // hs
	__shared__ double s_wv[MM_WARPS];
	__shared__ int    s_wf[MM_WARPS];

	const double NEUTRAL = WANT_MAX ? -__longlong_as_double (0x7ff0000000000000ll) * 1.0 : __longlong_as_double (0x7ff0000000000000ll);

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0   = sd.lo + tis * tileOut;
	const uint32_t nOut = (uint32_t) ((sd.hi - t0 < tileOut) ? (sd.hi - t0) : tileOut);
	const uint32_t count = nOut + Wn - 1;
	const int64_t  g0   = (int64_t) t0 - (int64_t) reachL;

	stage_tile<LOGE, true> (A, in, g0, count, sd.dlo, sd.dhi, NEUTRAL);
	for (uint32_t j = count + threadIdx.x; j < SCAP; j += MM_THREADS) A[mm_pad<LOGE> (j)] = NEUTRAL;
	__syncthreads ();

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t j0 = threadIdx.x * E;
	double a[E];
	#pragma unroll
	for (int e = 0; e < E; e++) a[e] = sanitize (A[mm_pad<LOGE> (j0 + e)], NEUTRAL);
	const uint32_t m0 = j0 % Wn;

	// ---------------- forward: blocks restart where j % Wn == 0 ----------------
	{
	double run = NEUTRAL;  int reset = 0;
	uint32_t m = m0;
	#pragma unroll
	for (int e = 0; e < E; e++)
		{
		if (m == 0) { run = a[e];  reset = 1; } else run = ext<WANT_MAX> (run, a[e]);
		if (++m == Wn) m = 0;
		}
	double val = run;  int f = reset;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		double v2 = shfl_up_f64 (val, d);
		int    f2 = __shfl_up_sync (0xffffffffu, f, d);
		if (lane >= d) { if (!f) val = ext<WANT_MAX> (v2, val);  f |= f2; }
		}
	double exv = shfl_up_f64 (val, 1);
	int    exf = __shfl_up_sync (0xffffffffu, f, 1);
	if (lane == 0) { exv = NEUTRAL;  exf = 0; }
	if (lane == 31) { s_wv[warp] = val;  s_wf[warp] = f; }
	__syncthreads ();                    // also: every thread has read its strip of A
	double cw = NEUTRAL;
	for (int w = 0; w < warp; w++) cw = s_wf[w] ? s_wv[w] : ext<WANT_MAX> (cw, s_wv[w]);
	run = exf ? exv : ext<WANT_MAX> (cw, exv);
	m = m0;
	#pragma unroll
	for (int e = 0; e < E; e++)
		{
		if (m == 0) run = a[e]; else run = ext<WANT_MAX> (run, a[e]);
		if (++m == Wn) m = 0;
		A[mm_pad<LOGE> (j0 + e)] = run;
		}
	}
	__syncthreads ();

	// ---------------- backward: blocks restart where j % Wn == Wn-1 -------------
	{
	double run = NEUTRAL;  int reset = 0;
	uint32_t m = m0 + (E - 1);  m %= Wn;          // position of the strip's last cell inside its block
	const uint32_t mLast = m;
	#pragma unroll
	for (int e = E - 1; e >= 0; e--)
		{
		if (m == Wn - 1) { run = a[e];  reset = 1; } else run = ext<WANT_MAX> (run, a[e]);
		m = (m == 0) ? Wn - 1 : m - 1;
		}
	double val = run;  int f = reset;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		double v2 = shfl_down_f64 (val, d);
		int    f2 = __shfl_down_sync (0xffffffffu, f, d);
		if (lane + d < 32) { if (!f) val = ext<WANT_MAX> (v2, val);  f |= f2; }
		}
	double exv = shfl_down_f64 (val, 1);
	int    exf = __shfl_down_sync (0xffffffffu, f, 1);
	if (lane == 31) { exv = NEUTRAL;  exf = 0; }
	if (lane == 0) { s_wv[warp] = val;  s_wf[warp] = f; }
	__syncthreads ();
	double cw = NEUTRAL;
	for (int w = MM_WARPS - 1; w > warp; w--) cw = s_wf[w] ? s_wv[w] : ext<WANT_MAX> (cw, s_wv[w]);
	run = exf ? exv : ext<WANT_MAX> (cw, exv);
	m = mLast;
	#pragma unroll
	for (int e = E - 1; e >= 0; e--)
		{
		if (m == Wn - 1) run = a[e]; else run = ext<WANT_MAX> (run, a[e]);
		m = (m == 0) ? Wn - 1 : m - 1;
		B[mm_pad<LOGE> (j0 + e)] = run;
		}
	}
	__syncthreads ();

	for (uint32_t c = threadIdx.x; c < nOut; c += MM_THREADS)
		{
		double w = ext<WANT_MAX> (B[mm_pad<LOGE> (c)], A[mm_pad<LOGE> (c + Wn - 1)]);
		if (MODE == 0) out[t0 + c] = w;
		else
			{
			double v = __ldg (in + t0 + c);
			bool beaten = WANT_MAX ? (w > v) : (w < v);
			out[t0 + c] = beaten ? fill : v;
			}
		}
	}

// ---------------------------------------------------------------------------
// small windows (W <= 64, i.e. every localmax/localmin neighbourhood in common
// use): sparse-table doubling in REGISTERS.  With P = 2^LOGP the largest power
// of two <= W, every thread builds A[j] = ext(X[j..j+P-1]) for 8 consecutive
// staged cells from 8+P-1 shared-memory loads (log2(P) rounds of register
// max), stores them, and the window [c, c+W-1] is ext(A[c], A[c+W-P]).
// ~40 instructions per base instead of ~150 for the general kernel.
// ---------------------------------------------------------------------------

#define MS_THREADS 256
#define MS_STRIP   8
#define MS_CELLS   (MS_THREADS * MS_STRIP)        // 2048 table cells per tile
#define MS_MARGIN  64
#define MS_TILE    (MS_CELLS - MS_MARGIN)         // 1984 outputs per tile

__device__ __forceinline__ uint32_t ms_pad (uint32_t j) { return j + (j >> 3); }

template <int LOGP, bool WANT_MAX, int MODE>
__global__ void __launch_bounds__(MS_THREADS)
k_extrema_small (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                 const double* __restrict__ in, double* __restrict__ out,
                 uint32_t reachL, uint32_t Wn, double fill)
	{
	constexpr int P = 1 << LOGP;
	constexpr uint32_t XN = MS_CELLS + P;                       // staged cells
	__shared__ double s_x[XN + (XN >> 3) + 2];
	__shared__ double s_a[MS_CELLS + (MS_CELLS >> 3) + 2];
	const double NEUTRAL = WANT_MAX ? -__longlong_as_double (0x7ff0000000000000ll) : __longlong_as_double (0x7ff0000000000000ll);

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0   = sd.lo + tis * MS_TILE;
	const uint32_t nOut = (uint32_t) ((sd.hi - t0 < MS_TILE) ? (sd.hi - t0) : MS_TILE);
	const int64_t  g0   = (int64_t) t0 - (int64_t) reachL;      // staged cell j <-> in[g0 + j]

	stage_tile<3, true> (s_x, in, g0, XN, sd.dlo, sd.dhi, NEUTRAL);
	__syncthreads ();

	// A[j] for the 8 cells of this thread, by doubling
	const uint32_t j0 = threadIdx.x * MS_STRIP;
	double x[MS_STRIP + P - 1];
	#pragma unroll
	for (int e = 0; e < MS_STRIP + P - 1; e++) x[e] = sanitize (s_x[ms_pad (j0 + e)], NEUTRAL);
	#pragma unroll
	for (int lev = 0; lev < LOGP; lev++)
		{
		const int step = 1 << lev;
		#pragma unroll
		for (int e = 0; e < MS_STRIP + P - 1; e++)
			if (e + step < MS_STRIP + P - 1 && e < MS_STRIP + P - (2 << lev) + 0)
				x[e] = ext<WANT_MAX> (x[e], x[e + step]);
		}
	#pragma unroll
	for (int e = 0; e < MS_STRIP; e++) s_a[ms_pad (j0 + e)] = x[e];
	__syncthreads ();

	const uint32_t shift = Wn - P;                              // 0 <= shift < P <= MS_MARGIN
	// every thread finishes its own 8 outputs: A[c] is still in registers, A[c+shift] and the centre
	// value come from the padded tables at a lane stride of 9 cells (conflict-free), and the 64 bytes
	// of results leave as two full-sector 256-bit stores
	if (j0 < nOut)
		{
		double w[MS_STRIP];
		#pragma unroll
		for (int e = 0; e < MS_STRIP; e++)
			{
			w[e] = ext<WANT_MAX> (x[e], s_a[ms_pad (j0 + e + shift)]);
			if (MODE == 1)
				{
				const double v = s_x[ms_pad (j0 + e + reachL)];
				w[e] = (WANT_MAX ? (w[e] > v) : (w[e] < v)) ? fill : v;
				}
			}
		double* o = out + t0 + j0;
		if (j0 + MS_STRIP <= nOut)
			{
			stg_stream4 (o,     w[0], w[1], w[2], w[3]);
			stg_stream4 (o + 4, w[4], w[5], w[6], w[7]);
			}
		else
			{
			#pragma unroll
			for (int e = 0; e < MS_STRIP; e++) if (j0 + e < nOut) o[e] = w[e];
			}
		}
	}

template <int LOGP, bool WANT_MAX, int MODE>
static int launch_extrema_small_t (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out,
                                   uint32_t reachL, uint32_t Wn, double fill)
	{
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, MS_TILE, &tm));
	k_extrema_small<LOGP, WANT_MAX, MODE><<<(unsigned) tm.ntiles, MS_THREADS, 0, c->stream>>>
		(L->d, tm.d_base, L->nseg, in, out, reachL, Wn, fill);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// localmax / localmin with a neighbourhood of at most 2*LD_HMAX+1 cells: the operator only asks whether SOME
// cell of the neighbourhood beats the centre (minmax.c:1163-1189 keeps v unless the window extremum differs),
// so no extremum is formed at all -- N-1 compares per cell, each one DSETP that ORs into the running
// predicate, instead of the ~70 instructions per cell of the doubling kernel above (sanitize, log2(P) rounds
// of 3-instruction compare-selects, two table passes): ncu had that kernel at 2.2 warp instructions per cell,
// 53 % issue -- as much instruction- as memory-bound.  NaN neighbours never beat anything and a NaN centre is
// never beaten (IEEE compares), which is what mapping NaN to the neutral element gave.
// ---------------------------------------------------------------------------

#define LD_THREADS 256
#define LD_STRIP   8
#define LD_TILE    (LD_THREADS * LD_STRIP)        // 2048 outputs per tile
#define LD_HMAX    16

template <int H, bool WANT_MAX>
__global__ void __launch_bounds__(LD_THREADS)
k_local_direct (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                const double* __restrict__ in, double* __restrict__ out, double fill)
	{
	constexpr uint32_t XN = LD_TILE + 2 * H;                    // staged cells
	__shared__ double s_x[XN + (XN >> 3) + 2];
	const double NEUTRAL = WANT_MAX ? -__longlong_as_double (0x7ff0000000000000ll) : __longlong_as_double (0x7ff0000000000000ll);

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0   = sd.lo + tis * LD_TILE;
	const uint32_t nOut = (uint32_t) ((sd.hi - t0 < LD_TILE) ? (sd.hi - t0) : LD_TILE);
	stage_tile<3, true> (s_x, in, (int64_t) t0 - H, XN, sd.dlo, sd.dhi, NEUTRAL);
	__syncthreads ();

	const uint32_t j0 = threadIdx.x * LD_STRIP;
	if (j0 >= nOut) return;
	double x[LD_STRIP + 2 * H];
	#pragma unroll
	for (int e = 0; e < LD_STRIP + 2 * H; e++) x[e] = s_x[ms_pad (j0 + e)];
	double w[LD_STRIP];
	#pragma unroll
	for (int e = 0; e < LD_STRIP; e++)
		{
		const double v = x[e + H];
		// one predicate, every compare ORs into it (setp.gt.or.f64): the C++ form of this loop compiled to a
		// select per compare
		int beaten;
		beaten = 0;
		#pragma unroll
		for (int d = 0; d <= 2 * H; d++)
			if (d != H)
				{
				if (WANT_MAX) asm ("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\tsetp.gt.or.f64 p, %1, %2, p;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "+r"(beaten) : "d"(x[e + d]), "d"(v));
				else          asm ("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %0, 0;\n\tsetp.lt.or.f64 p, %1, %2, p;\n\tselp.s32 %0, 1, 0, p;\n\t}" : "+r"(beaten) : "d"(x[e + d]), "d"(v));
				}
		w[e] = beaten ? fill : v;
		}
	double* o = out + t0 + j0;
	if (j0 + LD_STRIP <= nOut)
		{
		stg_stream4 (o,     w[0], w[1], w[2], w[3]);
		stg_stream4 (o + 4, w[4], w[5], w[6], w[7]);
		}
	else
		{
		#pragma unroll
		for (int e = 0; e < LD_STRIP; e++) if (j0 + e < nOut) o[e] = w[e];
		}
	}

template <int H, bool WANT_MAX>
static int launch_local_direct_h (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, uint32_t h, double fill)
	{
	if ((uint32_t) H != h) return launch_local_direct_h<(H > 1 ? H - 1 : 1), WANT_MAX> (c, L, in, out, (H > 1) ? h : 1u, fill);
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, LD_TILE, &tm));
	k_local_direct<H, WANT_MAX><<<(unsigned) tm.ntiles, LD_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, fill);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// 64 <= W <= 2049 (bestmax / bestmin, wide localmax / localmin): blocks of 16 staged cells, one per thread.
// A thread forms the prefix and suffix extrema of its block in registers (30 compare-selects per 16 cells), a
// sparse table over the 256 block extrema of the tile is built in shared memory (one compare-select per thread
// and level), and the window of output c = 16j+e is
//     ext( suffix of block j from e  [registers],  prefix of the block the window ends in  [shared memory],
//          the whole blocks in between  [two table entries] )
// where the whole blocks in between are the same for all 16 outputs of a thread up to one block at the far end:
// two table look-ups per thread.  About 30 instructions per cell; the van Herk kernel above (k_extrema: block-wide
// segmented scans in both directions, each as two passes over the strip, and a scalar output loop) needs ~56 and
// ran at 0.55 of the HBM peak.
// ---------------------------------------------------------------------------

#define XB_THREADS 256
#define XB_E       16
#define XB_CELLS   (XB_THREADS * XB_E)            // 4096 staged cells per tile
#define XB_LEVELS  8                              // table levels: windows of up to 2^7 whole blocks

__device__ __forceinline__ uint32_t xb_pad (uint32_t j) { return j + (j >> 4); }

template <bool WANT_MAX, int MODE>
__global__ void __launch_bounds__(XB_THREADS)
k_extrema_blocks (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                  const double* __restrict__ in, double* __restrict__ out,
                  uint32_t reachL, uint32_t Wn, uint32_t tileOut, double fill)
	{
	extern __shared__ double xb_smem[];
	double* const A  = xb_smem;                                 // staged cells, then the in-block prefix extrema (padded)
	double* const Tb = xb_smem + XB_CELLS + (XB_CELLS >> 4) + 2; // Tb[k * 256 + j]: extremum of blocks j .. j + 2^k - 1
	const double NEUTRAL = WANT_MAX ? -__longlong_as_double (0x7ff0000000000000ll) : __longlong_as_double (0x7ff0000000000000ll);

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0    = sd.lo + tis * tileOut;
	const uint32_t nOut  = (uint32_t) ((sd.hi - t0 < tileOut) ? (sd.hi - t0) : tileOut);
	const uint32_t count = nOut + Wn - 1;
	stage_tile<4, true> (A, in, (int64_t) t0 - (int64_t) reachL, count, sd.dlo, sd.dhi, NEUTRAL);
	for (uint32_t j = count + threadIdx.x; j < XB_CELLS; j += XB_THREADS) A[xb_pad (j)] = NEUTRAL;
	__syncthreads ();

	// own block: prefix extrema back to shared memory (in place), suffix extrema stay in registers
	const uint32_t j0 = threadIdx.x * XB_E;
	double suf[XB_E];
	{
	double a[XB_E];
	#pragma unroll
	for (int e = 0; e < XB_E; e++) a[e] = sanitize (A[xb_pad (j0) + e], NEUTRAL);
	double run = a[0];
	#pragma unroll
	for (int e = 1; e < XB_E; e++) { run = ext<WANT_MAX> (run, a[e]);  A[xb_pad (j0) + e] = run; }
	A[xb_pad (j0)] = a[0];
	Tb[threadIdx.x] = run;
	suf[XB_E - 1] = a[XB_E - 1];
	#pragma unroll
	for (int e = XB_E - 2; e >= 0; e--) suf[e] = ext<WANT_MAX> (suf[e + 1], a[e]);
	}
	// whole blocks between the first and the last block of a window: nfb or nfb+1 of them, the same for every thread
	const uint32_t nfb = ((Wn - 1) >> 4) - 1;                   // >= 2 (Wn >= 64)
	const int kTop = 31 - __clz (nfb + 1);
	for (int k = 1; k <= kTop; k++)
		{
		__syncthreads ();
		const uint32_t span = 1u << k;
		if (threadIdx.x + span <= XB_THREADS)
			Tb[k * XB_THREADS + threadIdx.x] = ext<WANT_MAX> (Tb[(k - 1) * XB_THREADS + threadIdx.x], Tb[(k - 1) * XB_THREADS + threadIdx.x + (span >> 1)]);
		}
	__syncthreads ();
	if (j0 >= nOut) return;

	// outputs 16j .. 16j+15: the window of output e ends at staged cell j0 + e + Wn - 1, in block j + 1 + nfb or the next
	const uint32_t r    = (Wn - 1) & 15u;                       // the window end crosses into the next block when e + r >= 16
	const int ka = 31 - __clz (nfb), kb = 31 - __clz (nfb + 1);
	const uint32_t b1 = threadIdx.x + 1;
	double midA = ext<WANT_MAX> (Tb[ka * XB_THREADS + b1], Tb[ka * XB_THREADS + b1 + nfb - (1u << ka)]);
	double midB = midA;
	if (r != 0 && b1 + nfb + 1 <= XB_THREADS)
		midB = ext<WANT_MAX> (Tb[kb * XB_THREADS + b1], Tb[kb * XB_THREADS + b1 + nfb + 1 - (1u << kb)]);
	double w[XB_E];
	#pragma unroll
	for (int e = 0; e < XB_E; e++)
		{
		uint32_t last = j0 + e + Wn - 1;
		if (last > XB_CELLS - 1) last = XB_CELLS - 1;           // (outputs past nOut: computed, never stored)
		const double pre = A[xb_pad (last)];
		const double mid = ((uint32_t) e + r >= 16u) ? midB : midA;
		w[e] = ext<WANT_MAX> (ext<WANT_MAX> (suf[e], pre), mid);
		}
	double* o = out + t0 + j0;
	if (MODE == 1)
		{
		const double* ctr = in + t0 + j0;
		#pragma unroll
		for (int e = 0; e < XB_E; e++)
			if (j0 + e < nOut)
				{
				const double v = __ldg (ctr + e);
				w[e] = (WANT_MAX ? (w[e] > v) : (w[e] < v)) ? fill : v;
				}
		}
	if (j0 + XB_E <= nOut)
		{
		#pragma unroll
		for (int e = 0; e < XB_E; e += 4) stg_stream4 (o + e, w[e], w[e + 1], w[e + 2], w[e + 3]);
		}
	else
		{
		#pragma unroll
		for (int e = 0; e < XB_E; e++) if (j0 + e < nOut) o[e] = w[e];
		}
	}

template <bool WANT_MAX, int MODE>
static int launch_extrema_blocks (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out,
                                  uint32_t reachL, uint32_t Wn, double fill)
	{
	uint32_t tileOut = XB_CELLS - (Wn - 1);
	tileOut &= ~63u;                                   // keep tile starts 512-byte aligned
	const size_t smem = sizeof (double) * ((size_t) XB_CELLS + (XB_CELLS >> 4) + 2 + (size_t) XB_LEVELS * XB_THREADS);
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, tileOut, &tm));
	GDSP_CUDA (cudaFuncSetAttribute (k_extrema_blocks<WANT_MAX, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_extrema_blocks<WANT_MAX, MODE><<<(unsigned) tm.ntiles, XB_THREADS, smem, c->stream>>>
		(L->d, tm.d_base, L->nseg, in, out, reachL, Wn, tileOut, fill);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// very wide windows: direct scan per output (correct for any width; slow)
template <bool WANT_MAX, int MODE>
__global__ void __launch_bounds__(256)
k_extrema_wide (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                const double* __restrict__ in, double* __restrict__ out,
                uint32_t reachL, uint32_t reachR, double fill)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	uint64_t i = sd.lo + tis * 256 + threadIdx.x;
	if (i >= sd.hi) return;
	uint64_t a = (i - sd.dlo > reachL) ? i - reachL : sd.dlo;
	uint64_t b = (sd.dhi - 1 - i > reachR) ? i + reachR : sd.dhi - 1;
	double w = in[a];
	for (uint64_t j = a + 1; j <= b; j++) w = ext_ieee<WANT_MAX> (w, in[j]);
	if (MODE == 0) out[i] = w;
	else
		{
		double v = in[i];
		bool beaten = WANT_MAX ? (w > v) : (w < v);
		out[i] = beaten ? fill : v;
		}
	}

template <int LOGE, int MM_THREADS, bool WANT_MAX, int MODE>
static int launch_extrema_t (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out,
                             uint32_t reachL, uint32_t Wn, double fill)
	{
	constexpr uint32_t SCAP = (1u << LOGE) * MM_THREADS;
	constexpr uint32_t PADN = SCAP + (SCAP >> LOGE) + 2;
	uint32_t tileOut = SCAP - (Wn - 1);
	tileOut &= ~63u;                                   // keep tile starts 512-byte aligned
	size_t smem = 2 * (size_t) PADN * sizeof (double);
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, tileOut, &tm));
	GDSP_CUDA (cudaFuncSetAttribute (k_extrema<LOGE, MM_THREADS, WANT_MAX, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_extrema<LOGE, MM_THREADS, WANT_MAX, MODE><<<(unsigned) tm.ntiles, MM_THREADS, smem, c->stream>>>
		(L->d, tm.d_base, L->nseg, in, out, reachL, Wn, tileOut, fill);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

template <bool WANT_MAX, int MODE>
static int launch_extrema (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out,
                           uint32_t reachL, uint32_t reachR, double fill)
	{
	uint64_t Wn64 = (uint64_t) reachL + reachR + 1;
	if (MODE == 1 && reachL == reachR && reachL >= 1 && reachL <= LD_HMAX)
		return launch_local_direct_h<LD_HMAX, WANT_MAX> (c, L, in, out, reachL, fill);
	if (Wn64 >= 2 && Wn64 < 64)
		{
		const uint32_t Wn = (uint32_t) Wn64;
		if (Wn < 4)  return launch_extrema_small_t<1, WANT_MAX, MODE> (c, L, in, out, reachL, Wn, fill);
		if (Wn < 8)  return launch_extrema_small_t<2, WANT_MAX, MODE> (c, L, in, out, reachL, Wn, fill);
		if (Wn < 16) return launch_extrema_small_t<3, WANT_MAX, MODE> (c, L, in, out, reachL, Wn, fill);
		if (Wn < 32) return launch_extrema_small_t<4, WANT_MAX, MODE> (c, L, in, out, reachL, Wn, fill);
		return launch_extrema_small_t<5, WANT_MAX, MODE> (c, L, in, out, reachL, Wn, fill);
		}
	if (Wn64 >= 64 && Wn64 <= 2049 && !getenv ("GDSP_EXTREMA_VANHERK"))
		return launch_extrema_blocks<WANT_MAX, MODE> (c, L, in, out, reachL, (uint32_t) Wn64, fill);
	if (Wn64 <= 2049) return launch_extrema_t<4, 256, WANT_MAX, MODE> (c, L, in, out, reachL, (uint32_t) Wn64, fill);
	if (Wn64 <= 6145) return launch_extrema_t<4, 512, WANT_MAX, MODE> (c, L, in, out, reachL, (uint32_t) Wn64, fill);
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, 256, &tm));
	k_extrema_wide<WANT_MAX, MODE><<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, reachL, reachR, fill);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

extern "C" int gdsp_local_extrema (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out,
                                   uint32_t neighborhood, int wantMax, double fill)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out, "gdsp_local_extrema: NULL argument");
	GDSP_REQUIRE (in != out, "gdsp_local_extrema: in and out must be different buffers");
	GDSP_REQUIRE_ALIGNED (out, "gdsp_local_extrema");
	GDSP_REQUIRE (neighborhood >= 1, "gdsp_local_extrema: neighborhood must be positive");
	uint32_t h = (neighborhood - 1) / 2;
	return wantMax ? launch_extrema<true, 1>  (c, L, in, out, h, h, fill)
	               : launch_extrema<false, 1> (c, L, in, out, h, h, fill);
	}

extern "C" int gdsp_best_extrema (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out,
                                  uint32_t window, int wantMax)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out, "gdsp_best_extrema: NULL argument");
	GDSP_REQUIRE (in != out, "gdsp_best_extrema: in and out must be different buffers");
	GDSP_REQUIRE_ALIGNED (out, "gdsp_best_extrema");
	GDSP_REQUIRE (window >= 1, "gdsp_best_extrema: window must be positive");
	uint32_t l = (window - 1) / 2, r = (window - 1) - l;
	return wantMax ? launch_extrema<true, 0>  (c, L, in, out, l, r, 0.0)
	               : launch_extrema<false, 0> (c, L, in, out, l, r, 0.0);
	}
