// gdsp_window.cu -- block sum, sliding sum and Hann smoothing.
//
// Replaces op_window_sum_apply (sum.c:211-252), op_sliding_sum_apply
// (sum.c:420-463) and op_smooth_apply (sum.c:616-676).
//
// The sliding sum and the direct FIR stage a tile of the signal plus its window halo in shared memory; the
// block sum streams (one thread per block), as does the shared-product FIR of gdsp_smooth_sym.cu (one thread
// per strip).  Cells outside the chromosome's readable range [dlo,dhi) count as 0.0, which is exactly the
// reference's "beyond the ends is zero" rule.
#include "gdsp_common.cuh"
#include <stdlib.h>

// ---------------------------------------------------------------------------
// staging helper: smem[idx(j)] = in[g0 + j] for j in [0,count), zero outside
// [dlo,dhi).  g0 may be "negative" (before the buffer start): signed 64-bit.
// ---------------------------------------------------------------------------

template <int PADSHIFT>
__device__ __forceinline__ uint32_t pad_idx (uint32_t j)
	{ return PADSHIFT ? j + (j >> PADSHIFT) : j; }

template <int PADSHIFT>
__device__ __forceinline__ void stage_zero_padded (double* smem, const double* __restrict__ in,
                                                   int64_t g0, uint32_t count,
                                                   uint64_t dlo, uint64_t dhi)
	{
	for (uint32_t j = threadIdx.x; j < count; j += blockDim.x)
		{
		int64_t g = g0 + (int64_t) j;
		double v = 0.0;
		if (g >= (int64_t) dlo && g < (int64_t) dhi) v = __ldg (in + g);
		smem[pad_idx<PADSHIFT> (j)] = v;
		}
	}

// ---------------------------------------------------------------------------
// sliding sum:  out[c] = sum in[c+h-W+1 .. c+h] / denom   (h = (W-1)/2)
// Tile-local inclusive prefix sums in shared memory; a window sum is the
// difference of two prefix values.  Sums of integer-valued (or dyadic) signals
// are exact in any association order, so the result is bit-identical to the
// reference's running sum there; for general reals it differs only by
// rounding (tile-local prefixes keep that below 1e-13 relative).
// ---------------------------------------------------------------------------

#define SS_THREADS 256
#define SS_TILE    4096

__global__ void __launch_bounds__(SS_THREADS)
k_sliding_sum (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
               const double* __restrict__ in, double* __restrict__ out,
               uint32_t W, uint32_t strip, double denom)
	{
	extern __shared__ double sm[];
	double* P = sm;                               // staged cells, then prefix sums (count = SS_TILE+W-1)
	__shared__ double s_warp[SS_THREADS / 32];

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SS_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < SS_TILE) ? (sd.hi - t0) : SS_TILE);
	const uint32_t h  = (W - 1) / 2;
	const uint32_t reachL = W - 1 - h;
	const uint32_t count  = n + W - 1;

	stage_tile<0> (P, in, (int64_t) t0 - (int64_t) reachL, count, sd.dlo, sd.dhi, 0.0);
	__syncthreads ();

	// each thread owns `strip` consecutive cells (strip is odd: conflict-free 64-bit accesses)
	const uint32_t j0 = threadIdx.x * strip;
	const uint32_t j1 = (j0 + strip < count) ? j0 + strip : count;
	double tot = 0.0;
	for (uint32_t j = j0; j < j1; j++) tot += P[j];

	// exclusive scan of the strip totals across the block
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double inc = tot;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		double up = shfl_up_f64 (inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads ();
	double carry = 0.0;
	for (int w = 0; w < warp; w++) carry += s_warp[w];
	double ex = shfl_up_f64 (inc, 1);
	if (lane == 0) ex = 0.0;
	double run = carry + ex;
	for (uint32_t j = j0; j < j1; j++) { run += P[j];  P[j] = run; }
	__syncthreads ();

	// window for local output c covers staged cells [c, c+W-1].  An IEEE division is ~20 instructions per cell; x / 1.0
	// is x, so the default denominator skips it (the kernel is otherwise 8 instructions per cell)
	if (denom == 1.0)
		for (uint32_t c = threadIdx.x; c < n; c += blockDim.x)
			{
			double hi = P[c + W - 1];
			double lo = (c > 0) ? P[c - 1] : 0.0;
			out[t0 + c] = hi - lo;
			}
	else
		for (uint32_t c = threadIdx.x; c < n; c += blockDim.x)
			{
			double hi = P[c + W - 1];
			double lo = (c > 0) ? P[c - 1] : 0.0;
			out[t0 + c] = (hi - lo) / denom;
			}
	}

// ---------------------------------------------------------------------------
// block sum: blocks [kW,(k+1)W) counted from chromosome coordinate 0; the
// left-to-right total of each block (seeded with its first cell, sum.c:230-236)
// over denom goes to the block's first cell, zeroVal to the rest.
// One THREAD owns one block: it streams the block's cells from global memory (256-bit loads when the block is
// 32-byte aligned, as with --window=100), folds them in the reference's order, and writes the block back.
// Neighbouring lanes own neighbouring blocks, so a warp touches one contiguous stretch of 32*W cells.
// (Round 1 staged 4096-cell tiles in shared memory and let one thread per block fold its row: 68 instructions per
// cell between the staging index arithmetic and the write-back, 0.70 of the HBM peak with 40 of 256 threads
// busy in the fold.  This form needs about 2 per cell.)
// ---------------------------------------------------------------------------

#define BS_THREADS 256

__device__ __forceinline__ int bs_block_seg (const uint64_t* __restrict__ base, int nseg, uint64_t t)
	{
	int lo = 0, hi = nseg - 1;                     // last s with base[s] <= t
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1;  if (__ldg (base + mid) <= t) lo = mid; else hi = mid - 1; }
	return lo;
	}

__global__ void __launch_bounds__(BS_THREADS)
k_block_sum (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t nblocks,
             double* __restrict__ sig, uint32_t W, double denom, int denomActual, double zeroVal)
	{
	const uint64_t gid = (uint64_t) blockIdx.x * BS_THREADS + threadIdx.x;
	if (gid >= nblocks) return;
	const int seg = bs_block_seg (base, nseg, gid);
	const SegDev sd = segs[seg];
	// blocks are counted in chromosome coordinates: the piece owns coordinates [pos0, pos0+len); its first block is
	// the one containing pos0
	const uint64_t c0   = ((uint64_t) sd.pos0 / W + (gid - __ldg (base + seg))) * W;       // chromosome coordinate of the block
	const uint64_t cEnd = ((uint64_t) sd.pos0 + (sd.dhi - sd.lo) < (uint64_t) sd.chromLen)
	                    ? (uint64_t) sd.pos0 + (sd.dhi - sd.lo) : (uint64_t) sd.chromLen;  // readable end
	uint64_t c1 = c0 + W;  if (c1 > cEnd) c1 = cEnd;
	if (c0 >= c1) return;
	const uint32_t bl = (uint32_t) (c1 - c0);
	// cell index of chromosome coordinate c:  sd.lo + (c - pos0)   (may precede sd.lo: halo / dlo)
	const int64_t g0 = (int64_t) sd.lo + ((int64_t) c0 - (int64_t) sd.pos0);
	const bool readable = (g0 >= (int64_t) sd.dlo) && (g0 + (int64_t) bl <= (int64_t) sd.dhi);
	const bool owned    = (g0 >= (int64_t) sd.lo)  && (g0 + (int64_t) bl <= (int64_t) sd.hi);

	double t;
	if (readable && owned && (bl & 3u) == 0 && (g0 & 3) == 0)
		{
		double* p = sig + g0;
		double a[4];
		ldg_stream4 (p, a[0], a[1], a[2], a[3]);
		t = a[0];  t += a[1];  t += a[2];  t += a[3];
		#pragma unroll 4
		for (uint32_t k = 4; k < bl; k += 4)
			{
			ldg_stream4 (p + k, a[0], a[1], a[2], a[3]);
			t += a[0];  t += a[1];  t += a[2];  t += a[3];
			}
		const double y = denomActual ? t / (double) bl : t / denom;
		stg_stream4 (p, y, zeroVal, zeroVal, zeroVal);
		for (uint32_t k = 4; k < bl; k += 4) stg_stream4 (p + k, zeroVal, zeroVal, zeroVal, zeroVal);
		return;
		}
	// any other block (odd widths, the ends of a chromosome, a slab boundary): cell by cell; cells outside the readable
	// range count as 0.0, only owned cells are written
	{
	const int64_t g = g0;
	t = (g >= (int64_t) sd.dlo && g < (int64_t) sd.dhi) ? sig[g] : 0.0;
	}
	for (uint32_t k = 1; k < bl; k++)
		{
		const int64_t g = g0 + (int64_t) k;
		t += (g >= (int64_t) sd.dlo && g < (int64_t) sd.dhi) ? sig[g] : 0.0;
		}
	const double y = denomActual ? t / (double) bl : t / denom;
	for (uint32_t k = 0; k < bl; k++)
		{
		const int64_t g = g0 + (int64_t) k;
		if (g >= (int64_t) sd.lo && g < (int64_t) sd.hi) sig[g] = (k == 0) ? y : zeroVal;
		}
	}

// The same with FOUR lanes per block, for widths that are a multiple of 4 (--window=100): lane q of a group takes
// cells 4q..4q+3 of every 16-cell chunk, so the group's 256-bit loads cover one whole 128-byte line and a warp-wide
// load touches 8 lines instead of 32 (with a thread per block the L1 tag stage was the limit: ncu lg_throttle 107
// warps per issue, 0.80 of the HBM peak).  The running sum still visits the cells in the reference's order: it is
// handed from lane to lane with a shuffle every four additions.
__global__ void __launch_bounds__(BS_THREADS)
k_block_sum4 (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t nblocks,
              double* __restrict__ sig, uint32_t W, double denom, int denomActual, double zeroVal)
	{
	const int lane = threadIdx.x & 31, q = lane & 3, g4 = lane & ~3;
	const uint64_t gid = ((uint64_t) blockIdx.x * BS_THREADS + threadIdx.x) >> 2;
	bool have = (gid < nblocks);
	SegDev sd;  uint32_t bl = 0;  int64_t g0 = 0;
	if (have)
		{
		const int seg = bs_block_seg (base, nseg, gid);
		sd = segs[seg];
		const uint64_t c0   = ((uint64_t) sd.pos0 / W + (gid - __ldg (base + seg))) * W;
		const uint64_t cEnd = ((uint64_t) sd.pos0 + (sd.dhi - sd.lo) < (uint64_t) sd.chromLen)
		                    ? (uint64_t) sd.pos0 + (sd.dhi - sd.lo) : (uint64_t) sd.chromLen;
		uint64_t c1 = c0 + W;  if (c1 > cEnd) c1 = cEnd;
		have = (c0 < c1);
		bl = have ? (uint32_t) (c1 - c0) : 0u;
		g0 = (int64_t) sd.lo + ((int64_t) c0 - (int64_t) sd.pos0);
		}
	const bool whole = have && (g0 >= (int64_t) sd.lo) && (g0 + (int64_t) bl <= (int64_t) sd.hi)      // readable and owned,
	                        && (bl & 3u) == 0 && (g0 & 3) == 0;                                     // and made of aligned quads
	// the four lanes of a group agree on `whole` (same block); groups of a warp may differ
	double t = 0.0;
	if (__any_sync (0xffffffffu, whole))
		{
		double* const p = sig + g0;
		const uint32_t nq = whole ? bl >> 2 : 0u;                   // quads of the block; quad j belongs to lane j & 3
		uint32_t nqMax = nq;
		#pragma unroll
		for (int d = 16; d >= 4; d >>= 1) { const uint32_t o = __shfl_xor_sync (0xffffffffu, nqMax, d);  nqMax = (o > nqMax) ? o : nqMax; }
		double a[4] = { 0.0, 0.0, 0.0, 0.0 }, b[4];
		if ((uint32_t) q < nq) ldg_stream4 (p + 4 * q, a[0], a[1], a[2], a[3]);
		for (uint32_t j0 = 0; j0 < nqMax; j0 += 4)                  // one 16-cell chunk per trip
			{
			const uint32_t jn = j0 + 4 + q;                         // this lane's quad of the next chunk: in flight during the hops
			if (jn < nq) ldg_stream4 (p + 4 * jn, b[0], b[1], b[2], b[3]);
			#pragma unroll
			for (int h = 0; h < 4; h++)
				{
				const double tin = __shfl_sync (0xffffffffu, t, g4 + ((h + 3) & 3));
				if (q == h && j0 + h < nq)
					{
					if (j0 + h == 0) t = a[0];                          // the fold is seeded with the first cell (sum.c:230)
					else { t = tin;  t += a[0]; }
					t += a[1];  t += a[2];  t += a[3];
					}
				}
			#pragma unroll
			for (int k = 0; k < 4; k++) a[k] = b[k];
			}
		const double tot = __shfl_sync (0xffffffffu, t, g4 + (int) ((nq - 1) & 3u));              // the lane that added the last quad
		if (whole)
			{
			const double y = denomActual ? tot / (double) bl : tot / denom;
			for (uint32_t j = q; j < nq; j += 4)
				stg_stream4 (p + 4 * j, (j == 0) ? y : zeroVal, zeroVal, zeroVal, zeroVal);
			}
		}
	if (whole || !have || q != 0) return;
	// any other block (the ends of a chromosome, a slab boundary): lane 0 of the group, cell by cell; cells outside the
	// readable range count as 0.0, only owned cells are written
	{
	const int64_t g = g0;
	t = (g >= (int64_t) sd.dlo && g < (int64_t) sd.dhi) ? sig[g] : 0.0;
	}
	for (uint32_t k = 1; k < bl; k++)
		{
		const int64_t g = g0 + (int64_t) k;
		t += (g >= (int64_t) sd.dlo && g < (int64_t) sd.dhi) ? sig[g] : 0.0;
		}
	const double y = denomActual ? t / (double) bl : t / denom;
	for (uint32_t k = 0; k < bl; k++)
		{
		const int64_t g = g0 + (int64_t) k;
		if (g >= (int64_t) sd.lo && g < (int64_t) sd.hi) sig[g] = (k == 0) ? y : zeroVal;
		}
	}

// big windows (W larger than a tile, or --window=chromosome): tile partial sums,
// then one thread per block folds the partials in order
#define BSB_TILE 4096
__global__ void __launch_bounds__(256)
k_block_sum_big_partial (const SegDev* __restrict__ segs, const uint64_t* __restrict__ wbase, int nseg,
                         const double* __restrict__ sig, uint32_t W, double* __restrict__ partial)
	{
	// wbase: prefix over segments of (number of W-blocks * tilesPerBlock)
	__shared__ double s_red[8];
	int seg;  uint64_t tin;
	tile_to_seg (wbase, nseg, blockIdx.x, seg, tin);
	const SegDev sd = segs[seg];
	const uint32_t tilesPerBlock = (W + BSB_TILE - 1) / BSB_TILE;
	const uint64_t blk  = tin / tilesPerBlock, tib = tin % tilesPerBlock;
	const uint64_t b0   = blk * (uint64_t) W;                   // chromosome coordinate (pos0 must be 0 here)
	uint64_t b1 = b0 + W;  if (b1 > sd.chromLen) b1 = sd.chromLen;
	uint64_t c0 = b0 + tib * BSB_TILE, c1 = c0 + BSB_TILE;  if (c1 > b1) c1 = b1;
	double t = 0.0;
	for (uint64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x) t += sig[sd.lo + c];
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) t += shfl_xor_f64 (t, d);
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		double a = 0.0;
		for (int w = 0; w < 8; w++) a += s_red[w];
		partial[blockIdx.x] = a;
		}
	}

__global__ void __launch_bounds__(256)
k_block_sum_big_write (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, const uint64_t* __restrict__ wbase,
                       int nseg, double* __restrict__ sig, uint32_t W, const double* __restrict__ partial,
                       double denom, int denomActual, double zeroVal)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint32_t tilesPerBlock = (W + BSB_TILE - 1) / BSB_TILE;
	uint64_t c0 = tis * BSB_TILE, c1 = c0 + BSB_TILE;
	if (c1 > sd.hi - sd.lo) c1 = sd.hi - sd.lo;
	for (uint64_t c = c0 + threadIdx.x; c < c1; c += blockDim.x)
		{
		double v = zeroVal;
		if (c % W == 0)
			{
			uint64_t blk = c / W;
			uint64_t bl  = (c + W <= sd.chromLen) ? W : sd.chromLen - c;
			uint32_t nt  = (uint32_t) ((bl + BSB_TILE - 1) / BSB_TILE);
			const double* p = partial + wbase[seg] + blk * tilesPerBlock;
			double t = p[0];
			for (uint32_t k = 1; k < nt; k++) t += p[k];
			v = denomActual ? t / (double) bl : t / denom;
			}
		sig[sd.lo + c] = v;
		}
	}

// ---------------------------------------------------------------------------
// Hann smoothing (direct FIR).  out[i] = fold over k ascending of
// acc + (w[k] * in[i-h+k]), product and sum rounded separately (no FMA), i.e.
// the reference's order.  Zero-staged cells outside the chromosome contribute
// +0.0 products, which leave the accumulator unchanged -- identical to the
// reference skipping those taps (sum.c:655-662).
// FP64-issue bound: 2*W double-precision instructions per base.
//
// Each thread produces SM_R consecutive outputs from a sliding register window,
// so one shared-memory load feeds SM_R multiply-adds; the staged tile uses a
// padded index (one pad cell per 8) so that the stride-8 accesses of a warp are
// bank-conflict free.
// ---------------------------------------------------------------------------

#define SM_THREADS 128
#define SM_R       8
#define SM_TILE    (SM_THREADS * SM_R)       // 1024 outputs per tile
#define SM_KC      512                        // taps per staged chunk

__global__ void __launch_bounds__(SM_THREADS)
k_smooth (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
          const double* __restrict__ in, double* __restrict__ out,
          uint32_t W, const double* __restrict__ taps)
	{
	__shared__ double s_w[SM_KC + 8];
	__shared__ double s_x[(SM_TILE + SM_KC + 16) + ((SM_TILE + SM_KC + 16) >> 3) + 2];

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SM_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < SM_TILE) ? (sd.hi - t0) : SM_TILE);
	const uint32_t h  = (W - 1) / 2;

	double acc[SM_R];
	#pragma unroll
	for (int r = 0; r < SM_R; r++) acc[r] = 0.0;

	const uint32_t i0 = threadIdx.x * SM_R;

	for (uint32_t kc = 0; kc < W; kc += SM_KC)
		{
		const uint32_t kn = (W - kc < SM_KC) ? (W - kc) : SM_KC;      // taps in this chunk
		__syncthreads ();
		for (uint32_t k = threadIdx.x; k < kn; k += blockDim.x) s_w[k] = taps[kc + k];
		// staged cell j  <->  input index t0 - h + kc + j ; need j in [0, SM_TILE + kn - 1 + 8)
		stage_tile<3> (s_x, in, (int64_t) t0 - (int64_t) h + (int64_t) kc, SM_TILE + kn + 8, sd.dlo, sd.dhi, 0.0);
		__syncthreads ();

		double x[SM_R], y[SM_R];
		#pragma unroll
		for (int m = 0; m < SM_R; m++) x[m] = s_x[pad_idx<3> (i0 + m)];

		uint32_t k0 = 0;
		for (; k0 + SM_R <= kn; k0 += SM_R)
			{
			#pragma unroll
			for (int m = 0; m < SM_R; m++) y[m] = s_x[pad_idx<3> (i0 + k0 + SM_R + m)];
			#pragma unroll
			for (int u = 0; u < SM_R; u++)
				{
				const double w = s_w[k0 + u];
				#pragma unroll
				for (int r = 0; r < SM_R; r++)
					{
					const double xv = (u + r < SM_R) ? x[u + r] : y[u + r - SM_R];
					acc[r] = __dadd_rn (acc[r], __dmul_rn (w, xv));
					}
				}
			#pragma unroll
			for (int m = 0; m < SM_R; m++) x[m] = y[m];
			}
		if (k0 < kn)
			{
			#pragma unroll
			for (int m = 0; m < SM_R; m++) y[m] = s_x[pad_idx<3> (i0 + k0 + SM_R + m)];
			#pragma unroll
			for (int u = 0; u < SM_R; u++)
				{
				if (k0 + u < kn)
					{
					const double w = s_w[k0 + u];
					#pragma unroll
					for (int r = 0; r < SM_R; r++)
						{
						const double xv = (u + r < SM_R) ? x[u + r] : y[u + r - SM_R];
						acc[r] = __dadd_rn (acc[r], __dmul_rn (w, xv));
						}
					}
				}
			}
		}

	// SM_R consecutive outputs per thread: 64 contiguous bytes, 128-bit stores
	double* o = out + t0 + i0;
	if (i0 + SM_R <= n)
		{
		#pragma unroll
		for (int r = 0; r < SM_R; r += 4) stg_stream4 (o + r, acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
		}
	else
		{
		#pragma unroll
		for (int r = 0; r < SM_R; r++) if (i0 + r < n) o[r] = acc[r];
		}
	}

// ---------------------------------------------------------------------------
// k_smooth_ct: the same FIR for W <= SC_MAXW with the taps passed BY VALUE as a
// __grid_constant__ kernel parameter.  The tap of step k is the same for every
// thread, so it is read through the uniform datapath straight from the constant
// bank (SASS: LDCU.64 UR, c[0x0][UR+off]; DMUL R, R, UR) and costs no
// shared-memory load; the only LDS left in the inner loop is the one new input
// cell per 2*R FP64 instructions.  scripts/micro/fp64_peak.cu measures what
// that buys on this part: 1.745e13 DMUL+DADD/s with constant-bank taps against
// 1.673e13 with a broadcast LDS per tap (nominal 64 lanes * 148 SMs * 1.965 GHz
// = 1.861e13).
// (A persistent variant fed by 1-D bulk async copies (cp.async.bulk + mbarrier)
// was measured at 43.9 ms on hg38 against 39.9 ms for k_smooth: the kernel is
// FP64-issue bound, so hiding the global-load latency buys nothing and the
// per-tile bookkeeping costs issue slots.  It was removed.)
// ---------------------------------------------------------------------------

#define SC_MAXW    1024
#define SC_LOGR    4            // 16 outputs per thread: hg38 W=101 35.3 ms; 8 per thread 37.1 ms
struct SmoothTaps { double w[SC_MAXW]; };

template <int LOGR>
__global__ void __launch_bounds__(SM_THREADS)
k_smooth_ct (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
             const double* __restrict__ in, double* __restrict__ out,
             uint32_t W, const __grid_constant__ SmoothTaps tp)
	{
	constexpr int      R    = 1 << LOGR;
	constexpr uint32_t TILE = SM_THREADS * R;
	__shared__ double s_x[(TILE + SM_KC + 2 * R) + ((TILE + SM_KC + 2 * R) >> LOGR) + 2];

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < TILE) ? (sd.hi - t0) : TILE);
	const uint32_t h  = (W - 1) / 2;

	double acc[R];
	#pragma unroll
	for (int r = 0; r < R; r++) acc[r] = 0.0;

	const uint32_t i0 = threadIdx.x * R;
	// padded position of staged cell i0 + q*R + m  (0 <= m < R)  is  (threadIdx.x + q) * (R + 1) + m
	const double* xrow = s_x + (size_t) threadIdx.x * (R + 1);

	for (uint32_t kc = 0; kc < W; kc += SM_KC)
		{
		const uint32_t kn = (W - kc < SM_KC) ? (W - kc) : SM_KC;      // taps in this chunk
		if (kc) __syncthreads ();
		stage_tile<LOGR> (s_x, in, (int64_t) t0 - (int64_t) h + (int64_t) kc, TILE + kn + R, sd.dlo, sd.dhi, 0.0);
		__syncthreads ();

		double x[R], y[R];
		#pragma unroll
		for (int m = 0; m < R; m++) x[m] = xrow[m];

		uint32_t k0 = 0;
		const double* yrow = xrow + (R + 1);
		for (; k0 + R <= kn; k0 += R, yrow += R + 1)
			{
			#pragma unroll
			for (int m = 0; m < R; m++) y[m] = yrow[m];
			#pragma unroll
			for (int u = 0; u < R; u++)
				{
				const double w = tp.w[kc + k0 + u];
				#pragma unroll
				for (int r = 0; r < R; r++)
					{
					const double xv = (u + r < R) ? x[u + r] : y[u + r - R];
					acc[r] = __dadd_rn (acc[r], __dmul_rn (w, xv));
					}
				}
			#pragma unroll
			for (int m = 0; m < R; m++) x[m] = y[m];
			}
		if (k0 < kn)
			{
			#pragma unroll
			for (int m = 0; m < R; m++) y[m] = yrow[m];
			#pragma unroll
			for (int u = 0; u < R; u++)
				{
				if (k0 + u < kn)
					{
					const double w = tp.w[kc + k0 + u];
					#pragma unroll
					for (int r = 0; r < R; r++)
						{
						const double xv = (u + r < R) ? x[u + r] : y[u + r - R];
						acc[r] = __dadd_rn (acc[r], __dmul_rn (w, xv));
						}
					}
				}
			}
		}

	double* o = out + t0 + i0;
	if (i0 + R <= n)
		{
		#pragma unroll
		for (int r = 0; r < R; r += 4) stg_stream4 (o + r, acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
		}
	else
		{
		#pragma unroll
		for (int r = 0; r < R; r++) if (i0 + r < n) o[r] = acc[r];
		}
	}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------

extern "C" int gdsp_sliding_sum (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out,
                                 uint32_t W, double denom)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out, "gdsp_sliding_sum: NULL argument");
	GDSP_REQUIRE (in != out, "gdsp_sliding_sum: in and out must be different buffers");
	GDSP_REQUIRE (W >= 1, "gdsp_sliding_sum: window must be positive");
	if (c->exact_order) return gdsp_sliding_sum_exact (c, L, in, out, W, denom);
	uint32_t count = SS_TILE + W - 1;
	size_t smem = (size_t) count * sizeof (double);
	GDSP_REQUIRE (smem + 1024 <= c->smem_optin,
	              "gdsp_sliding_sum: window %u needs %zu bytes of shared memory (limit %zu)", W, smem, c->smem_optin);
	uint32_t strip = (count + SS_THREADS - 1) / SS_THREADS;
	strip |= 1;
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SS_TILE, &tm));
	GDSP_CUDA (cudaFuncSetAttribute (k_sliding_sum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_sliding_sum<<<(unsigned) tm.ntiles, SS_THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, W, strip, denom);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

extern "C" int gdsp_block_sum (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint32_t W, int windowIsChrom,
                               double denom, int denomActual, double zeroVal)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig, "gdsp_block_sum: NULL argument");
	GDSP_REQUIRE (windowIsChrom || W >= 1, "gdsp_block_sum: window must be positive");
	if (!windowIsChrom && W <= 4096)
		{
		// blocks per segment, in chromosome coordinates (the first block is the one containing pos0)
		std::vector<uint64_t> base (L->nseg + 1);
		uint64_t nb = 0;
		for (int s = 0; s < L->nseg; s++)
			{
			const gdsp_seg& g = L->h[s];
			uint64_t cs = ((uint64_t) g.pos0 / W) * W, ce = (uint64_t) g.pos0 + (g.hi - g.lo);
			base[s] = nb;
			nb += (ce - cs + W - 1) / W;
			}
		base[L->nseg] = nb;
		if (nb == 0) return GDSP_OK;
		void* ws;
		GDSP_TRY (gdsp_ws (c, 2, sizeof (uint64_t) * (L->nseg + 1), &ws));
		GDSP_CUDA (cudaMemcpyAsync (ws, base.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
		GDSP_CUDA (cudaStreamSynchronize (c->stream));      // base[] is a host temporary
		if ((W & 3u) == 0 && W >= 16 && !getenv ("GDSP_BLOCKSUM_THREAD"))
			k_block_sum4<<<(unsigned) ((4 * nb + BS_THREADS - 1) / BS_THREADS), BS_THREADS, 0, c->stream>>> (L->d, (const uint64_t*) ws, L->nseg, nb, sig, W,
			                                                                                              denom, denomActual, zeroVal);
		else
			k_block_sum<<<(unsigned) ((nb + BS_THREADS - 1) / BS_THREADS), BS_THREADS, 0, c->stream>>> (L->d, (const uint64_t*) ws, L->nseg, nb, sig, W,
			                                                                                         denom, denomActual, zeroVal);
		GDSP_KERNEL_CHECK ();
		return GDSP_OK;
		}

	// big windows: whole chromosomes must be resident on this GPU
	for (int s = 0; s < L->nseg; s++)
		GDSP_REQUIRE (L->h[s].pos0 == 0 && L->h[s].hi - L->h[s].lo == L->h[s].chrom_len,
		              "gdsp_block_sum: windows above 4096 need whole chromosomes on one GPU");
	std::vector<uint64_t> wbase (L->nseg + 1);
	uint64_t np = 0;
	uint32_t maxW = 0;
	// with --window=chromosome every segment is one block; use the longest as W and clip per segment
	if (windowIsChrom) { for (int s = 0; s < L->nseg; s++) if (L->h[s].chrom_len > maxW) maxW = L->h[s].chrom_len;  W = maxW; }
	uint32_t tpb = (W + BSB_TILE - 1) / BSB_TILE;
	for (int s = 0; s < L->nseg; s++)
		{
		wbase[s] = np;
		np += (((uint64_t) L->h[s].chrom_len + W - 1) / W) * tpb;
		}
	wbase[L->nseg] = np;
	void* ws;  void* wpart;
	GDSP_TRY (gdsp_ws (c, 2, sizeof (uint64_t) * (L->nseg + 1), &ws));
	GDSP_TRY (gdsp_ws (c, 3, sizeof (double) * np, &wpart));
	GDSP_CUDA (cudaMemcpyAsync (ws, wbase.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	k_block_sum_big_partial<<<(unsigned) np, 256, 0, c->stream>>> (L->d, (const uint64_t*) ws, L->nseg, sig, W, (double*) wpart);
	GDSP_KERNEL_CHECK ();
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, BSB_TILE, &tm));
	k_block_sum_big_write<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, (const uint64_t*) ws, L->nseg, sig, W,
	                                                                   (const double*) wpart, denom, denomActual, zeroVal);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

extern "C" int gdsp_smooth (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out,
                            uint32_t W, const double* h_taps)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out && h_taps, "gdsp_smooth: NULL argument");
	GDSP_REQUIRE (in != out, "gdsp_smooth: in and out must be different buffers");
	GDSP_REQUIRE_ALIGNED (out, "gdsp_smooth");
	GDSP_REQUIRE (W >= 1, "gdsp_smooth: window must be positive");
	TileMap tm;
	// symmetric window (the reference's Hann window always is): every product of a tap pair computed once,
	// same summation order (gdsp_smooth_sym.cu)
	if (!c->smooth_direct && gdsp_smooth_sym_plan (L, in, W, h_taps))
		return gdsp_smooth_sym_launch (c, L, in, out, W, h_taps);
	if (W <= SC_MAXW)
		{
		SmoothTaps tp;                          // by-value kernel parameter (copied at launch)
		memcpy (tp.w, h_taps, sizeof (double) * W);
		GDSP_TRY (gdsp_layout_tilemap (L, SM_THREADS << SC_LOGR, &tm));
		k_smooth_ct<SC_LOGR><<<(unsigned) tm.ntiles, SM_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, W, tp);
		GDSP_KERNEL_CHECK ();
		return GDSP_OK;
		}
	// the taps are uploaded once and reused while the caller keeps passing the same window
	if (c->taps_n != W || c->taps_host == NULL || memcmp (c->taps_host, h_taps, sizeof (double) * W) != 0)
		{
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		if (c->taps_dev)  { cudaFree (c->taps_dev);  c->taps_dev = NULL; }
		if (c->taps_host) { free (c->taps_host);     c->taps_host = NULL; }
		c->taps_n = 0;
		c->taps_host = (double*) malloc (sizeof (double) * W);
		GDSP_REQUIRE (c->taps_host != NULL, "gdsp_smooth: out of host memory");
		memcpy (c->taps_host, h_taps, sizeof (double) * W);
		GDSP_CUDA (cudaMalloc (&c->taps_dev, sizeof (double) * W));
		GDSP_CUDA (cudaMemcpyAsync (c->taps_dev, c->taps_host, sizeof (double) * W, cudaMemcpyHostToDevice, c->stream));
		c->taps_n = W;
		}
	void* dt = c->taps_dev;
	GDSP_TRY (gdsp_layout_tilemap (L, SM_TILE, &tm));
	k_smooth<<<(unsigned) tm.ntiles, SM_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, W, (const double*) dt);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
