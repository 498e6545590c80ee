// gdsp_clump.cu -- clump / anticlump.
//
// Replaces clump_search (clump.c:494-736).  The reference walks each chromosome
// once with a prefix sum P of d = v-T (T-v for anticlump), a stack of strictly
// decreasing prefix minima and a moving pointer to find, for every end i, the
// earliest start j with P[j] <= P[i]; [j+1,i] is marked when i-j >= minLength,
// and every maximal marked run is trimmed to its first..last cell with v>=T.
//
// Closed form used here (DESIGN.md derives it): with M[j] = min(0, P[0..j]),
//     end i is "valid"   iff  i+1 >= Lmin and M[i-Lmin] <= P[i]      (M[-1] = 0)
//     cell p is marked   iff  max{ P[i] : i >= p, i valid } >= M[p-1]
// so the whole search is three streaming passes of segmented scans:
//   A  forward : P (prefix sum) and M (prefix min of P)              24 B/bp
//   B  backward: suffix max of valid P -> marked(p); suffix "a cell with v>=T
//                follows inside the marked run" flag                  ~33 B/bp
//   C  forward : prefix "a cell with v>=T precedes inside the run" flag,
//                output one/zero                                       9 B/bp
// Prefix sums of integer-valued / dyadic signals are exact in any order, so the
// result is bit-identical to the reference there (the reference's P is a
// sequentially rounded sum; for general reals the comparison P[j]<=P[i] can
// differ at ties within rounding -- BASELINE.json's stated tolerance class).
#include <stdlib.h>
#include "gdsp_common.cuh"
#include "gdsp_scan.cuh"

#define CL_THREADS 256
#define CL_WARPS   8
#define CL_ROWS    4
#define CL_TILE    (CL_WARPS * CL_ROWS * 128)       // 4096

template <typename T> struct Shfl;
template <> struct Shfl<double>
	{
	static __device__ __forceinline__ double up (double v, int d)  { return shfl_up_f64 (v, d); }
	static __device__ __forceinline__ double idx (double v, int s) { return shfl_idx_f64 (v, s); }
	};
template <> struct Shfl<int>
	{
	static __device__ __forceinline__ int up (int v, int d)  { return __shfl_up_sync (0xffffffffu, v, d); }
	static __device__ __forceinline__ int idx (int v, int s) { return __shfl_sync (0xffffffffu, v, s); }
	};

// element e of the tile lives in x[r][c] of thread (warp,lane): e = warp*512 + r*128 + lane*4 + c
__device__ __forceinline__ uint32_t cl_elem (int warp, int lane, int r, int c)
	{ return warp * (CL_ROWS * 128) + r * 128 + lane * 4 + c; }

// Inclusive scan of the tile in element order with a (possibly non-commutative)
// associative op(earlier, later).  On return x holds tile-local inclusive values,
// warpExcl the fold of all earlier warps (identity for warp 0) and tileAgg the
// fold of the whole tile; the caller folds (tileCarry, warpExcl, x).
template <typename T, typename Op>
__device__ __forceinline__ void tile_scan (T x[CL_ROWS][4], T identity, Op op, T* s_warp, T& warpExcl, T& tileAgg)
	{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	T rowCarry = identity;
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		x[r][1] = op (x[r][0], x[r][1]);  x[r][2] = op (x[r][1], x[r][2]);  x[r][3] = op (x[r][2], x[r][3]);
		T g = x[r][3];
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			T up = Shfl<T>::up (g, d);
			if (lane >= d) g = op (up, g);
			}
		T ex = Shfl<T>::up (g, 1);
		if (lane == 0) ex = identity;
		const T pre = op (rowCarry, ex);
		#pragma unroll
		for (int c = 0; c < 4; c++) x[r][c] = op (pre, x[r][c]);
		rowCarry = op (rowCarry, Shfl<T>::idx (g, 31));
		}
	__syncthreads ();                 // s_warp may still be read from a previous scan
	if (lane == 31) s_warp[warp] = rowCarry;
	__syncthreads ();
	warpExcl = identity;  tileAgg = identity;
	#pragma unroll
	for (int w = 0; w < CL_WARPS; w++)
		{
		T t = s_warp[w];
		if (w < warp) warpExcl = op (warpExcl, t);
		tileAgg = op (tileAgg, t);
		}
	}

// segmented OR state: bit0 = value, bit1 = "a break occurred" (later element wins across a break)
__device__ __forceinline__ int seg_or (int a, int b) { return (b & 2) ? b : ((a & 2) | ((a | b) & 1)); }

struct ClumpWork
	{
	double*        P;          // prefix sums
	double*        M;          // prefix minima (min(0, P[0..i]))
	unsigned char* F;          // per cell: bit0 marked, bit1 qualifying cell follows in run, bit2 qualifying
	int*           segAllNeg;  // per segment: 1 while every d<0
	};

// ---------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, 4)
k_clump_a (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
           const double* __restrict__ sig, double T, int above, ClumpWork wk,
           ScanStatus<double> stSum, ScanStatus<double> stMin)
	{
	__shared__ double s_warp[CL_WARPS];
	__shared__ double s_carry[2];
	__shared__ int    s_anyNonNeg;
	const uint32_t tile = scan_take_ticket (stSum.ticket);
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_anyNonNeg = 0;
	__syncthreads ();

	double x[CL_ROWS][4];
	bool nonNeg = false;
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const uint32_t e = cl_elem (warp, lane, r, c);
			double d = 0.0;
			if (e < n)
				{
				const double v = sig[t0 + e];
				d = above ? __dsub_rn (v, T) : __dsub_rn (T, v);
				if (d >= 0.0) nonNeg = true;
				}
			x[r][c] = d;
			}
	if (nonNeg) s_anyNonNeg = 1;

	double warpExcl, tileAgg;
	tile_scan<double> (x, 0.0, [] (double a, double b) { return a + b; }, s_warp, warpExcl, tileAgg);
	if (threadIdx.x < 32)
		{
		const double e = scan_lookback<double> (stSum, tile, tis == 0, tileAgg, 0.0, [] (double a, double b) { return a + b; });
		if (threadIdx.x == 0)
			{
			s_carry[0] = e;
			if (s_anyNonNeg) atomicAnd (&wk.segAllNeg[seg], 0);
			}
		}
	__syncthreads ();
	const double addP = s_carry[0] + warpExcl;

	double m[CL_ROWS][4];
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const uint32_t e = cl_elem (warp, lane, r, c);
			x[r][c] = addP + x[r][c];
			m[r][c] = (e < n) ? x[r][c] : __longlong_as_double (0x7ff0000000000000ll);
			}
	double wExM, tAggM;
	tile_scan<double> (m, __longlong_as_double (0x7ff0000000000000ll), [] (double a, double b) { return (b < a) ? b : a; },
	                   s_warp, wExM, tAggM);
	if (threadIdx.x < 32)
		{
		const double e = scan_lookback<double> (stMin, tile, tis == 0, tAggM, __longlong_as_double (0x7ff0000000000000ll),
		                                        [] (double a, double b) { return (b < a) ? b : a; });
		if (threadIdx.x == 0) s_carry[1] = e;
		}
	__syncthreads ();
	const double carryM = fmin (fmin (s_carry[1], wExM), 0.0);        // P[-1] = 0 takes part in every prefix minimum

	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		const uint32_t e0 = cl_elem (warp, lane, r, 0);
		if (e0 >= n) continue;
		double mm[4];
		#pragma unroll
		for (int c = 0; c < 4; c++) mm[c] = fmin (carryM, m[r][c]);
		if (e0 + 4 <= n)
			{
			stg_stream4 (wk.P + t0 + e0, x[r][0], x[r][1], x[r][2], x[r][3]);
			stg_stream4 (wk.M + t0 + e0, mm[0], mm[1], mm[2], mm[3]);
			}
		else
			for (int c = 0; c < 4 && e0 + c < n; c++) { wk.P[t0 + e0 + c] = x[r][c];  wk.M[t0 + e0 + c] = mm[c]; }
		}
	}

// ---------------------------------------------------------------------------
// pass B (backward: tile element e <-> cell t1-1-e)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, 4)
k_clump_b (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
           const double* __restrict__ sig, double T, int above, uint32_t minLength, double relLength,
           ClumpWork wk, ScanStatus<double> stMax, ScanStatus<int> stOr)
	{
	__shared__ double s_warp[CL_WARPS];
	__shared__ int    s_warpI[CL_WARPS];
	__shared__ double s_carryD;
	__shared__ int    s_carryI;
	// reversed tile order: ticket k handles the k-th tile from the END of the launch, so that
	// every tile a block waits on (the tiles to its right) has already started
	const uint32_t ticket = scan_take_ticket (stMax.ticket);
	const uint64_t tile = ntiles - 1 - ticket;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t tilesInSeg = base[seg + 1] - base[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const uint64_t t1 = t0 + n;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const double NEG = -__longlong_as_double (0x7ff0000000000000ll);

	uint32_t Lmin = minLength;
	if (relLength > 0.0)
		{
		uint32_t rl = (uint32_t) (relLength * sd.chromLen);       // clump.c:516-522
		if (rl > Lmin) Lmin = rl;
		}

	// status arrays are indexed by ticket order (position from the end), so the "previous" tile of
	// the scan is ticket-1 and the first tile of a segment's scan is that segment's LAST tile
	const bool firstOfScan = (tis == tilesInSeg - 1);

	double q[CL_ROWS][4];       // valid P or -inf
	double pm[CL_ROWS][4];      // M[p-1]
	int    ql[CL_ROWS][4];      // qualifying cell?
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const uint32_t e = cl_elem (warp, lane, r, c);
			double qq = NEG, mprev = 0.0;  int qu = 0;
			if (e < n)
				{
				const uint64_t cell = t1 - 1 - e;
				const uint64_t i = cell - sd.lo;                       // index inside the chromosome piece
				const double P = wk.P[cell];
				if (i + 1 >= (uint64_t) Lmin)
					{
					const double mj = (i >= (uint64_t) Lmin) ? wk.M[cell - Lmin] : 0.0;    // M[i-Lmin], M[-1]=0
					if (mj <= P) qq = P;
					}
				mprev = (i > 0) ? wk.M[cell - 1] : 0.0;
				const double v = sig[cell];
				qu = above ? (v >= T) : (v <= T);
				}
			q[r][c] = qq;  pm[r][c] = mprev;  ql[r][c] = qu;
			}

	double wEx, tAgg;
	tile_scan<double> (q, NEG, [] (double a, double b) { return (b > a) ? b : a; }, s_warp, wEx, tAgg);
	if (threadIdx.x < 32)
		{
		const double e = scan_lookback<double> (stMax, ticket, firstOfScan, tAgg, NEG, [] (double a, double b) { return (b > a) ? b : a; });
		if (threadIdx.x == 0) s_carryD = e;
		}
	__syncthreads ();
	const double carryQ = fmax (s_carryD, wEx);

	// marked(p) and the backward segmented OR of "qualifying" inside marked runs
	int st[CL_ROWS][4];
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const uint32_t e = cl_elem (warp, lane, r, c);
			const bool marked = (e < n) && (fmax (carryQ, q[r][c]) >= pm[r][c]);
			// unmarked cell: break with value 0; marked: value = qualifying
			st[r][c] = (e < n) ? (marked ? (ql[r][c] & 1) : 2) : 0;       // cells past the tile: neutral (no break, 0)
			ql[r][c] |= marked ? 4 : 0;                                    // remember marked in bit2 of ql
			}
	int wExI, tAggI;
	tile_scan<int> (st, 0, [] (int a, int b) { return seg_or (a, b); }, s_warpI, wExI, tAggI);
	if (threadIdx.x < 32)
		{
		const int e = scan_lookback<int> (stOr, ticket, firstOfScan, tAggI, 0, [] (int a, int b) { return seg_or (a, b); });
		if (threadIdx.x == 0) s_carryI = e;
		}
	__syncthreads ();
	const int carryI = seg_or (s_carryI, wExI);

	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		const uint32_t e0 = cl_elem (warp, lane, r, 0);
		if (e0 >= n) continue;
		unsigned int fb[4];
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const int s = seg_or (carryI, st[r][c]);
			const int marked = (ql[r][c] >> 2) & 1;
			fb[c] = (unsigned int) (marked | ((marked & s & 1) << 1) | ((ql[r][c] & 1) << 2));
			}
		// elements e0..e0+3 are the cells t1-1-e0 down to t1-4-e0: one 32-bit store when that group is
		// whole and 4-aligned (byte stores cost a 32-byte sector transaction each)
		if (e0 + 4 <= n && ((t1 - e0) & 3) == 0)
			*reinterpret_cast<unsigned int*> (wk.F + (t1 - 4 - e0)) = fb[3] | (fb[2] << 8) | (fb[1] << 16) | (fb[0] << 24);
		else
			{
			#pragma unroll
			for (int c = 0; c < 4; c++) if (e0 + c < n) wk.F[t1 - 1 - (e0 + c)] = (unsigned char) fb[c];
			}
		}
	}

// ---------------------------------------------------------------------------
// pass C (forward)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CL_THREADS, 4)
k_clump_c (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
           double* __restrict__ sig, double oneVal, double zeroVal, ClumpWork wk, ScanStatus<int> stOr)
	{
	__shared__ int s_warpI[CL_WARPS];
	__shared__ int s_carryI;
	const uint32_t tile = scan_take_ticket (stOr.ticket);
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool allNeg = wk.segAllNeg[seg] != 0;               // clump.c:545-565: nothing can clump

	int f[CL_ROWS][4], st[CL_ROWS][4];
	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		const uint32_t e0 = cl_elem (warp, lane, r, 0);
		unsigned int packed = 0;
		if (e0 + 4 <= n) packed = *reinterpret_cast<const unsigned int*> (wk.F + t0 + e0);     // t0+e0 is 4-aligned
		else for (int c = 0; c < 4 && e0 + c < n; c++) packed |= (unsigned int) wk.F[t0 + e0 + c] << (8 * c);
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			f[r][c] = (packed >> (8 * c)) & 255;
			const bool in = (e0 + c < n);
			st[r][c] = in ? ((f[r][c] & 1) ? ((f[r][c] >> 2) & 1) : 2) : 0;
			}
		}
	int wExI, tAggI;
	tile_scan<int> (st, 0, [] (int a, int b) { return seg_or (a, b); }, s_warpI, wExI, tAggI);
	if (threadIdx.x < 32)
		{
		const int e = scan_lookback<int> (stOr, tile, tis == 0, tAggI, 0, [] (int a, int b) { return seg_or (a, b); });
		if (threadIdx.x == 0) s_carryI = e;
		}
	__syncthreads ();
	const int carryI = seg_or (s_carryI, wExI);

	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		const uint32_t e0 = cl_elem (warp, lane, r, 0);
		if (e0 >= n) continue;
		double y[4];
		#pragma unroll
		for (int c = 0; c < 4; c++)
			{
			const int s = seg_or (carryI, st[r][c]);
			const bool one = !allNeg && (f[r][c] & 1) && (f[r][c] & 2) && (s & 1);
			y[c] = one ? oneVal : zeroVal;
			}
		if (e0 + 4 <= n)
			{
			stg_stream4 (sig + t0 + e0, y[0], y[1], y[2], y[3]);
			}
		else
			for (int c = 0; c < 4 && e0 + c < n; c++) sig[t0 + e0 + c] = y[c];
		}
	}

// ===========================================================================
// Fast path (every chromosome's minimum length <= CLF_MAX_HALO): the prefix arrays
// P and M are never stored.
//   k_clump_groups     every warp folds one 512-cell group to {sum of d, minimum of its
//                      group-relative prefix sums} -- a pure streaming read, no chain
//   k_clump_groupscan  one block per chromosome turns those into the carry of every group:
//                      {P before the group, M before the group}
//   k_clump_mark       rebuilds P and M of its tile and of the `Lmin` cells before it from
//                      the carries (the halo is the tail of the previous tile: L2 hits),
//                      needs ONE chained scan (the suffix maximum) and leaves two bits per
//                      cell (marked, marked-and-qualifying)
//   k_clump_tilesum / k_clump_tilecarry / k_clump_emit
//                      the trimming of every marked run to its first..last qualifying cell
//                      is carry propagation on those bit words: `fill upwards from the seeds
//                      through the mask` is (M & ~(M + S)) | S on a 32-cell word, a word
//                      passes a carry on like a full adder's generate/propagate pair, and
//                      the pairs are folded per tile, per chromosome and per word, in both
//                      directions
// Rounding is monotone, so min_i fl(c + x_i) = fl(c + min_i x_i): the group minima published
// by the first kernel give exactly the minima of the prefix sums the mark kernel computes
// (P = carry + group-local prefix, the same association in both kernels).
// DRAM traffic: 8 + 8 + 8 B/bp (+0.6 of bits and carries) instead of 66, one look-back chain
// instead of five.
// ===========================================================================
#define CLF_MAX_HALO 4096                 // cells; larger minimum lengths take the stored-prefix passes above
#define CLF_GROUP    512                  // cells per carry group: one warp, 16 consecutive cells per lane
#define CLF_GROUPS   (CL_TILE / CLF_GROUP)    // 8
#define CLF_WORDS    (CL_TILE / 32)       // 128 bit words per tile

struct ClumpFast
	{
	double2*       carry;      // per group (tile*8 + warp): first {group sum, group-relative prefix minimum}, then
	                           // {P before the group, M before the group} (M includes P[-1] = 0)
	uint32_t*      Bm;         // per word (tile*128 + w): marked
	uint32_t*      Bq;         // marked and qualifying (v >= T, <= T for anticlump)
	unsigned char* tsum;       // per tile: bit0/1 generate/propagate upwards, bit2/3 downwards
	unsigned char* tcin;       // per tile: bit0 carry entering from below, bit1 from above
	unsigned char* tquiet;     // per tile: 1 = no cell of the tile can be a valid end (k_clump_classify, from the records alone)
	double*        gmid;       // per group: its prefix sum after 256 cells, relative to the group (NaN when the group holds a
	                           // qualifying cell or the tile is not whole): with the carries, P at every 256-cell boundary
	int*           segAllNeg;
	// slab-sharded chromosomes (gdsp_clump_slab_*); all NULL for whole chromosomes
	const double2*       segCarryIn;  // per segment: {P, M} just before the segment's first cell
	double*              segSufMax;   // per segment: maximum valid P over the owned cells (out)
	const unsigned char* segFlags;    // per segment: bit0 = the first tile holds the neighbour's cells (halo tile)
	};

#define CLF_INF (__longlong_as_double (0x7ff0000000000000ll))

__device__ __forceinline__ double dmin2 (double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double dmax2 (double a, double b) { return (b > a) ? b : a; }

// d = v-T (T-v for anticlump) of the 16 cells of a lane: cells [c0, c0+16) of the signal, `valid` of them
// inside the tile (the others read as d = 0); returns the bit mask of qualifying cells (v >= T, resp.
// v <= T, which is d >= 0: a difference of two distinct doubles never rounds to zero)
template <bool FULL, bool ABOVE>
__device__ __forceinline__ unsigned clf_load16 (const double* __restrict__ sig, uint64_t c0, int valid, double T, double x[16])
	{
	if (FULL || valid >= 16)
		{
		#pragma unroll
		for (int k = 0; k < 16; k += 4) ldg_stream4 (sig + c0 + k, x[k], x[k + 1], x[k + 2], x[k + 3]);
		}
	else
		{
		#pragma unroll
		for (int k = 0; k < 16; k++) x[k] = (k < valid) ? sig[c0 + k] : T;
		}
	unsigned q = 0;
	#pragma unroll
	for (int k = 0; k < 16; k++)
		{
		x[k] = ABOVE ? __dsub_rn (x[k], T) : __dsub_rn (T, x[k]);
		if (!FULL && k >= valid) x[k] = 0.0;
		else if (x[k] >= 0.0) q |= 1u << k;
		}
	return q;
	}

// Inclusive prefix sums inside the 512-cell group a warp holds (lane l: cells 16l..16l+15).  The group
// and the mark kernel both use exactly this association, so they compute the same values bit for bit.
__device__ __forceinline__ void group_sum_scan (double x[16], double& total)
	{
	const int lane = threadIdx.x & 31;
	#pragma unroll
	for (int k = 1; k < 16; k++) x[k] = x[k - 1] + x[k];
	double g = x[15];
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const double up = shfl_up_f64 (g, d);
		if (lane >= d) g = up + g;
		}
	double ex = shfl_up_f64 (g, 1);
	if (lane == 0) ex = 0.0;
	#pragma unroll
	for (int k = 0; k < 16; k++) x[k] = ex + x[k];
	total = shfl_idx_f64 (g, 31);
	}

// ---- group aggregates ----------------------------------------------------------
template <bool ABOVE>
__global__ void __launch_bounds__(CL_THREADS, 4)
k_clump_groups (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                const double* __restrict__ sig, double T, ClumpFast wk)
	{
	const uint64_t tile = blockIdx.x;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t e0 = warp * CLF_GROUP + lane * 16;
	const int valid = (e0 >= n) ? 0 : (int) ((n - e0 < 16) ? (n - e0) : 16);

	double x[16], total;
	const unsigned q = clf_load16<false, ABOVE> (sig, t0 + e0, valid, T, x);
	group_sum_scan (x, total);
	double m = CLF_INF;
	#pragma unroll
	for (int k = 0; k < 16; k++) if (k < valid) m = dmin2 (m, x[k]);
	#pragma unroll
	for (int d = 16; d >= 1; d >>= 1) m = dmin2 (m, shfl_xor_f64 (m, d));
	const bool any = __any_sync (0xffffffffu, q != 0);
	// (for k_clump_classify) the prefix sum half way through the group, if the group holds no qualifying cell
	if (lane == 15) wk.gmid[tile * CLF_GROUPS + warp] = (any || n != CL_TILE) ? __longlong_as_double (0x7ff8000000000000ll) : x[15];
	if (lane == 0)
		{
		wk.carry[tile * CLF_GROUPS + warp] = make_double2 (total, m);
		// clump.c:545-565; the flag is read first: 6 M atomics on 24 addresses would serialise
		if (any && *((volatile int*) &wk.segAllNeg[seg]) != 0) atomicAnd (&wk.segAllNeg[seg], 0);
		}
	}

// ---- group carries: one block per chromosome -------------------------------------
#define CLF_GS_THREADS 1024
__global__ void __launch_bounds__(CLF_GS_THREADS)
k_clump_groupscan (const uint64_t* __restrict__ base, ClumpFast wk)
	{
	__shared__ double s_w[32];
	const uint64_t g0 = base[blockIdx.x] * CLF_GROUPS, g1 = base[blockIdx.x + 1] * CLF_GROUPS;
	const uint64_t chunk = (g1 - g0 + CLF_GS_THREADS - 1) / CLF_GS_THREADS;
	const uint64_t lo = (g0 + threadIdx.x * chunk < g1) ? g0 + threadIdx.x * chunk : g1;
	const uint64_t hi = (lo + chunk < g1) ? lo + chunk : g1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	// 1: sum of the chunk, exclusive prefix over the block
	double sum = 0.0;
	for (uint64_t g = lo; g < hi; g++) sum = sum + wk.carry[g].x;
	double inc = sum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const double up = shfl_up_f64 (inc, d);
		if (lane >= d) inc = up + inc;
		}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads ();
	double wex = 0.0;
	for (int w = 0; w < warp; w++) wex = wex + s_w[w];
	double ex = shfl_up_f64 (inc, 1);
	if (lane == 0) ex = 0.0;
	double run0 = wex + ex;                            // P before the chunk
	double mex0 = 0.0;                                 // P[-1] = 0 takes part in every prefix minimum
	if (wk.segCarryIn != NULL)                         // a slab piece: the sums continue the left neighbour's
		{
		const double2 ci = wk.segCarryIn[blockIdx.x];
		run0 = ci.x + run0;  mex0 = ci.y;
		}
	__syncthreads ();

	// 2: minimum prefix sum inside the chunk, exclusive prefix minimum over the block
	double run = run0, mn = CLF_INF;
	for (uint64_t g = lo; g < hi; g++)
		{
		const double2 a = wk.carry[g];
		mn = dmin2 (mn, run + a.y);
		run = run + a.x;
		}
	double minc = mn;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const double up = shfl_up_f64 (minc, d);
		if (lane >= d) minc = dmin2 (up, minc);
		}
	if (lane == 31) s_w[warp] = minc;
	__syncthreads ();
	double mex = mex0;
	for (int w = 0; w < warp; w++) mex = dmin2 (mex, s_w[w]);
	double e2 = shfl_up_f64 (minc, 1);
	if (lane != 0) mex = dmin2 (mex, e2);

	// 3: carries, in place
	run = run0;
	for (uint64_t g = lo; g < hi; g++)
		{
		const double2 a = wk.carry[g];
		wk.carry[g] = make_double2 (run, mex);
		mex = dmin2 (mex, run + a.y);
		run = run + a.x;
		}
	}

// ---- marks ---------------------------------------------------------------------
// shared-memory slot of the prefix minimum of the cell `j` cells after the first halo cell: one pad per
// 16 cells, so that lanes 16 cells apart are 17 slots apart (conflict-free 64-bit accesses both when a
// lane writes its own cells and when it reads cells shifted by Lmin)
__device__ __forceinline__ uint32_t clf_slot (uint32_t j) { return j + (j >> 4); }

// prefix minima of a group's prefix sums P[k] (cells past `valid` are skipped), including the carry
template <bool FULL>
__device__ __forceinline__ void clf_group_min_store (const double P[16], int valid, double carryM, double* s_row)
	{
	const int lane = threadIdx.x & 31;
	double m = CLF_INF;
	#pragma unroll
	for (int k = 0; k < 16; k++) if (FULL || k < valid) m = dmin2 (m, P[k]);
	double g = m;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const double up = shfl_up_f64 (g, d);
		if (lane >= d) g = dmin2 (up, g);
		}
	double ex = shfl_up_f64 (g, 1);
	if (lane == 0) ex = CLF_INF;
	m = dmin2 (ex, carryM);                            // everything before this lane's cells
	#pragma unroll
	for (int k = 0; k < 16; k++)
		{
		if (FULL || k < valid) m = dmin2 (m, P[k]);
		s_row[lane * 17 + k] = m;
		}
	}

// FAST: the tile is whole and at least Lmin cells away from the chromosome's first cell -- every cell is
// valid and no index test is needed (all but the first and last tiles of a chromosome)
template <bool FAST, bool ABOVE>
__device__ __forceinline__ void clf_mark_tile (const double* __restrict__ sig, double T, const ClumpFast& wk, const ScanStatus<double>& stMax,
                                               double* s_M, double* s_warp, double* s_carryD,
                                               uint64_t tile, uint32_t ticket, bool firstOfScan, uint64_t c0, uint64_t t0, uint32_t n,
                                               uint32_t Lmin, uint32_t hrows, double* sufMaxOut, bool haveLook, double lookE)
	{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const double NEG = -CLF_INF;
	const uint32_t hcells = hrows * CLF_GROUP;

	// own group: P stays in registers, M goes to shared memory
	const uint32_t e0 = warp * CLF_GROUP + lane * 16;
	const int valid = FAST ? 16 : ((e0 >= n) ? 0 : (int) ((n - e0 < 16) ? (n - e0) : 16));
	double P[16];
	unsigned qual;
	double* const ownRow = s_M + (hrows + warp) * (CLF_GROUP / 16 * 17) + lane * 17;
		{
		double tot;
		qual = clf_load16<FAST, ABOVE> (sig, t0 + e0, valid, T, P);
		group_sum_scan (P, tot);
		const double2 cr = wk.carry[tile * CLF_GROUPS + warp];
		#pragma unroll
		for (int k = 0; k < 16; k++) P[k] = cr.x + P[k];
		clf_group_min_store<FAST> (P, valid, cr.y, s_M + (hrows + warp) * (CLF_GROUP / 16 * 17));
		}
	__syncthreads ();

	// q = P where the cell is a valid end, -inf elsewhere; then its suffix maximum inside the lane.
	// The cell Lmin before cell k of this lane: slot(J0 + k), J0 = hcells + e0 - Lmin = A + r0 with A a
	// multiple of 16 and r0 = (-Lmin) mod 16 the same for every lane, so slot = slot(A) + u + (u >> 4),
	// u = r0 + k: the per-cell part of the address is warp-uniform
	const uint32_t r0 = (0u - Lmin) & 15u;
	const uint32_t A = hcells + e0 - Lmin - r0;
	const double* const shifted = s_M + A + (A >> 4);
	const uint64_t i0 = c0 + e0;                                  // index of the lane's first cell inside the chromosome
	#pragma unroll
	for (int k = 0; k < 16; k++)
		{
		double qq = NEG;
		if (FAST)
			{
			const uint32_t u = r0 + k;
			if (shifted[u + (u >> 4)] <= P[k]) qq = P[k];
			}
		else if (k < valid && i0 + k + 1 >= (uint64_t) Lmin)
			{
			const uint32_t u = r0 + k;
			const double mj = (i0 + k >= (uint64_t) Lmin) ? shifted[u + (u >> 4)] : 0.0;     // M[i-Lmin], M[-1] = 0
			if (mj <= P[k]) qq = P[k];
			}
		P[k] = qq;
		}
	#pragma unroll
	for (int k = 14; k >= 0; k--) P[k] = dmax2 (P[k], P[k + 1]);
	double g = P[0];
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const double dn = shfl_down_f64 (g, d);
		if (lane + d < 32) g = dmax2 (g, dn);
		}
	double ex = shfl_down_f64 (g, 1);
	if (lane == 31) ex = NEG;
	if (lane == 0) s_warp[warp] = g;
	__syncthreads ();
	double wEx = NEG, tAgg = NEG;
	#pragma unroll
	for (int w = 0; w < CL_WARPS; w++)
		{
		const double t = s_warp[w];
		if (w > warp) wEx = dmax2 (wEx, t);
		tAgg = dmax2 (tAgg, t);
		}
	if (threadIdx.x < 32)
		{
		// (haveLook: the tile already took part in the look-back chain with the aggregate -inf, see k_clump_mark)
		const double e = haveLook ? lookE
		               : scan_lookback<double> (stMax, ticket, firstOfScan, tAgg, NEG, [] (double a, double b) { return (b > a) ? b : a; });
		if (threadIdx.x == 0)
			{
			*s_carryD = e;
			if (sufMaxOut != NULL) *sufMaxOut = dmax2 (e, tAgg);   // first owned tile of a slab piece: the piece's maximum
			}
		}
	__syncthreads ();
	const double cq = dmax2 (dmax2 (*s_carryD, wEx), ex);         // everything after this lane's cells

	// two bits per cell; a 32-cell word is shared by two lanes.  M[p-1] of cell k >= 1 is the slot before
	// its own; of cell 0 it is the last slot of the previous lane's 16 (one pad in between)
	unsigned mk = 0;
	#pragma unroll
	for (int k = 0; k < 16; k++)
		if (FAST || k < valid)
			{
			double mp = (k == 0) ? ownRow[-2] : ownRow[k - 1];
			if (!FAST && i0 + k == 0) mp = 0.0;                   // M[-1] = 0
			if (dmax2 (cq, P[k]) >= mp) mk |= 1u << k;
			}
	// Bq holds the qualifying bit of EVERY cell (the trimming kernels and the slab fix-up AND it with Bm)
	unsigned wm = mk << ((lane & 1) * 16), wq = qual << ((lane & 1) * 16);
	wm |= __shfl_xor_sync (0xffffffffu, wm, 1);
	wq |= __shfl_xor_sync (0xffffffffu, wq, 1);
	if ((lane & 1) == 0)
		{
		const uint64_t w = tile * CLF_WORDS + warp * (CLF_GROUP / 32) + (lane >> 1);
		wk.Bm[w] = wm;  wk.Bq[w] = wq;
		}
	}

// ---- tiles that cannot hold a valid end, decided from the group records alone (no signal read) -------------
// A group without a qualifying cell (every d < 0) has a prefix sum that never rises (rounding is monotone).  The
// records give P at every 256-cell boundary: the carries at the group boundaries, carry + gmid half way.  Take a
// 256-cell half h of the tile; i-Lmin for its cells i lies at least two halves back (Lmin >= 512), so a whole half
// lies in between, and if the recorded P falls STRICTLY across it -- P at the end of the half holding (last cell of
// h) - Lmin is above P before h -- then P[i-Lmin] > P[i] for every cell of h, as computed, without looking at a cell.
// If all groups from the one holding (first cell of the group) - Lmin to the group itself never rise, and the
// prefix minimum BEFORE the first of them is above P before the group, then M[i-Lmin] = min (that minimum,
// P[..i-Lmin]) > P[i]: no cell of the group is a valid end.  One thread per tile; k_clump_mark reads the flag with
// its first loads.
__device__ __forceinline__ double clf_half_p (const ClumpFast& wk, uint64_t H)    // P before 256-cell half H (global index)
	{
	const double2 cr = wk.carry[H >> 1];
	return (H & 1) ? cr.x + wk.gmid[H >> 1] : cr.x;
	}

__global__ void __launch_bounds__(256)
k_clump_classify (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
                  uint32_t minLength, double relLength, ClumpFast wk)
	{
	const uint64_t tile = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (tile >= ntiles) return;
	int lo = 0, hi = nseg - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1;  if (base[mid] <= tile) lo = mid; else hi = mid - 1; }
	const SegDev sd = segs[lo];
	const uint64_t tis = tile - base[lo], tilesInSeg = base[lo + 1] - base[lo];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint64_t c0 = (uint64_t) sd.pos0 + tis * CL_TILE;
	uint32_t Lmin = minLength;
	if (relLength > 0.0) { const uint32_t rl = (uint32_t) (relLength * sd.chromLen);  if (rl > Lmin) Lmin = rl; }
	// (the groups the test looks back at must belong to this segment: a slab piece starts at chromosome coordinate
	// pos0 > 0, and what lies before its first tile in the record arrays is another segment)
	bool pass = (sd.hi - t0 >= CL_TILE) && (Lmin >= CLF_GROUP) && (tis * CL_TILE >= (uint64_t) Lmin) && (tis + 1 < tilesInSeg);
	// (the first own tile of a slab piece also reports the piece's suffix maximum: it takes the full path)
	if (wk.segSufMax != NULL && wk.segFlags != NULL && tis == (uint64_t) (wk.segFlags[lo] & 1u)) pass = false;
	// (slab pieces: the tile of the neighbour's cells in front of a piece is never marked at all, k_clump_mark returns
	// before it looks at this flag; its groups do serve as the Lmin cells before the piece's first own tile)
	for (int w = 0; w < CLF_GROUPS && pass; w++)
		{
		const uint64_t G  = tile * CLF_GROUPS + w;
		const uint64_t a  = c0 + (uint64_t) w * CLF_GROUP;                          // first cell of the group (chromosome index)
		const uint64_t ga = G - ((a >> 9) - ((a - Lmin) >> 9));                      // group of the cell Lmin before it (512-aligned groups)
		pass = (wk.carry[ga].y > wk.carry[G].x);
		for (uint64_t g = ga; g <= G && pass; g++) { const double m = wk.gmid[g];  pass = (m == m); }   // none of them holds a qualifying cell
		for (int h = 0; h < 2 && pass; h++)
			{
			const uint64_t H  = 2 * G + h;
			const uint64_t bh = a + 256u * h + 255u;                                  // last cell of the half
			const uint64_t Hl = H - ((bh >> 8) - ((bh - Lmin) >> 8));                  // half holding the cell Lmin before it
			pass = (clf_half_p (wk, Hl + 1) > clf_half_p (wk, H));
			}
		}
	wk.tquiet[tile] = pass ? 1 : 0;
	}

#ifndef GDSP_CLUMP_MARK_OCC
#define GDSP_CLUMP_MARK_OCC 4
#endif
template <bool ABOVE>
__global__ void __launch_bounds__(CL_THREADS, GDSP_CLUMP_MARK_OCC)
k_clump_mark (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
              const double* __restrict__ sig, double T, uint32_t minLength, double relLength,
              ClumpFast wk, ScanStatus<double> stMax)
	{
	extern __shared__ double s_M[];                    // (hrows + 8) * 544 prefix minima
	__shared__ double s_warp[CL_WARPS];
	__shared__ double s_carryD;
	// reversed tile order: ticket k handles the k-th tile from the END of the launch, so that every tile a
	// block waits on (the tiles to its right) has already started
	const uint32_t ticket = scan_take_ticket (stMax.ticket);
	const uint64_t tile = ntiles - 1 - ticket;

	// A tile k_clump_classify found without a valid end has the aggregate -inf in the suffix-maximum chain -- published
	// at once, so the tiles on its left do not wait for any work here -- and with e the maximum valid prefix sum to its
	// right, a cell p is marked iff e >= M[p-1] >= M at the tile's end: if e is below that, the tile's words are zero and
	// the signal is never read.  On a thresholded track that is mostly below the threshold (clump after open/close in
	// BASELINE config 4) that is nearly every tile, so this comes first: it needs the tile's index and nothing else
	// (a quiet tile is a whole one, neither the last of its chromosome nor a slab piece's halo or first own tile).
	bool haveLook = false;  double lookE = 0.0;
	if (wk.tquiet[tile] != 0)
		{
		__shared__ double s_look;
		const double mEnd = wk.carry[(tile + 1) * CLF_GROUPS].y;   // M before the next tile = M at this tile's end
		if (threadIdx.x < 32)
			{
			const double e = scan_lookback<double> (stMax, ticket, false, -CLF_INF, -CLF_INF, [] (double a, double b) { return (b > a) ? b : a; });
			if (threadIdx.x == 0) s_look = e;
			}
		__syncthreads ();
		lookE = s_look;  haveLook = true;
		if (lookE < mEnd)
			{
			if (threadIdx.x < CLF_WORDS) { wk.Bm[tile * CLF_WORDS + threadIdx.x] = 0u;  wk.Bq[tile * CLF_WORDS + threadIdx.x] = 0u; }
			return;
			}
		}

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t tilesInSeg = base[seg + 1] - base[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool firstOfScan = (tis == tilesInSeg - 1);
	const uint64_t c0 = (uint64_t) sd.pos0 + tis * CL_TILE;       // chromosome index of the tile's first cell
	const uint32_t haloTiles = (wk.segFlags != NULL) ? (wk.segFlags[seg] & 1u) : 0u;
	if (tis < haloTiles)
		{
		// the neighbour's cells (they only supply the prefix minima of the next tile's halo): no marks, and
		// nothing waits on this tile in the look-back chain (the tile after it in scan order starts a segment)
		if (threadIdx.x < CLF_WORDS) { wk.Bm[tile * CLF_WORDS + threadIdx.x] = 0u;  wk.Bq[tile * CLF_WORDS + threadIdx.x] = 0u; }
		return;
		}
	double* const sufMaxOut = (wk.segSufMax != NULL && tis == haloTiles) ? &wk.segSufMax[seg] : NULL;

	uint32_t Lmin = minLength;
	if (relLength > 0.0)
		{
		uint32_t rl = (uint32_t) (relLength * sd.chromLen);       // clump.c:516-522
		if (rl > Lmin) Lmin = rl;
		}
	const uint32_t reach = (Lmin > 0) ? Lmin : 1;                 // M[p-1] is needed even when Lmin is 0
	const uint32_t hrows = (reach + CLF_GROUP - 1) / CLF_GROUP;   // <= 8 (the host checked): at most one per warp

	// prefix minima of the halo groups (the tail of the previous tile of this chromosome)
	if ((uint32_t) warp < hrows && (uint64_t) (hrows - warp) * CLF_GROUP <= c0)
		{
		const uint32_t back = (hrows - warp) * CLF_GROUP;         // cells between the group's first cell and t0
		double x[16], tot;
		clf_load16<true, ABOVE> (sig, t0 - back + lane * 16, 16, T, x);
		group_sum_scan (x, tot);
		const double2 cr = wk.carry[tile * CLF_GROUPS - (hrows - warp)];
		#pragma unroll
		for (int k = 0; k < 16; k++) x[k] = cr.x + x[k];
		clf_group_min_store<true> (x, 16, cr.y, s_M + warp * (CLF_GROUP / 16 * 17));
		}

	if (n == CL_TILE && c0 >= (uint64_t) Lmin && c0 > 0)
		clf_mark_tile<true, ABOVE> (sig, T, wk, stMax, s_M, s_warp, &s_carryD, tile, ticket, firstOfScan, c0, t0, n, Lmin, hrows, sufMaxOut, haveLook, lookE);
	else
		clf_mark_tile<false, ABOVE> (sig, T, wk, stMax, s_M, s_warp, &s_carryD, tile, ticket, firstOfScan, c0, t0, n, Lmin, hrows, sufMaxOut, haveLook, lookE);
	}

// ---- run trimming on the bit words -------------------------------------------
// cells reached from the seeds S (a subset of M) going UP through set bits of M; `cin`: the cell below
// bit 0 is reached
__device__ __forceinline__ uint32_t clf_fill_up (uint32_t M, uint32_t S, uint32_t cin)
	{
	S |= cin & M & 1u;
	return (M & ~(M + S)) | S;
	}
// bit0 = the word reaches the cell above it on its own (generate), bit1 = it passes a carry on (propagate)
__device__ __forceinline__ int clf_gp (uint32_t M, uint32_t S)
	{
	return (int) (clf_fill_up (M, S, 0) >> 31) | ((M == 0xffffffffu) ? 2 : 0);
	}
// a then b (b is the later word in the direction of travel)
__device__ __forceinline__ int clf_comb (int a, int b) { return ((b | ((b >> 1) & a)) & 1) | (a & b & 2); }
__device__ __forceinline__ int clf_apply (int gp, int cin) { return (gp | ((gp >> 1) & cin)) & 1; }

// one warp per tile: fold the 128 words of the tile in both directions
__global__ void __launch_bounds__(256)
k_clump_tilesum (uint64_t ntiles, ClumpFast wk)
	{
	const uint64_t tile = (uint64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
	const int lane = threadIdx.x & 31;
	if (tile >= ntiles) return;
	const uint4 m = *reinterpret_cast<const uint4*> (wk.Bm + tile * CLF_WORDS + lane * 4);
	uint4 q = *reinterpret_cast<const uint4*> (wk.Bq + tile * CLF_WORDS + lane * 4);
	q.x &= m.x;  q.y &= m.y;  q.z &= m.z;  q.w &= m.w;
	int up = clf_comb (clf_comb (clf_comb (clf_gp (m.x, q.x), clf_gp (m.y, q.y)), clf_gp (m.z, q.z)), clf_gp (m.w, q.w));
	int dn = clf_comb (clf_comb (clf_comb (clf_gp (__brev (m.w), __brev (q.w)), clf_gp (__brev (m.z), __brev (q.z))),
	                             clf_gp (__brev (m.y), __brev (q.y))), clf_gp (__brev (m.x), __brev (q.x)));
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const int o = __shfl_up_sync (0xffffffffu, up, d);         // earlier lanes
		if (lane >= d) up = clf_comb (o, up);
		const int p = __shfl_down_sync (0xffffffffu, dn, d);       // lanes above come first going down
		if (lane + d < 32) dn = clf_comb (p, dn);
		}
	const int upAll = __shfl_sync (0xffffffffu, up, 31), dnAll = __shfl_sync (0xffffffffu, dn, 0);
	if (lane == 0) wk.tsum[tile] = (unsigned char) (upAll | (dnAll << 2));
	}

// one block per chromosome: the carry entering every tile from below and from above
__global__ void __launch_bounds__(256)
k_clump_tilecarry (const uint64_t* __restrict__ base, ClumpFast wk, const unsigned char* __restrict__ segCin)
	{
	__shared__ int s_up[256], s_dn[256];
	uint64_t b0 = base[blockIdx.x];
	const uint64_t b1 = base[blockIdx.x + 1];
	if (wk.segFlags != NULL && (wk.segFlags[blockIdx.x] & 1u) && b0 < b1)
		{
		if (threadIdx.x == 0) wk.tcin[b0] = 0;                     // halo tile: not emitted
		b0++;
		}
	const int cin0 = (segCin != NULL) ? segCin[blockIdx.x] : 0;   // slab piece: bit0 from the left neighbour, bit1 from the right
	const uint64_t nt = b1 - b0, chunk = (nt + 255) / 256;
	const uint64_t lo = b0 + threadIdx.x * chunk < b1 ? b0 + threadIdx.x * chunk : b1;
	const uint64_t hi = lo + chunk < b1 ? lo + chunk : b1;
	int up = 2, dn = 2;                                            // identity: propagate only
	for (uint64_t t = lo; t < hi; t++) up = clf_comb (up, wk.tsum[t] & 3);
	for (uint64_t t = hi; t > lo; t--) dn = clf_comb (dn, (wk.tsum[t - 1] >> 2) & 3);
	s_up[threadIdx.x] = up;  s_dn[threadIdx.x] = dn;
	__syncthreads ();
	int cu = cin0 & 1, cd = (cin0 >> 1) & 1;                       // nothing enters a whole chromosome from outside
	for (int k = 0; k < (int) threadIdx.x; k++) cu = clf_apply (s_up[k], cu);
	for (int k = 255; k > (int) threadIdx.x; k--) cd = clf_apply (s_dn[k], cd);
	for (uint64_t t = lo; t < hi; t++)
		{
		wk.tcin[t] = (unsigned char) cu;                           // bit1 is OR-ed in below
		cu = clf_apply (wk.tsum[t] & 3, cu);
		}
	for (uint64_t t = hi; t > lo; t--)
		{
		wk.tcin[t - 1] |= (unsigned char) (cd << 1);
		cd = clf_apply ((wk.tsum[t - 1] >> 2) & 3, cd);
		}
	}

// one block per tile: the words of the trimmed runs, then the one/zero cells
__global__ void __launch_bounds__(CL_THREADS, 6)
k_clump_emit (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
              double* __restrict__ sig, double oneVal, double zeroVal, ClumpFast wk)
	{
	__shared__ uint32_t s_out[CLF_WORDS];
	__shared__ int s_wu[4], s_wd[4];
	const uint64_t tile = blockIdx.x;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * CL_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < CL_TILE) ? (sd.hi - t0) : CL_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool allNeg = wk.segAllNeg[seg] != 0;                   // clump.c:545-565: nothing can clump
	const int tc = wk.tcin[tile];
	if (wk.segFlags != NULL && (wk.segFlags[seg] & 1u) && tis == 0) return;     // the neighbour's cells

	uint32_t M = 0, S = 0;
	int up = 2, dn = 2, upIn = 2, dnIn = 2;
	if (warp < 4)
		{
		M = wk.Bm[tile * CLF_WORDS + threadIdx.x];
		S = wk.Bq[tile * CLF_WORDS + threadIdx.x] & M;
		up = clf_gp (M, S);  dn = clf_gp (__brev (M), __brev (S));
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			const int o = __shfl_up_sync (0xffffffffu, up, d);
			if (lane >= d) up = clf_comb (o, up);
			const int p = __shfl_down_sync (0xffffffffu, dn, d);
			if (lane + d < 32) dn = clf_comb (p, dn);
			}
		upIn = __shfl_up_sync (0xffffffffu, up, 1);                // fold of the earlier words of this warp
		if (lane == 0) upIn = 2;
		dnIn = __shfl_down_sync (0xffffffffu, dn, 1);
		if (lane == 31) dnIn = 2;
		if (lane == 31) s_wu[warp] = up;
		if (lane == 0)  s_wd[warp] = dn;
		}
	__syncthreads ();
	if (warp < 4)
		{
		int cu = tc & 1, cd = (tc >> 1) & 1;
		for (int w = 0; w < warp; w++) cu = clf_apply (s_wu[w], cu);
		for (int w = 3; w > warp; w--) cd = clf_apply (s_wd[w], cd);
		cu = clf_apply (upIn, cu);  cd = clf_apply (dnIn, cd);
		const uint32_t fu = clf_fill_up (M, S, (uint32_t) cu);
		const uint32_t fd = __brev (clf_fill_up (__brev (M), __brev (S), (uint32_t) cd));
		s_out[threadIdx.x] = allNeg ? 0u : (fu & fd);
		}
	__syncthreads ();

	#pragma unroll
	for (int r = 0; r < CL_ROWS; r++)
		{
		const uint32_t e0 = cl_elem (warp, lane, r, 0);
		if (e0 >= n) continue;
		const uint32_t nib = s_out[e0 >> 5] >> (e0 & 31);
		double y[4];
		#pragma unroll
		for (int c = 0; c < 4; c++) y[c] = ((nib >> c) & 1) ? oneVal : zeroVal;
		if (e0 + 4 <= n) stg_stream4 (sig + t0 + e0, y[0], y[1], y[2], y[3]);
		else
			for (int c = 0; c < 4 && e0 + c < n; c++) sig[t0 + e0 + c] = y[c];
		}
	}

__global__ void k_fill_int (int* p, int n, int v)
	{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
	}

// ===========================================================================
// Slab-sharded chromosomes (one GPU owns a contiguous piece): the same kernels, with the four
// chromosome-wide dependencies of the search exchanged as per-piece CARRIES instead of moving the
// signal (BASELINE north_star):
//   forward   {sum of d, minimum prefix sum} of every piece  -> {P, M} entering the next piece
//   backward  maximum valid P of every piece                 -> suffix maximum entering the previous piece
//   both ways generate/propagate pair of the run trimming    -> carry bits entering from either side
// Slab cuts are multiples of CL_TILE in chromosome coordinates, so a piece's tiles and 512-cell groups
// coincide with the whole chromosome's.  A piece that continues on the left starts one tile early
// ("halo tile": the last CL_TILE cells of the left neighbour, supplied by the halo exchange): its groups
// are reduced and scanned like the piece's own -- they are the Lmin-cell look-back of the first owned
// tile -- but it is neither marked nor emitted.  A piece therefore publishes its aggregate in two parts:
// head (all but its last tile) and tail (the last tile = the right neighbour's halo tile), so that the
// neighbour can start its sums at the beginning of its halo tile.
// ===========================================================================
#define CLS_HALO   1u        // the first tile of the segment holds the left neighbour's cells
#define CLS_CONT_R 2u        // the chromosome continues on the right neighbour

struct AggSM { double S, m; };
__device__ __forceinline__ AggSM agg_comb (AggSM a, AggSM b) { AggSM r;  r.S = a.S + b.S;  r.m = dmin2 (a.m, a.S + b.m);  return r; }

// one block per segment: {sum, minimum prefix sum} of the head groups and of the tail groups, out[seg*4..]
__global__ void __launch_bounds__(1024)
k_clump_groupagg (const uint64_t* __restrict__ base, ClumpFast wk, double* __restrict__ out)
	{
	__shared__ AggSM s_w[32];
	const unsigned fl = wk.segFlags[blockIdx.x];
	const uint64_t g0 = (base[blockIdx.x] + (fl & CLS_HALO)) * CLF_GROUPS, g1 = base[blockIdx.x + 1] * CLF_GROUPS;
	const uint64_t gT = ((fl & CLS_CONT_R) && g1 >= g0 + CLF_GROUPS) ? g1 - CLF_GROUPS : g1;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int part = 0; part < 2; part++)
		{
		const uint64_t a = part ? gT : g0, b = part ? g1 : gT;
		const uint64_t chunk = (b - a + 1023) / 1024;
		const uint64_t lo = (a + threadIdx.x * chunk < b) ? a + threadIdx.x * chunk : b;
		const uint64_t hi = (lo + chunk < b) ? lo + chunk : b;
		AggSM v;  v.S = 0.0;  v.m = CLF_INF;
		for (uint64_t g = lo; g < hi; g++)
			{
			const double2 r = wk.carry[g];
			AggSM x;  x.S = r.x;  x.m = r.y;
			v = agg_comb (v, x);
			}
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)                           // ordered fold: lane l+d is later
			{
			AggSM o;  o.S = shfl_down_f64 (v.S, d);  o.m = shfl_down_f64 (v.m, d);
			if (lane + d < 32 && ((lane & (2 * d - 1)) == 0)) v = agg_comb (v, o);
			}
		__syncthreads ();
		if (lane == 0) s_w[warp] = v;
		__syncthreads ();
		if (threadIdx.x == 0)
			{
			AggSM t = s_w[0];
			for (int w = 1; w < 32; w++) t = agg_comb (t, s_w[w]);
			out[blockIdx.x * 4 + part * 2] = t.S;  out[blockIdx.x * 4 + part * 2 + 1] = t.m;
			}
		}
	}

// one block per segment: cells p of the owned part with M[p-1] <= sufIn are marked because a valid end
// with that prefix sum lies in a piece further right.  M is non-increasing, so they form a suffix:
// whole groups from the first group whose carry M is <= sufIn, plus a tail of the group before it.
template <bool ABOVE>
__global__ void __launch_bounds__(1024)
k_clump_fixup (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, const double* __restrict__ sig, double T,
               ClumpFast wk, const double* __restrict__ sufIn)
	{
	const double sIn = sufIn[blockIdx.x];
	if (!(sIn > -CLF_INF)) return;
	const unsigned fl = wk.segFlags[blockIdx.x];
	const uint64_t gs = base[blockIdx.x] * CLF_GROUPS;             // first group of the (extended) segment
	const uint64_t g0 = gs + (fl & CLS_HALO) * CLF_GROUPS, g1 = base[blockIdx.x + 1] * CLF_GROUPS;
	uint64_t lo = g0, hi = g1;                                     // smallest g in [g0,g1) with carry[g].y <= sIn
	while (lo < hi)
		{
		const uint64_t mid = (lo + hi) >> 1;
		if (wk.carry[mid].y <= sIn) hi = mid; else lo = mid + 1;
		}
	const uint64_t gstar = lo;
	for (uint64_t w = gstar * (CLF_GROUP / 32) + threadIdx.x; w < g1 * (CLF_GROUP / 32); w += 1024) wk.Bm[w] = 0xffffffffu;
	if (gstar > g0 && threadIdx.x < 32)
		{
		const int lane = threadIdx.x;
		const uint64_t g = gstar - 1;
		const SegDev sd = segs[blockIdx.x];
		double P[16], tot;
		clf_load16<true, ABOVE> (sig, sd.lo + (g - gs) * CLF_GROUP + lane * 16, 16, T, P);
		group_sum_scan (P, tot);
		const double2 cr = wk.carry[g];
		double m = CLF_INF;
		#pragma unroll
		for (int k = 0; k < 16; k++) { P[k] = cr.x + P[k];  m = dmin2 (m, P[k]); }
		double inc = m;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			const double up = shfl_up_f64 (inc, d);
			if (lane >= d) inc = dmin2 (up, inc);
			}
		double ex = shfl_up_f64 (inc, 1);
		if (lane == 0) ex = CLF_INF;
		double run = dmin2 (ex, cr.y);                             // M[p-1] of this lane's first cell
		unsigned mk = 0;
		#pragma unroll
		for (int k = 0; k < 16; k++)
			{
			if (run <= sIn) mk |= 1u << k;
			run = dmin2 (run, P[k]);
			}
		unsigned wm = mk << ((lane & 1) * 16);
		wm |= __shfl_xor_sync (0xffffffffu, wm, 1);
		if ((lane & 1) == 0 && wm) wk.Bm[g * (CLF_GROUP / 32) + (lane >> 1)] |= wm;
		}
	}

// one block per segment: generate/propagate pair of the owned tiles, both directions (bits 0-1 up, 2-3 down)
__global__ void __launch_bounds__(256)
k_clump_seggp (const uint64_t* __restrict__ base, ClumpFast wk, int* __restrict__ out)
	{
	__shared__ int s_up[256], s_dn[256];
	const uint64_t b0 = base[blockIdx.x] + (wk.segFlags[blockIdx.x] & CLS_HALO), b1 = base[blockIdx.x + 1];
	const uint64_t nt = b1 - b0, chunk = (nt + 255) / 256;
	const uint64_t lo = b0 + threadIdx.x * chunk < b1 ? b0 + threadIdx.x * chunk : b1;
	const uint64_t hi = lo + chunk < b1 ? lo + chunk : b1;
	int up = 2, dn = 2;
	for (uint64_t t = lo; t < hi; t++) up = clf_comb (up, wk.tsum[t] & 3);
	for (uint64_t t = hi; t > lo; t--) dn = clf_comb (dn, (wk.tsum[t - 1] >> 2) & 3);
	s_up[threadIdx.x] = up;  s_dn[threadIdx.x] = dn;
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		int u = 2, d = 2;
		for (int k = 0; k < 256; k++) u = clf_comb (u, s_up[k]);
		for (int k = 255; k >= 0; k--) d = clf_comb (d, s_dn[k]);
		out[blockIdx.x] = u | (d << 2);
		}
	}

struct gdsp_clump_slab
	{
	gdsp_ctx*    c;
	gdsp_layout* E;            // the segments, each extended by its halo tile
	int          nseg;
	double       average, relLength, oneVal, zeroVal;
	uint32_t     minLength;
	int          above;
	ClumpFast    wf;
	TileMap      tm;
	size_t       smem;
	void*        ws1;
	std::vector<unsigned char> flags;
	char*        dev;          // one allocation holding the per-segment arrays below
	double2*       d_carryIn;
	double*        d_sufMax;
	double*        d_sufIn;
	double*        d_agg;
	int*           d_gp;
	unsigned char* d_flags;
	unsigned char* d_cin;
	};

extern "C" void gdsp_clump_slab_destroy (gdsp_clump_slab* cs)
	{
	if (cs == NULL) return;
	if (cs->E) gdsp_layout_destroy (cs->E);
	if (cs->dev) cudaFree (cs->dev);
	delete cs;
	}

extern "C" int gdsp_clump_slab_create (gdsp_ctx* c, const gdsp_layout* L_, uint64_t buffer_cells, void* work,
                                       double average, uint32_t minLength, double relLength, int above,
                                       double oneVal, double zeroVal, gdsp_clump_slab** out)
	{
	const gdsp_layout* L = L_;
	GDSP_REQUIRE (c && L && work && out, "gdsp_clump_slab_create: NULL argument");
	GDSP_REQUIRE_ALIGNED (work, "gdsp_clump_slab_create");
	GDSP_REQUIRE (L->nseg <= 65536, "gdsp_clump_slab_create: more than 65536 segments");
	uint32_t maxLmin = minLength;
	std::vector<gdsp_seg> ext (L->nseg);
	std::vector<unsigned char> flags (L->nseg);
	for (int s = 0; s < L->nseg; s++)
		{
		gdsp_seg g = L->h[s];
		if (relLength > 0.0)
			{
			const uint32_t rl = (uint32_t) (relLength * g.chrom_len);
			if (rl > maxLmin) maxLmin = rl;
			}
		unsigned fl = 0;
		if (g.pos0 > 0)
			{
			GDSP_REQUIRE (g.pos0 % CL_TILE == 0, "gdsp_clump_slab: a slab cut must be a multiple of 4096 in chromosome coordinates");
			GDSP_REQUIRE (g.lo - g.dlo >= CL_TILE, "gdsp_clump_slab: needs 4096 halo cells on the left of a cut chromosome");
			g.lo -= CL_TILE;  g.pos0 -= CL_TILE;  fl |= CLS_HALO;
			}
		const uint64_t endPos = (uint64_t) L->h[s].pos0 + (L->h[s].hi - L->h[s].lo);
		if (endPos < g.chrom_len)
			{
			GDSP_REQUIRE (endPos % CL_TILE == 0 && L->h[s].hi - L->h[s].lo >= CL_TILE,
			              "gdsp_clump_slab: a slab cut must be a multiple of 4096 in chromosome coordinates (and a piece at least 4096 cells)");
			fl |= CLS_CONT_R;
			}
		ext[s] = g;  flags[s] = (unsigned char) fl;
		}
	GDSP_REQUIRE (maxLmin <= CLF_MAX_HALO, "gdsp_clump_slab: minimum lengths above 4096 are not supported on slab-sharded chromosomes");
	gdsp_clump_slab* cs = new gdsp_clump_slab ();
	cs->c = c;  cs->E = NULL;  cs->dev = NULL;  cs->nseg = L->nseg;  cs->flags = flags;
	cs->average = average;  cs->relLength = relLength;  cs->oneVal = oneVal;  cs->zeroVal = zeroVal;
	cs->minLength = minLength;  cs->above = above;
	int st = gdsp_layout_create (c, ext.data (), L->nseg, &cs->E);
	if (st != GDSP_OK) { delete cs;  return st; }
	st = gdsp_layout_tilemap (cs->E, CL_TILE, &cs->tm);
	if (st != GDSP_OK) { gdsp_clump_slab_destroy (cs);  return st; }
	const uint64_t ntiles = cs->tm.ntiles;
	const size_t fastBytes = (size_t) ntiles * (CLF_GROUPS * 16 + 2 * CLF_WORDS * 4 + CLF_GROUPS * 8) + 256 + 3 * (((size_t) ntiles + 255) / 256) * 256 + (size_t) L->nseg * 4;
	if (fastBytes > gdsp_clump_work_bytes (buffer_cells))
		{
		gdsp_clump_slab_destroy (cs);
		gdsp_set_error ("gdsp_clump_slab_create: work buffer too small");
		return GDSP_ERR_ARG;
		}
	const size_t n = (size_t) L->nseg;
	const size_t bytes = n * (16 + 8 + 8 + 32 + 4 + 1 + 1) + 256;
	if (cudaMalloc (&cs->dev, bytes) != cudaSuccess)
		{
		gdsp_clump_slab_destroy (cs);
		gdsp_set_error ("gdsp_clump_slab_create: out of device memory");
		return GDSP_ERR_NOMEM;
		}
	char* q = cs->dev;
	cs->d_carryIn = (double2*) q;  q += n * 16;
	cs->d_agg     = (double*) q;   q += n * 32;
	cs->d_sufMax  = (double*) q;   q += n * 8;
	cs->d_sufIn   = (double*) q;   q += n * 8;
	cs->d_gp      = (int*) q;      q += n * 4;
	cs->d_flags   = (unsigned char*) q;  q += n;
	cs->d_cin     = (unsigned char*) q;
	cudaMemcpyAsync (cs->d_flags, flags.data (), n, cudaMemcpyHostToDevice, c->stream);
	cudaStreamSynchronize (c->stream);
	ClumpFast& wf = cs->wf;
	char* p = (char*) work;
	wf.carry = (double2*) p;            p += (size_t) ntiles * CLF_GROUPS * 16;
	wf.Bm = (uint32_t*) p;              p += (size_t) ntiles * CLF_WORDS * 4;
	wf.Bq = (uint32_t*) p;              p += (size_t) ntiles * CLF_WORDS * 4;
	wf.tsum = (unsigned char*) p;       p += (((size_t) ntiles + 255) / 256) * 256;
	wf.tcin = (unsigned char*) p;       p += (((size_t) ntiles + 255) / 256) * 256;
	wf.tquiet = (unsigned char*) p;     p += (((size_t) ntiles + 255) / 256) * 256;
	wf.gmid = (double*) p;              p += (((size_t) ntiles * CLF_GROUPS * 8 + 255) / 256) * 256;
	wf.segAllNeg = (int*) p;
	wf.segCarryIn = cs->d_carryIn;  wf.segSufMax = cs->d_sufMax;  wf.segFlags = cs->d_flags;
	const uint32_t reach = maxLmin ? maxLmin : 1;
	cs->smem = ((size_t) ((reach + CLF_GROUP - 1) / CLF_GROUP) + CLF_GROUPS) * (CLF_GROUP / 16 * 17) * sizeof (double);
	if (cs->smem > 48 * 1024)
		{
		cudaFuncSetAttribute (k_clump_mark<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) cs->smem);
		cudaFuncSetAttribute (k_clump_mark<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) cs->smem);
		}
	cs->ws1 = NULL;
	*out = cs;
	return GDSP_OK;
	}

// phase 1: group aggregates.  h_agg[s*5..] = {head sum, head minimum prefix, tail sum, tail minimum prefix, all-negative flag}
extern "C" int gdsp_clump_slab_reduce (gdsp_clump_slab* cs, const double* sig, double* h_agg)
	{
	GDSP_REQUIRE (cs && sig && h_agg, "gdsp_clump_slab_reduce: NULL argument");
	gdsp_ctx* c = cs->c;  gdsp_layout* E = cs->E;
	k_fill_int<<<(cs->nseg + 255) / 256, 256, 0, c->stream>>> (cs->wf.segAllNeg, cs->nseg, 1);
	GDSP_KERNEL_CHECK ();
	if (cs->above) k_clump_groups<true><<<(unsigned) cs->tm.ntiles, CL_THREADS, 0, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, sig, cs->average, cs->wf);
	else           k_clump_groups<false><<<(unsigned) cs->tm.ntiles, CL_THREADS, 0, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, sig, cs->average, cs->wf);
	GDSP_KERNEL_CHECK ();
	k_clump_groupagg<<<cs->nseg, 1024, 0, c->stream>>> (cs->tm.d_base, cs->wf, cs->d_agg);
	GDSP_KERNEL_CHECK ();
	std::vector<double> agg ((size_t) cs->nseg * 4);
	std::vector<int> neg (cs->nseg);
	GDSP_CUDA (cudaMemcpyAsync (agg.data (), cs->d_agg, sizeof (double) * 4 * cs->nseg, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (neg.data (), cs->wf.segAllNeg, sizeof (int) * cs->nseg, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	for (int s = 0; s < cs->nseg; s++)
		{
		for (int k = 0; k < 4; k++) h_agg[s * 5 + k] = agg[(size_t) s * 4 + k];
		h_agg[s * 5 + 4] = neg[s] ? 1.0 : 0.0;
		}
	return GDSP_OK;
	}

// phase 2: h_carry_in[s*2..] = {P, M} just before the first cell of segment s's halo tile (or of the segment
// when it starts the chromosome: {0, 0}); marks; h_sufmax[s] = the maximum valid P of the owned cells
extern "C" int gdsp_clump_slab_mark (gdsp_clump_slab* cs, const double* sig, const double* h_carry_in, double* h_sufmax)
	{
	GDSP_REQUIRE (cs && sig && h_carry_in && h_sufmax, "gdsp_clump_slab_mark: NULL argument");
	gdsp_ctx* c = cs->c;  gdsp_layout* E = cs->E;
	GDSP_TRY (gdsp_ws (c, 0, scan_status_bytes<double> (cs->tm.ntiles), &cs->ws1));
	GDSP_CUDA (cudaMemcpyAsync (cs->d_carryIn, h_carry_in, sizeof (double) * 2 * cs->nseg, cudaMemcpyHostToDevice, c->stream));
	k_clump_groupscan<<<cs->nseg, CLF_GS_THREADS, 0, c->stream>>> (cs->tm.d_base, cs->wf);
	GDSP_KERNEL_CHECK ();
	k_clump_classify<<<(unsigned) ((cs->tm.ntiles + 255) / 256), 256, 0, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, cs->tm.ntiles, cs->minLength, cs->relLength, cs->wf);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemsetAsync (cs->ws1, 0, scan_status_clear_bytes<double> (cs->tm.ntiles), c->stream));
	if (cs->above) k_clump_mark<true><<<(unsigned) cs->tm.ntiles, CL_THREADS, cs->smem, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, cs->tm.ntiles, sig,
	        cs->average, cs->minLength, cs->relLength, cs->wf, scan_status_carve<double> (cs->ws1, cs->tm.ntiles));
	else           k_clump_mark<false><<<(unsigned) cs->tm.ntiles, CL_THREADS, cs->smem, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, cs->tm.ntiles, sig,
	        cs->average, cs->minLength, cs->relLength, cs->wf, scan_status_carve<double> (cs->ws1, cs->tm.ntiles));
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemcpyAsync (h_sufmax, cs->d_sufMax, sizeof (double) * cs->nseg, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

// phase 3: h_sufmax_in[s] = maximum valid P over the pieces to the right of segment s (-inf if none); the cells
// that maximum reaches are marked, the tile summaries of the trimming are built; h_gp[s] = generate/propagate
// pair of the owned tiles (bits 0-1 going up, 2-3 going down)
extern "C" int gdsp_clump_slab_trim (gdsp_clump_slab* cs, const double* sig, const double* h_sufmax_in, int* h_gp)
	{
	GDSP_REQUIRE (cs && sig && h_sufmax_in && h_gp, "gdsp_clump_slab_trim: NULL argument");
	gdsp_ctx* c = cs->c;  gdsp_layout* E = cs->E;
	GDSP_CUDA (cudaMemcpyAsync (cs->d_sufIn, h_sufmax_in, sizeof (double) * cs->nseg, cudaMemcpyHostToDevice, c->stream));
	if (cs->above) k_clump_fixup<true><<<cs->nseg, 1024, 0, c->stream>>> (E->d, cs->tm.d_base, sig, cs->average, cs->wf, cs->d_sufIn);
	else           k_clump_fixup<false><<<cs->nseg, 1024, 0, c->stream>>> (E->d, cs->tm.d_base, sig, cs->average, cs->wf, cs->d_sufIn);
	GDSP_KERNEL_CHECK ();
	k_clump_tilesum<<<(unsigned) ((cs->tm.ntiles + 7) / 8), 256, 0, c->stream>>> (cs->tm.ntiles, cs->wf);
	GDSP_KERNEL_CHECK ();
	k_clump_seggp<<<cs->nseg, 256, 0, c->stream>>> (cs->tm.d_base, cs->wf, cs->d_gp);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemcpyAsync (h_gp, cs->d_gp, sizeof (int) * cs->nseg, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

// phase 4: h_cin[s] bit0 = a trimmed run reaches segment s from the left neighbour, bit1 = from the right;
// h_allneg[s] = every d < 0 on the WHOLE chromosome (clump.c:545-565); writes one/zero into the owned cells
extern "C" int gdsp_clump_slab_emit (gdsp_clump_slab* cs, double* sig, const unsigned char* h_cin, const int* h_allneg)
	{
	GDSP_REQUIRE (cs && sig && h_cin && h_allneg, "gdsp_clump_slab_emit: NULL argument");
	gdsp_ctx* c = cs->c;  gdsp_layout* E = cs->E;
	GDSP_CUDA (cudaMemcpyAsync (cs->d_cin, h_cin, cs->nseg, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (cs->wf.segAllNeg, h_allneg, sizeof (int) * cs->nseg, cudaMemcpyHostToDevice, c->stream));
	k_clump_tilecarry<<<cs->nseg, 256, 0, c->stream>>> (cs->tm.d_base, cs->wf, cs->d_cin);
	GDSP_KERNEL_CHECK ();
	k_clump_emit<<<(unsigned) cs->tm.ntiles, CL_THREADS, 0, c->stream>>> (E->d, cs->tm.d_base, cs->nseg, sig, cs->oneVal, cs->zeroVal, cs->wf);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}



extern "C" size_t gdsp_clump_work_bytes (uint64_t buffer_cells)
	{
	return (size_t) (buffer_cells * 17 + 3 * 256 + 65536 * 4);
	}

extern "C" int gdsp_clump (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint64_t buffer_cells, void* work,
                           double average, uint32_t minLength, double relLength, int above,
                           double oneVal, double zeroVal)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && work, "gdsp_clump: NULL argument");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_clump");  GDSP_REQUIRE_ALIGNED (work, "gdsp_clump");
	GDSP_REQUIRE (L->nseg <= 65536, "gdsp_clump: more than 65536 segments");
	for (int s = 0; s < L->nseg; s++)
		GDSP_REQUIRE (L->h[s].pos0 == 0 && L->h[s].hi - L->h[s].lo == L->h[s].chrom_len,
		              "gdsp_clump: a slab-sharded chromosome needs the carry variant (gdsp_clump_slab_*)");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, CL_TILE, &tm));

	// two scan status blocks (reused by the passes)
	size_t sb = scan_status_bytes<double> (tm.ntiles);
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, 2 * sb, &ws));
	void* ws1 = ws;  void* ws2 = (char*) ws + sb;

	// the longest minimum length of any chromosome decides the path (clump.c:516-522)
	uint32_t maxLmin = minLength;
	if (relLength > 0.0)
		for (int s = 0; s < L->nseg; s++)
			{
			const uint32_t rl = (uint32_t) (relLength * L->h[s].chrom_len);
			if (rl > maxLmin) maxLmin = rl;
			}
	const size_t fastBytes = (size_t) tm.ntiles * (CLF_GROUPS * 16 + 2 * CLF_WORDS * 4 + CLF_GROUPS * 8) + 256 + 3 * (((size_t) tm.ntiles + 255) / 256) * 256
	                       + (size_t) L->nseg * 4;
	if (maxLmin <= CLF_MAX_HALO && fastBytes <= gdsp_clump_work_bytes (buffer_cells) && !getenv ("GDSP_CLUMP_STORED") && !c->exact_order)
		{
		ClumpFast wf;
		char* p = (char*) work;
		wf.carry = (double2*) p;            p += (size_t) tm.ntiles * CLF_GROUPS * 16;
		wf.Bm = (uint32_t*) p;              p += (size_t) tm.ntiles * CLF_WORDS * 4;
		wf.Bq = (uint32_t*) p;              p += (size_t) tm.ntiles * CLF_WORDS * 4;
		wf.tsum = (unsigned char*) p;       p += (((size_t) tm.ntiles + 255) / 256) * 256;
		wf.tcin = (unsigned char*) p;       p += (((size_t) tm.ntiles + 255) / 256) * 256;
		wf.tquiet = (unsigned char*) p;     p += (((size_t) tm.ntiles + 255) / 256) * 256;
		wf.gmid = (double*) p;              p += (((size_t) tm.ntiles * CLF_GROUPS * 8 + 255) / 256) * 256;
		wf.segAllNeg = (int*) p;
		wf.segCarryIn = NULL;  wf.segSufMax = NULL;  wf.segFlags = NULL;
		const uint32_t reach = maxLmin ? maxLmin : 1;
		const size_t smem = ((size_t) ((reach + CLF_GROUP - 1) / CLF_GROUP) + CLF_GROUPS) * (CLF_GROUP / 16 * 17) * sizeof (double);
		if (smem > 48 * 1024)
			{
			GDSP_CUDA (cudaFuncSetAttribute (k_clump_mark<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			GDSP_CUDA (cudaFuncSetAttribute (k_clump_mark<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			}

		k_fill_int<<<(L->nseg + 255) / 256, 256, 0, c->stream>>> (wf.segAllNeg, L->nseg, 1);
		GDSP_KERNEL_CHECK ();
		if (above) k_clump_groups<true><<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, average, wf);
		else       k_clump_groups<false><<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, average, wf);
		GDSP_KERNEL_CHECK ();
		k_clump_groupscan<<<L->nseg, CLF_GS_THREADS, 0, c->stream>>> (tm.d_base, wf);
		GDSP_KERNEL_CHECK ();
		k_clump_classify<<<(unsigned) ((tm.ntiles + 255) / 256), 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, minLength, relLength, wf);
		GDSP_KERNEL_CHECK ();
		GDSP_CUDA (cudaMemsetAsync (ws1, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
		if (above) k_clump_mark<true><<<(unsigned) tm.ntiles, CL_THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, average,
		        minLength, relLength, wf, scan_status_carve<double> (ws1, tm.ntiles));
		else       k_clump_mark<false><<<(unsigned) tm.ntiles, CL_THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, average,
		        minLength, relLength, wf, scan_status_carve<double> (ws1, tm.ntiles));
		GDSP_KERNEL_CHECK ();
		k_clump_tilesum<<<(unsigned) ((tm.ntiles + 7) / 8), 256, 0, c->stream>>> (tm.ntiles, wf);
		GDSP_KERNEL_CHECK ();
		k_clump_tilecarry<<<L->nseg, 256, 0, c->stream>>> (tm.d_base, wf, NULL);
		GDSP_KERNEL_CHECK ();
		k_clump_emit<<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, oneVal, zeroVal, wf);
		GDSP_KERNEL_CHECK ();
		return GDSP_OK;
		}

	// stored-prefix passes (minimum lengths beyond CLF_MAX_HALO)
	ClumpWork wk;
	char* p = (char*) work;
	wk.P = (double*) p;                 p += ((buffer_cells * 8 + 255) / 256) * 256;
	wk.M = (double*) p;                 p += ((buffer_cells * 8 + 255) / 256) * 256;
	wk.F = (unsigned char*) p;          p += ((buffer_cells + 255) / 256) * 256;
	wk.segAllNeg = (int*) p;

	k_fill_int<<<(L->nseg + 255) / 256, 256, 0, c->stream>>> (wk.segAllNeg, L->nseg, 1);
	GDSP_KERNEL_CHECK ();

	if (c->exact_order)
		{
		// the prefix sums in the reference's own order (clump.c:600): one warp per chromosome (gdsp_exact.cu)
		GDSP_TRY (gdsp_clump_prefix_exact (c, L, sig, average, above ? 1 : 0, wk.P, wk.M, wk.segAllNeg));
		}
	else
		{
		// pass A: sum scan (ws1) + min scan (ws2); the ticket of ws1 drives the tile order
		GDSP_CUDA (cudaMemsetAsync (ws1, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
		GDSP_CUDA (cudaMemsetAsync (ws2, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
		k_clump_a<<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, average, above ? 1 : 0, wk,
		        scan_status_carve<double> (ws1, tm.ntiles), scan_status_carve<double> (ws2, tm.ntiles));
		GDSP_KERNEL_CHECK ();
		}

	// pass B: suffix max (ws1) + backward segmented OR (ws2)
	GDSP_CUDA (cudaMemsetAsync (ws1, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	GDSP_CUDA (cudaMemsetAsync (ws2, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	k_clump_b<<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, average, above ? 1 : 0,
	        minLength, relLength, wk, scan_status_carve<double> (ws1, tm.ntiles), scan_status_carve<int> (ws2, tm.ntiles));
	GDSP_KERNEL_CHECK ();

	// pass C: forward segmented OR (ws1)
	GDSP_CUDA (cudaMemsetAsync (ws1, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	k_clump_c<<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, oneVal, zeroVal, wk,
	        scan_status_carve<int> (ws1, tm.ntiles));
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
