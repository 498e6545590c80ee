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

__global__ void k_fill_int (int* p, int n, int v)
	{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
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
		              "gdsp_clump: slab-sharded chromosomes need the carry variant (not in this build)");
	ClumpWork wk;
	char* p = (char*) work;
	wk.P = (double*) p;                 p += ((buffer_cells * 8 + 255) / 256) * 256;
	wk.M = (double*) p;                 p += ((buffer_cells * 8 + 255) / 256) * 256;
	wk.F = (unsigned char*) p;          p += ((buffer_cells + 255) / 256) * 256;
	wk.segAllNeg = (int*) p;
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, CL_TILE, &tm));

	// two scan status blocks (reused by the three passes)
	size_t sb = scan_status_bytes<double> (tm.ntiles);
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, 2 * sb, &ws));
	void* ws1 = ws;  void* ws2 = (char*) ws + sb;

	k_fill_int<<<(L->nseg + 255) / 256, 256, 0, c->stream>>> (wk.segAllNeg, L->nseg, 1);
	GDSP_KERNEL_CHECK ();

	// pass A: sum scan (ws1) + min scan (ws2); the ticket of ws1 drives the tile order
	GDSP_CUDA (cudaMemsetAsync (ws1, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	GDSP_CUDA (cudaMemsetAsync (ws2, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	k_clump_a<<<(unsigned) tm.ntiles, CL_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, average, above ? 1 : 0, wk,
	        scan_status_carve<double> (ws1, tm.ntiles), scan_status_carve<double> (ws2, tm.ntiles));
	GDSP_KERNEL_CHECK ();

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
