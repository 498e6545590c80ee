// gdsp_accumulate.cu -- interval accumulation (difference array + segmented scan)
// and cumulative sum.
//
// Replaces the accumulate loops of read_intervals (genodsp.c:1307-1330, sum
// overlap) and op_cumulative_sum_apply (sum.c:776-792).
//
// Unit-weight intervals (the `--novalue` depth case, the benchmark path):
//   k_bin_count / k_bin_offsets / k_bin_scatter / k_bin_final -- every interval becomes one 32-bit
//   record in the bucket of the 8192-cell tile it starts in; one block per tile turns its records
//   into depth with shared-memory counters and writes fp64 once (no difference array in HBM).
// Intervals with integer or dyadic values (`add`, `subtract`, valued `input`):
//   K1  k_diff_*      one thread per interval: +w at the interval start, -w at
//                     its end (clipped to the owned piece of the chromosome)
//   K2  k_scan_tiles  segmented inclusive prefix sum of the difference array (int32 or fp64),
//                     chain-free: k_diff also maintains the tile sums
// (other real values are applied exactly, in file order, by the host's layered pointwise launches:
// genodsp_b200/host/gd_device.c, gd_apply_intervals_exact)
// Cumulative sum: k_cumsum, a chained (decoupled look-back) two-phase tile scan.
//
// Algorithmic bytes: 16 B/bp + 28 B/interval (binned and int32 paths), 24 B/bp +
// 52 B/interval (fp64 difference array), 16 B/bp (cumulative sum).
#include "gdsp_common.cuh"
#include "gdsp_scan.cuh"

#define SCAN_THREADS 256
#define SCAN_WARPS   (SCAN_THREADS / 32)
#define SCAN_ROWS    4
#define SCAN_TILE    (SCAN_WARPS * SCAN_ROWS * 128)      // 4096 cells

template <typename T> __device__ __forceinline__ T warp_shfl_up (T v, int d);
template <typename T> __device__ __forceinline__ T warp_shfl_idx (T v, int s);

// ---------------------------------------------------------------------------
// K1: difference array
// ---------------------------------------------------------------------------

// Besides the difference array, every interval also updates the SUM of its tile of the difference
// array (tileSum, a few MB that stay in L2): the prefix over those sums gives every tile of the scan
// its starting value up front, so the scan below needs no inter-block chain at all.
template <typename T>
__global__ void __launch_bounds__(256)
k_diff (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, T* __restrict__ diff,
        T* __restrict__ tileSum,
        const uint32_t* __restrict__ iseg, const uint32_t* __restrict__ istart,
        const uint32_t* __restrict__ iend, const double* __restrict__ ival, uint64_t n)
	{
	uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	uint32_t sg = iseg[k];
	if (sg >= (uint32_t) nseg) return;
	uint32_t s = istart[k], e = iend[k];
	if (s >= e) return;                              // empty interval: the reference loop runs zero times
	const SegDev sd = segs[sg];
	uint64_t len = sd.hi - sd.lo;
	uint64_t p0 = sd.pos0, p1 = p0 + len;
	if ((uint64_t) e <= p0 || (uint64_t) s >= p1) return;
	uint64_t a = ((uint64_t) s > p0 ? (uint64_t) s : p0) - p0;
	uint64_t b = ((uint64_t) e < p1 ? (uint64_t) e : p1) - p0;
	T w = (ival != NULL) ? (T) ival[k] : (T) 1;
	atomicAdd (diff + sd.lo + a, w);
	const uint64_t ta = a / SCAN_TILE;
	if (b < len)
		{
		atomicAdd (diff + sd.lo + b, (T) (-w));
		const uint64_t tb = b / SCAN_TILE;
		if (tb != ta)                              // same tile: the two updates cancel in the tile sum
			{
			atomicAdd (tileSum + base[sg] + ta, w);
			atomicAdd (tileSum + base[sg] + tb, (T) (-w));
			}
		}
	else atomicAdd (tileSum + base[sg] + ta, w);
	}

// exclusive prefix of the tile sums inside every segment (one block per segment)
template <typename T>
__global__ void __launch_bounds__(1024)
k_tile_prefix (const uint64_t* __restrict__ base, const T* __restrict__ tileSum, T* __restrict__ tilePrefix)
	{
	__shared__ T s_w[32];
	__shared__ T s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t t0 = base[blockIdx.x], t1 = base[blockIdx.x + 1];
	if (threadIdx.x == 0) s_carry = (T) 0;
	__syncthreads ();
	for (uint64_t c0 = t0; c0 < t1; c0 += 1024)
		{
		const uint64_t i = c0 + threadIdx.x;
		const T v = (i < t1) ? tileSum[i] : (T) 0;
		T inc = v;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			T up = warp_shfl_up<T> (inc, d);
			if (lane >= d) inc += up;
			}
		if (lane == 31) s_w[warp] = inc;
		__syncthreads ();
		T wex = (T) 0, tot = (T) 0;
		for (int w = 0; w < 32; w++) { if (w < warp) wex += s_w[w];  tot += s_w[w]; }
		const T carry = s_carry;
		if (i < t1) tilePrefix[i] = carry + wex + inc - v;
		__syncthreads ();
		if (threadIdx.x == 0) s_carry = carry + tot;
		__syncthreads ();
		}
	}

// ---------------------------------------------------------------------------
// K2: segmented inclusive scan of a tile-major array
// ---------------------------------------------------------------------------

template <typename T> struct Vec4;
template <> struct Vec4<int>
	{
	static __device__ __forceinline__ void load (const int* p, int x[4])
		{ int4 v = *reinterpret_cast<const int4*> (p);  x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
	};
template <> struct Vec4<double>
	{
	static __device__ __forceinline__ void load (const double* p, double x[4])
		{ ldg_stream4 (p, x[0], x[1], x[2], x[3]); }
	};

template <> __device__ __forceinline__ int    warp_shfl_up<int>    (int v, int d)    { return __shfl_up_sync (0xffffffffu, v, d); }
template <> __device__ __forceinline__ double warp_shfl_up<double> (double v, int d) { return shfl_up_f64 (v, d); }
template <> __device__ __forceinline__ int    warp_shfl_idx<int>    (int v, int s)    { return __shfl_sync (0xffffffffu, v, s); }
template <> __device__ __forceinline__ double warp_shfl_idx<double> (double v, int s) { return shfl_idx_f64 (v, s); }

// MODE 0: out = scan ; MODE 1: out += scan
// CHAINED: the tile's starting value comes from the decoupled look-back (generic input);
// otherwise it is read from tilePrefix (accumulate: known from the intervals themselves)
template <typename T, int MODE, bool CHAINED>
__global__ void __launch_bounds__(SCAN_THREADS, 4)
k_scan_tiles (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
              const T* __restrict__ in, double* __restrict__ out, ScanStatus<T> st,
              const T* __restrict__ tilePrefix)
	{
	__shared__ T s_warp[SCAN_WARPS];
	__shared__ T s_excl;

	const uint32_t tile = CHAINED ? scan_take_ticket (st.ticket) : blockIdx.x;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SCAN_TILE;
	const uint64_t n  = (sd.hi - t0 < SCAN_TILE) ? (sd.hi - t0) : SCAN_TILE;

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	// each warp owns SCAN_ROWS rows of 128 cells; a lane holds 4 consecutive cells of each row
	T x[SCAN_ROWS][4];
	#pragma unroll
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		uint32_t off = warp * (SCAN_ROWS * 128) + r * 128 + lane * 4;
		if (off < n)
			{
			Vec4<T>::load (in + t0 + off, x[r]);
			#pragma unroll
			for (int c = 0; c < 4; c++) if (off + c >= n) x[r][c] = (T) 0;
			}
		else
			{
			#pragma unroll
			for (int c = 0; c < 4; c++) x[r][c] = (T) 0;
			}
		}

	// lane-local inclusive scans, then warp scan of the group totals row by row
	T rowCarry = (T) 0;
	#pragma unroll
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		x[r][1] += x[r][0];  x[r][2] += x[r][1];  x[r][3] += x[r][2];
		T g = x[r][3];
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			T up = warp_shfl_up<T> (g, d);
			if (lane >= d) g += up;
			}
		T exclLane = warp_shfl_up<T> (g, 1);           // sum of everything before this lane's group in the row
		if (lane == 0) exclLane = (T) 0;
		exclLane += rowCarry;
		#pragma unroll
		for (int c = 0; c < 4; c++) x[r][c] += exclLane;
		rowCarry = warp_shfl_idx<T> (g, 31) + rowCarry;
		}
	if (lane == 31) s_warp[warp] = rowCarry;          // warp total
	__syncthreads ();

	T warpExcl = (T) 0, tileAgg = (T) 0;
	#pragma unroll
	for (int w = 0; w < SCAN_WARPS; w++)
		{
		T t = s_warp[w];
		if (w < warp) warpExcl += t;
		tileAgg += t;
		}

	T tileExcl;
	if (CHAINED)
		{
		if (threadIdx.x < 32)
			{
			const T e = scan_lookback<T> (st, tile, tis == 0, tileAgg, (T) 0, [] (T a, T b) { return a + b; });
			if (threadIdx.x == 0) s_excl = e;
			}
		__syncthreads ();
		tileExcl = s_excl;
		}
	else tileExcl = tilePrefix[tile];
	const T add = tileExcl + warpExcl;

	#pragma unroll
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		uint32_t off = warp * (SCAN_ROWS * 128) + r * 128 + lane * 4;
		if (off >= n) continue;
		double y[4];
		#pragma unroll
		for (int c = 0; c < 4; c++) y[c] = to_f64 (x[r][c] + add);
		double* o = out + t0 + off;
		if (off + 4 <= n)
			{
			if (MODE == 1)
				{
				double a0, a1, a2, a3;
				ldg_stream4 (o, a0, a1, a2, a3);
				y[0] += a0;  y[1] += a1;  y[2] += a2;  y[3] += a3;
				}
			stg_stream4 (o, y[0], y[1], y[2], y[3]);
			}
		else
			{
			for (int c = 0; c < 4 && off + c < n; c++)
				o[c] = (MODE == 1) ? o[c] + y[c] : y[c];
			}
		}
	}

// ---------------------------------------------------------------------------
// Cumulative sum (chained).  A chained tile waits for its predecessors between reading and writing,
// so what bounds the kernel is how many tiles an SM keeps in flight, i.e. registers per thread.
// k_cumsum therefore reads its tile twice: once to form the tile's sum (published at once), and --
// after the look-back -- row by row from L2 to scan and write, holding 4 cells at a time instead of
// 16 (32 registers of data).  DRAM traffic is unchanged (the second read hits L2).
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS, 8)
k_cumsum (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
          const double* __restrict__ in, double* __restrict__ out, ScanStatus<double> st)
	{
	__shared__ double s_warp[SCAN_WARPS];
	__shared__ double s_excl;
	const uint32_t tile = scan_take_ticket (st.ticket);
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SCAN_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < SCAN_TILE) ? (sd.hi - t0) : SCAN_TILE);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	// phase 1: the warp's total, rows in order, lanes in order (the same association as phase 2)
	double x[SCAN_ROWS][4];
	#pragma unroll
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		const uint32_t off = warp * (SCAN_ROWS * 128) + r * 128 + lane * 4;
		if (off + 4 <= n) ldg_stream4 (in + t0 + off, x[r][0], x[r][1], x[r][2], x[r][3]);
		else
			{
			#pragma unroll
			for (int c = 0; c < 4; c++) x[r][c] = (off + c < n) ? in[t0 + off + c] : 0.0;
			}
		}
	double warpTot = 0.0;
	#pragma unroll
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		double g = ((x[r][0] + x[r][1]) + x[r][2]) + x[r][3];
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			const double up = shfl_up_f64 (g, d);
			if (lane >= d) g += up;
			}
		warpTot += shfl_idx_f64 (g, 31);
		}
	if (lane == 0) s_warp[warp] = warpTot;
	__syncthreads ();
	double warpExcl = 0.0, tileAgg = 0.0;
	#pragma unroll
	for (int w = 0; w < SCAN_WARPS; w++) { const double t = s_warp[w];  if (w < warp) warpExcl += t;  tileAgg += t; }
	if (threadIdx.x < 32)
		{
		const double e = scan_lookback<double> (st, tile, tis == 0, tileAgg, 0.0, [] (double a, double b) { return a + b; });
		if (threadIdx.x == 0) s_excl = e;
		}
	__syncthreads ();

	// phase 2: row by row from L2
	double run = s_excl + warpExcl;
	#pragma unroll 1
	for (int r = 0; r < SCAN_ROWS; r++)
		{
		const uint32_t off = warp * (SCAN_ROWS * 128) + r * 128 + lane * 4;
		double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
		if (off + 4 <= n)
			{
			const double2 p = *reinterpret_cast<const double2*> (in + t0 + off), q = *reinterpret_cast<const double2*> (in + t0 + off + 2);
			a0 = p.x;  a1 = p.y;  a2 = q.x;  a3 = q.y;
			}
		else
			{
			if (off     < n) a0 = in[t0 + off];
			if (off + 1 < n) a1 = in[t0 + off + 1];
			if (off + 2 < n) a2 = in[t0 + off + 2];
			}
		a1 += a0;  a2 += a1;  a3 += a2;
		double g = a3;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			const double up = shfl_up_f64 (g, d);
			if (lane >= d) g += up;
			}
		double ex = shfl_up_f64 (g, 1);
		if (lane == 0) ex = 0.0;
		ex += run;
		run += shfl_idx_f64 (g, 31);
		if (off >= n) continue;
		double y0 = a0 + ex, y1 = a1 + ex, y2 = a2 + ex, y3 = a3 + ex;
		double* o = out + t0 + off;
		if (off + 4 <= n)
			{
			if (MODE == 1)
				{
				double b0, b1, b2, b3;
				ldg_stream4 (o, b0, b1, b2, b3);
				y0 += b0;  y1 += b1;  y2 += b2;  y3 += b3;
				}
			stg_stream4 (o, y0, y1, y2, y3);
			}
		else
			{
			const double y[4] = { y0, y1, y2, y3 };
			for (int c = 0; c < 4 && off + c < n; c++) o[c] = (MODE == 1) ? o[c] + y[c] : y[c];
			}
		}
	}

static int launch_scan (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, int addTo)
	{
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SCAN_TILE, &tm));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, scan_status_bytes<double> (tm.ntiles), &ws));
	ScanStatus<double> st = scan_status_carve<double> (ws, tm.ntiles);
	GDSP_CUDA (cudaMemsetAsync (ws, 0, scan_status_clear_bytes<double> (tm.ntiles), c->stream));
	if (addTo) k_cumsum<1><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st);
	else       k_cumsum<0><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// accumulate: tile sums (slot 0) were filled by k_diff
template <typename T>
static int launch_scan_prefixed (gdsp_ctx* c, gdsp_layout* L, const T* in, double* out, int addTo,
                                 const TileMap& tm, T* tileSum, T* tilePrefix)
	{
	k_tile_prefix<T><<<L->nseg, 1024, 0, c->stream>>> (tm.d_base, tileSum, tilePrefix);
	GDSP_KERNEL_CHECK ();
	ScanStatus<T> st;  st.ticket = NULL;  st.rec = NULL;
	if (addTo) k_scan_tiles<T, 1, false><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st, tilePrefix);
	else       k_scan_tiles<T, 0, false><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st, tilePrefix);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// Binned accumulate (unit-weight intervals, the `--novalue` depth case).
//
// The difference-array path above pays two random global atomics per interval (measured: 64 B read
// + 32 B written per atomic in DRAM) plus zeroing and re-reading a 4 B/bp array.  When every
// interval has weight 1 the array never has to exist in global memory:
//   K1b k_bin_count    per interval, count one record in the tile (2^LOG cells) its start falls in
//                      (one RED on an L2-resident counter) and keep the tiles' net start-end counts
//   K1c k_bin_offsets  exclusive prefix of the record counts (bucket offsets); the segment prefix of
//                      the net counts is every tile's starting depth
//   K1d k_bin_scatter  write every interval as ONE 32-bit record {kind, cell in tile, length} into
//                      its tile's bucket; lanes that hit the same tile share one cursor atomic
//   K2b k_bin_final    one block per tile: +1/-1 into shared-memory counters from the tile's own
//                      bucket and from the records of the previous tile whose end spills over,
//                      prefix sum, fp64 depth written once -- the only large global traffic
// An interval that ends in its own or the next tile and is shorter than a tile is one record
// (kind 0); anything longer becomes a start-only record (kind 1) and, unless it runs to the end of
// the chromosome, an end-only record (kind 2) in the tile of its end.
// Record order inside a bucket is irrelevant (integer adds), so the result is deterministic and
// identical to the reference loop (genodsp.c:1325-1329) for any interval order.
// ---------------------------------------------------------------------------

#define BIN_REC_SHORT 0u
#define BIN_REC_START 1u
#define BIN_REC_END   2u

template <int LOG>
struct BinGeom
	{
	static constexpr uint32_t TILE = 1u << LOG;
	bool     ok, isShort, hasB;
	uint64_t ta, tb;           // tiles of the start cell and of the end cell
	uint32_t ca, cb, len;      // cells inside those tiles; clipped length

	__device__ __forceinline__ void of (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
	                                    uint32_t sg, uint32_t s, uint32_t e)
		{
		ok = false;
		if (sg >= (uint32_t) nseg || s >= e) return;
		const SegDev sd = segs[sg];
		const uint64_t n = sd.hi - sd.lo;
		const uint64_t p0 = sd.pos0, p1 = p0 + n;
		if ((uint64_t) e <= p0 || (uint64_t) s >= p1) return;
		const uint64_t a = ((uint64_t) s > p0 ? (uint64_t) s : p0) - p0;
		const uint64_t b = ((uint64_t) e < p1 ? (uint64_t) e : p1) - p0;
		const uint64_t tb0 = base[sg];
		ok = true;
		ta = tb0 + (a >> LOG);  ca = (uint32_t) (a & (TILE - 1));
		tb = tb0 + (b >> LOG);  cb = (uint32_t) (b & (TILE - 1));
		hasB = (b < n);
		len = (uint32_t) ((b - a < TILE) ? (b - a) : TILE);
		isShort = hasB && (b - a < TILE);                   // then tb is ta or ta+1
		}
	__device__ __forceinline__ uint32_t rec_first () const
		{ return isShort ? ((BIN_REC_SHORT << 30) | (ca << LOG) | len) : ((BIN_REC_START << 30) | (ca << LOG)); }
	__device__ __forceinline__ bool     has_second () const { return ok && !isShort && hasB; }
	__device__ __forceinline__ uint32_t rec_second () const { return (BIN_REC_END << 30) | (cb << LOG); }
	};

template <int LOG>
__global__ void __launch_bounds__(256)
k_bin_count (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
             uint32_t* __restrict__ nRec, int* __restrict__ tileSum,
             const uint32_t* __restrict__ iseg, const uint32_t* __restrict__ istart,
             const uint32_t* __restrict__ iend, uint64_t n)
	{
	const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	BinGeom<LOG> g;
	g.of (segs, base, nseg, iseg[k], istart[k], iend[k]);
	if (!g.ok) return;
	atomicAdd (nRec + g.ta, 1u);
	if (g.has_second ()) atomicAdd (nRec + g.tb, 1u);
	if (!g.hasB) atomicAdd (tileSum + g.ta, 1);
	else if (g.tb != g.ta)                                   // same tile: +1 and -1 cancel in the tile's net count
		{
		atomicAdd (tileSum + g.ta, 1);
		atomicAdd (tileSum + g.tb, -1);
		}
	}

// off[t] = records in tiles before t, cursor[t] = off[t].  One block (ntiles is a few 100k).
__global__ void __launch_bounds__(1024)
k_bin_offsets (uint64_t ntiles, const uint32_t* __restrict__ nRec, uint32_t* __restrict__ off, uint32_t* __restrict__ cursor)
	{
	__shared__ uint32_t s_w[32];
	__shared__ uint32_t s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads ();
	// 4 tiles per thread per round
	for (uint64_t c0 = 0; c0 < ntiles; c0 += 4096)
		{
		const uint64_t i = c0 + 4ull * threadIdx.x;
		uint32_t v[4];
		#pragma unroll
		for (int q = 0; q < 4; q++) v[q] = (i + q < ntiles) ? nRec[i + q] : 0u;
		const uint32_t mine = v[0] + v[1] + v[2] + v[3];
		uint32_t inc = mine;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			uint32_t up = __shfl_up_sync (0xffffffffu, inc, d);
			if (lane >= d) inc += up;
			}
		if (lane == 31) s_w[warp] = inc;
		__syncthreads ();
		uint32_t wex = 0, tot = 0;
		for (int w = 0; w < 32; w++) { if (w < warp) wex += s_w[w];  tot += s_w[w]; }
		const uint32_t carry = s_carry;
		uint32_t o = carry + wex + inc - mine;
		#pragma unroll
		for (int q = 0; q < 4; q++)
			{
			if (i + q < ntiles) { off[i + q] = o;  cursor[i + q] = o; }
			o += v[q];
			}
		__syncthreads ();
		if (threadIdx.x == 0) s_carry = carry + tot;
		__syncthreads ();
		}
	if (threadIdx.x == 0) off[ntiles] = s_carry;
	}

// One slot in bucket `tile` per calling lane.  Neighbouring lanes that name the same tile share one atomic: a lane
// starts a run when the lane below it is inactive or names another tile, and the run's first lane reserves for all
// of it.  Position-sorted input gives one or two runs per warp (32 same-address atomics would serialise in L2);
// random input gives a run per lane -- there MATCH.ANY found nothing more to share either and was what bound the
// scatter kernel (ncu: 28 warps per issue in mio_throttle, 24 on the short scoreboard, L2 atomic unit 12 % busy).
__device__ __forceinline__ uint32_t bin_reserve (uint32_t* __restrict__ cursor, uint64_t tile, bool active)
	{
	const int lane = threadIdx.x & 31;
	const uint32_t t32 = (uint32_t) tile;
	const uint32_t left = __shfl_up_sync (0xffffffffu, t32, 1);
	const unsigned act = __ballot_sync (0xffffffffu, active);
	const bool head = active && (lane == 0 || ((act >> (lane - 1)) & 1u) == 0 || left != t32);
	const unsigned heads = __ballot_sync (0xffffffffu, head);
	uint32_t pos = 0;
	if (active)
		{
		const unsigned upto = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);
		const int leader = 31 - __clz (heads & upto);                       // the run's first lane
		const unsigned above = (leader == 31) ? 0u : ~((2u << leader) - 1u);
		const unsigned stop = (heads | ~act) & above;                       // the next run's head or an inactive lane
		const int end = stop ? __ffs (stop) - 1 : 32;
		uint32_t b0 = 0;
		if (lane == leader) b0 = atomicAdd (cursor + tile, (uint32_t) (end - leader));
		b0 = __shfl_sync (act, b0, leader);
		pos = b0 + (uint32_t) (lane - leader);
		}
	return pos;
	}

template <int LOG>
__global__ void __launch_bounds__(256)
k_bin_scatter (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
               uint32_t* __restrict__ cursor, uint32_t* __restrict__ recs,
               const uint32_t* __restrict__ iseg, const uint32_t* __restrict__ istart,
               const uint32_t* __restrict__ iend, uint64_t n)
	{
	const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	BinGeom<LOG> g;
	g.ok = false;  g.ta = g.tb = 0;
	if (k < n) g.of (segs, base, nseg, iseg[k], istart[k], iend[k]);
	const uint32_t pa = bin_reserve (cursor, g.ta, g.ok);
	if (g.ok) recs[pa] = g.rec_first ();
	const bool second = g.has_second ();
	if (__any_sync (0xffffffffu, second))
		{
		const uint32_t pb = bin_reserve (cursor, g.tb, second);
		if (second) recs[pb] = g.rec_second ();
		}
	}

// ---------------------------------------------------------------------------
// Fixed-capacity buckets (round 2): the count pass and the offset prefix disappear.  Every tile owns
// CAP = 2^capLog record slots (CAP ~ 2.5x the mean number of records per tile); a read takes its slot with the
// same warp-aggregated returning atomic as before, on a per-tile counter that starts at zero, and the net
// start-minus-end count per tile (the tile's starting depth after the per-chromosome prefix) is kept in the
// same pass -- only reads that cross a tile boundary touch it.  A record that finds its bucket full goes to a
// small overflow list {tile, record}; the final kernel scans that list only in tiles whose counter says they
// overflowed.  The host looks at the overflow count once: beyond BIN_OVF_MAX entries (a pile-up far above the
// mean) the whole input takes the exact-size path above instead.
// ---------------------------------------------------------------------------
#define BIN_OVF_MAX 65536u

template <int LOG>
__global__ void __launch_bounds__(256)
k_bin_scatter_fixed (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                     uint32_t* __restrict__ cursor, int* __restrict__ tileSum, uint32_t* __restrict__ recs, uint32_t capLog,
                     uint32_t* __restrict__ ovf, uint32_t* __restrict__ ovfCount,
                     const uint32_t* __restrict__ iseg, const uint32_t* __restrict__ istart,
                     const uint32_t* __restrict__ iend, uint64_t n)
	{
	const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t CAP = 1u << capLog;
	BinGeom<LOG> g;
	g.ok = false;  g.ta = g.tb = 0;
	if (k < n) g.of (segs, base, nseg, iseg[k], istart[k], iend[k]);
	const uint32_t pa = bin_reserve (cursor, g.ta, g.ok);
	if (g.ok)
		{
		if (pa < CAP) recs[(g.ta << capLog) + pa] = g.rec_first ();
		else
			{
			const uint32_t o = atomicAdd (ovfCount, 1u);
			if (o < BIN_OVF_MAX) { ovf[2 * o] = (uint32_t) g.ta;  ovf[2 * o + 1] = g.rec_first (); }
			}
		if (!g.hasB) atomicAdd (tileSum + g.ta, 1);
		else if (g.tb != g.ta) { atomicAdd (tileSum + g.ta, 1);  atomicAdd (tileSum + g.tb, -1); }
		}
	const bool second = g.has_second ();
	if (__any_sync (0xffffffffu, second))
		{
		const uint32_t pb = bin_reserve (cursor, g.tb, second);
		if (second)
			{
			if (pb < CAP) recs[(g.tb << capLog) + pb] = g.rec_second ();
			else
				{
				const uint32_t o = atomicAdd (ovfCount, 1u);
				if (o < BIN_OVF_MAX) { ovf[2 * o] = (uint32_t) g.tb;  ovf[2 * o + 1] = g.rec_second (); }
				}
			}
		}
	}

// the two ways a record changes a tile's counters
template <int LOG>
__device__ __forceinline__ void bin_apply_own (int* s_cnt, uint32_t r)
	{
	constexpr uint32_t TILE = 1u << LOG;
	const uint32_t kind = r >> 30, cell = (r >> LOG) & (TILE - 1), len = r & (TILE - 1);
	if (kind == BIN_REC_END) atomicAdd (&s_cnt[cell], -1);
	else
		{
		atomicAdd (&s_cnt[cell], 1);
		if (kind == BIN_REC_SHORT && cell + len < TILE) atomicAdd (&s_cnt[cell + len], -1);
		}
	}
template <int LOG>
__device__ __forceinline__ void bin_apply_spill (int* s_cnt, uint32_t r)        // a record of the PREVIOUS tile
	{
	constexpr uint32_t TILE = 1u << LOG;
	const uint32_t kind = r >> 30, cell = (r >> LOG) & (TILE - 1), len = r & (TILE - 1);
	if (kind == BIN_REC_SHORT && cell + len >= TILE) atomicAdd (&s_cnt[cell + len - TILE], -1);
	}

// FIXED: `off` is the per-tile record counter and a tile's bucket starts at tile << capLog
template <int LOG, int THREADS, int MODE, bool FIXED = false>
__global__ void __launch_bounds__(THREADS)
k_bin_final (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
             const uint32_t* __restrict__ off, const uint32_t* __restrict__ recs,
             const int* __restrict__ tilePrefix, double* __restrict__ out,
             uint32_t capLog = 0, const uint32_t* __restrict__ ovf = NULL, const uint32_t* __restrict__ ovfCount = NULL)
	{
	constexpr uint32_t TILE  = 1u << LOG;
	constexpr int      WARPS = THREADS / 32;
	constexpr int      ROWS  = TILE / (WARPS * 128);        // rows of 128 cells per warp
	extern __shared__ __align__(16) int s_cnt[];            // TILE counters
	__shared__ int s_warp[WARPS];

	const uint64_t tile = blockIdx.x;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < TILE) ? (sd.hi - t0) : TILE);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

	// bucket bounds first: the loads are in flight while the counters are cleared
	uint32_t e0 = 0, e1 = 0, p0 = 0, nOwn = 0, nPrev = 0;
	if (FIXED) { nOwn = off[tile];  nPrev = (tis > 0) ? off[tile - 1] : 0u; }
	else
		{
		e0 = off[tile];  e1 = off[tile + 1];
		p0 = (tis > 0) ? off[tile - 1] : e0;                // previous tile of the same chromosome: [p0, e0)
		}
	for (uint32_t i = threadIdx.x; i < TILE / 4; i += THREADS)
		reinterpret_cast<int4*> (s_cnt)[i] = make_int4 (0, 0, 0, 0);
	__syncthreads ();

	if (FIXED)
		{
		const uint32_t CAP = 1u << capLog;
		const uint32_t cOwn = (nOwn < CAP) ? nOwn : CAP, cPrev = (nPrev < CAP) ? nPrev : CAP;
		const uint32_t* own  = recs + (tile << capLog);
		const uint32_t* prev = recs + ((tile - (tis > 0 ? 1 : 0)) << capLog);
		for (uint32_t i = threadIdx.x; i < cOwn; i += THREADS)  bin_apply_own<LOG> (s_cnt, __ldg (own + i));
		for (uint32_t i = threadIdx.x; i < cPrev; i += THREADS) bin_apply_spill<LOG> (s_cnt, __ldg (prev + i));
		if (nOwn > CAP || nPrev > CAP)                       // this tile or the one before it overflowed: the list holds their rest
			{
			const uint32_t no = (*ovfCount < BIN_OVF_MAX) ? *ovfCount : BIN_OVF_MAX;
			for (uint32_t i = threadIdx.x; i < no; i += THREADS)
				{
				const uint32_t t = __ldg (ovf + 2 * i), r = __ldg (ovf + 2 * i + 1);
				if (t == (uint32_t) tile) bin_apply_own<LOG> (s_cnt, r);
				else if (tis > 0 && t == (uint32_t) (tile - 1)) bin_apply_spill<LOG> (s_cnt, r);
				}
			}
		}
	else
	for (uint32_t i = p0 + threadIdx.x; i < e1; i += THREADS)
		{
		const uint32_t r = __ldg (recs + i);
		const uint32_t kind = r >> 30, cell = (r >> LOG) & (TILE - 1), len = r & (TILE - 1);
		if (i >= e0)
			{
			if (kind == BIN_REC_END) atomicAdd (&s_cnt[cell], -1);
			else
				{
				atomicAdd (&s_cnt[cell], 1);
				if (kind == BIN_REC_SHORT && cell + len < TILE) atomicAdd (&s_cnt[cell + len], -1);
				}
			}
		else if (kind == BIN_REC_SHORT && cell + len >= TILE) atomicAdd (&s_cnt[cell + len - TILE], -1);
		}
	__syncthreads ();

	// every warp scans its own stripe of ROWS * 128 cells in place
	int* stripe = s_cnt + warp * (ROWS * 128);
	int rowCarry = 0;
	#pragma unroll
	for (int r = 0; r < ROWS; r++)
		{
		int4 v = reinterpret_cast<int4*> (stripe + r * 128)[lane];
		v.y += v.x;  v.z += v.y;  v.w += v.z;
		int g = v.w;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			int up = __shfl_up_sync (0xffffffffu, g, d);
			if (lane >= d) g += up;
			}
		int ex = __shfl_up_sync (0xffffffffu, g, 1);
		if (lane == 0) ex = 0;
		ex += rowCarry;
		v.x += ex;  v.y += ex;  v.z += ex;  v.w += ex;
		reinterpret_cast<int4*> (stripe + r * 128)[lane] = v;
		rowCarry += __shfl_sync (0xffffffffu, g, 31);
		}
	if (lane == 0) s_warp[warp] = rowCarry;
	__syncthreads ();
	int add = tilePrefix[tile];
	#pragma unroll
	for (int w = 0; w < WARPS; w++) if (w < warp) add += s_warp[w];

	#pragma unroll
	for (int r = 0; r < ROWS; r++)
		{
		const uint32_t o0 = warp * (ROWS * 128) + r * 128 + lane * 4;
		if (o0 >= n) continue;
		const int4 v = reinterpret_cast<const int4*> (stripe + r * 128)[lane];
		double y[4] = { i32_to_f64 (v.x + add), i32_to_f64 (v.y + add), i32_to_f64 (v.z + add), i32_to_f64 (v.w + add) };
		double* o = out + t0 + o0;
		if (o0 + 4 <= n)
			{
			if (MODE == 1)
				{
				double a0, a1, a2, a3;
				ldg_stream4 (o, a0, a1, a2, a3);
				y[0] += a0;  y[1] += a1;  y[2] += a2;  y[3] += a3;
				}
			stg_stream4 (o, y[0], y[1], y[2], y[3]);
			}
		else
			{
			for (int c = 0; c < 4 && o0 + c < n; c++) o[c] = (MODE == 1) ? o[c] + y[c] : y[c];
			}
		}
	}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------

extern "C" size_t gdsp_accumulate_work_bytes (const gdsp_layout* L, uint64_t buffer_cells, int mode)
	{
	(void) L;
	return (size_t) buffer_cells * (mode == GDSP_ACC_I32 ? sizeof (int) : sizeof (double));
	}

struct AccTiles { TileMap tm;  void* tileSum;  void* tilePrefix; };

static int accumulate_begin (gdsp_ctx* c, gdsp_layout* L, uint64_t buffer_cells, void* work, int mode, AccTiles* at)
	{
	const size_t esz = (mode == GDSP_ACC_I32) ? sizeof (int) : sizeof (double);
	GDSP_CUDA (cudaMemsetAsync (work, 0, (size_t) buffer_cells * esz, c->stream));
	GDSP_TRY (gdsp_layout_tilemap (L, SCAN_TILE, &at->tm));
	const size_t tb = ((at->tm.ntiles * esz + 255) / 256) * 256;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, 2 * tb, &ws));
	at->tileSum = ws;  at->tilePrefix = (char*) ws + tb;
	GDSP_CUDA (cudaMemsetAsync (ws, 0, tb, c->stream));
	return GDSP_OK;
	}

static int accumulate_chunk (gdsp_ctx* c, gdsp_layout* L, void* work, const AccTiles& at, const uint32_t* d_seg,
                             const uint32_t* d_start, const uint32_t* d_end, const double* d_val,
                             uint64_t n, int mode)
	{
	if (n == 0) return GDSP_OK;
	unsigned blocks = (unsigned) ((n + 255) / 256);
	if (mode == GDSP_ACC_I32)
		k_diff<int><<<blocks, 256, 0, c->stream>>> (L->d, at.tm.d_base, L->nseg, (int*) work, (int*) at.tileSum,
		                                            d_seg, d_start, d_end, d_val, n);
	else
		k_diff<double><<<blocks, 256, 0, c->stream>>> (L->d, at.tm.d_base, L->nseg, (double*) work, (double*) at.tileSum,
		                                               d_seg, d_start, d_end, d_val, n);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

static int accumulate_finish (gdsp_ctx* c, gdsp_layout* L, double* sig, void* work, const AccTiles& at, int mode, int addTo)
	{
	if (mode == GDSP_ACC_I32)
		return launch_scan_prefixed<int> (c, L, (const int*) work, sig, addTo, at.tm, (int*) at.tileSum, (int*) at.tilePrefix);
	return launch_scan_prefixed<double> (c, L, (const double*) work, sig, addTo, at.tm, (double*) at.tileSum, (double*) at.tilePrefix);
	}

// ---- binned path: carving of the caller's work buffer (buffer_cells * 4 bytes in I32 mode) ----
//   [records u32 x 2n][nRec][off (+1)][cursor][tileSum][tilePrefix]  (+ [seg][start][end] staging
//   for host arrays)
struct BinCarve
	{
	uint32_t *recs, *nRec, *off, *cursor;  int *tileSum, *tilePrefix;
	uint32_t *seg, *start, *end;
	size_t total;
	};

static inline size_t up256 (size_t x) { return (x + 255) / 256 * 256; }

static BinCarve bin_carve (void* work, uint64_t ntiles, uint64_t n, bool staging)
	{
	BinCarve b;
	char* p = (char*) work;
	const size_t tb = up256 ((ntiles + 1) * sizeof (uint32_t));
	b.recs   = (uint32_t*) p;   p += up256 (2 * n * sizeof (uint32_t));
	b.nRec   = (uint32_t*) p;   p += tb;
	b.tileSum    = (int*) p;    p += tb;
	b.off    = (uint32_t*) p;   p += tb;
	b.cursor = (uint32_t*) p;   p += tb;
	b.tilePrefix = (int*) p;    p += tb;
	b.seg = b.start = b.end = NULL;
	if (staging)
		{
		b.seg   = (uint32_t*) p;  p += up256 (n * sizeof (uint32_t));
		b.start = (uint32_t*) p;  p += up256 (n * sizeof (uint32_t));
		b.end   = (uint32_t*) p;  p += up256 (n * sizeof (uint32_t));
		}
	b.total = (size_t) (p - (char*) work);
	return b;
	}

// 8192-cell tiles (32 KB of counters, 6 blocks of 256 threads per SM): hg38 depth 5 takes 10.1 ms
// against 11.5 ms with 16384-cell tiles and 512 threads
#define BIN_LOG     13
#define BIN_THREADS 256

static bool binned_fits (gdsp_layout* L, uint64_t buffer_cells, uint64_t n, bool staging)
	{
	if (getenv ("GDSP_ACCUMULATE_DIFF")) return false;            // force the difference-array path (tests, A/B timing)
	if (2 * n >= 0xffffffffull) return false;                     // 32-bit bucket offsets
	const uint64_t T = 1ull << BIN_LOG;
	uint64_t ntiles = 0;
	for (int s = 0; s < L->nseg; s++) ntiles += (L->h[s].hi - L->h[s].lo + T - 1) / T;
	return bin_carve (NULL, ntiles, n, staging).total <= (size_t) buffer_cells * sizeof (int);
	}

// fixed-capacity buckets: returns *done = 0 (nothing written) when the buckets do not fit the work buffer or
// more than BIN_OVF_MAX records overflowed; the caller then takes the exact-size path
template <int LOG, int THREADS>
static int accumulate_binned_fixed_t (gdsp_ctx* c, gdsp_layout* L, double* sig, void* work, size_t workBytes,
                                      const uint32_t* d_seg, const uint32_t* d_start, const uint32_t* d_end,
                                      uint64_t n, int addTo, int* done)
	{
	*done = 0;
	if (getenv ("GDSP_ACCUMULATE_EXACT_BUCKETS")) return GDSP_OK;           // tests / A-B timing: the count + prefix path
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, 1u << LOG, &tm));
	if (tm.ntiles == 0 || tm.ntiles >= 0xffffffffull) return GDSP_OK;
	// CAP: a power of two >= 2x the mean number of records per tile (+ slack for short inputs), 64 .. 4096
	const uint64_t mean = (n + n / 64 + tm.ntiles - 1) / tm.ntiles;
	uint32_t capLog = 6;
	while ((1ull << capLog) < 2 * mean + 32 && capLog < 12) capLog++;
	const size_t tb = up256 ((tm.ntiles + 1) * sizeof (uint32_t));
	const size_t need = up256 (((size_t) tm.ntiles << capLog) * sizeof (uint32_t)) + 3 * tb + 256 + up256 (2 * (size_t) BIN_OVF_MAX * sizeof (uint32_t));
	if (need > workBytes) return GDSP_OK;
	char* p = (char*) work;
	uint32_t* recs   = (uint32_t*) p;   p += up256 (((size_t) tm.ntiles << capLog) * sizeof (uint32_t));
	uint32_t* cursor = (uint32_t*) p;   p += tb;
	int* tileSum     = (int*) p;        p += tb;
	uint32_t* ovfCount = (uint32_t*) p; p += 256;
	int* tilePrefix  = (int*) p;        p += tb;
	uint32_t* ovf    = (uint32_t*) p;
	GDSP_CUDA (cudaMemsetAsync (cursor, 0, 2 * tb + 256, c->stream));       // cursor, tileSum, ovfCount
	const unsigned blocks = (unsigned) ((n + 255) / 256);
	k_bin_scatter_fixed<LOG><<<blocks, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, cursor, tileSum, recs, capLog, ovf, ovfCount,
	                                                        d_seg, d_start, d_end, n);
	GDSP_KERNEL_CHECK ();
	void* hp;
	GDSP_TRY (gdsp_host_scratch (c, 64, &hp));
	uint32_t* h_ovf = (uint32_t*) hp;
	GDSP_CUDA (cudaMemcpyAsync (h_ovf, ovfCount, sizeof (uint32_t), cudaMemcpyDeviceToHost, c->stream));
	k_tile_prefix<int><<<L->nseg, 1024, 0, c->stream>>> (tm.d_base, tileSum, tilePrefix);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (*h_ovf > BIN_OVF_MAX) return GDSP_OK;                               // a pile-up far above the mean: exact-size buckets
	const int smem = (int) (sizeof (int) << LOG);
	if (addTo)
		{
		GDSP_CUDA (cudaFuncSetAttribute (k_bin_final<LOG, THREADS, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k_bin_final<LOG, THREADS, 1, true><<<(unsigned) tm.ntiles, THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, cursor, recs, tilePrefix, sig, capLog, ovf, ovfCount);
		}
	else
		{
		GDSP_CUDA (cudaFuncSetAttribute (k_bin_final<LOG, THREADS, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k_bin_final<LOG, THREADS, 0, true><<<(unsigned) tm.ntiles, THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, cursor, recs, tilePrefix, sig, capLog, ovf, ovfCount);
		}
	GDSP_KERNEL_CHECK ();
	*done = 1;
	return GDSP_OK;
	}

template <int LOG, int THREADS>
static int accumulate_binned_t (gdsp_ctx* c, gdsp_layout* L, double* sig, void* work,
                                const uint32_t* d_seg, const uint32_t* d_start, const uint32_t* d_end,
                                uint64_t n, int addTo)
	{
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, 1u << LOG, &tm));
	BinCarve b = bin_carve (work, tm.ntiles, n, false);
	GDSP_CUDA (cudaMemsetAsync (b.nRec, 0, (char*) b.off - (char*) b.nRec, c->stream));      // nRec and tileSum
	const unsigned blocks = (unsigned) ((n + 255) / 256);
	k_bin_count<LOG><<<blocks, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, b.nRec, b.tileSum, d_seg, d_start, d_end, n);
	GDSP_KERNEL_CHECK ();
	k_bin_offsets<<<1, 1024, 0, c->stream>>> (tm.ntiles, b.nRec, b.off, b.cursor);
	GDSP_KERNEL_CHECK ();
	k_tile_prefix<int><<<L->nseg, 1024, 0, c->stream>>> (tm.d_base, b.tileSum, b.tilePrefix);
	GDSP_KERNEL_CHECK ();
	k_bin_scatter<LOG><<<blocks, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, b.cursor, b.recs, d_seg, d_start, d_end, n);
	GDSP_KERNEL_CHECK ();
	const int smem = (int) (sizeof (int) << LOG);
	if (addTo)
		{
		GDSP_CUDA (cudaFuncSetAttribute (k_bin_final<LOG, THREADS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k_bin_final<LOG, THREADS, 1><<<(unsigned) tm.ntiles, THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, b.off, b.recs, b.tilePrefix, sig);
		}
	else
		{
		GDSP_CUDA (cudaFuncSetAttribute (k_bin_final<LOG, THREADS, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
		k_bin_final<LOG, THREADS, 0><<<(unsigned) tm.ntiles, THREADS, smem, c->stream>>> (L->d, tm.d_base, L->nseg, b.off, b.recs, b.tilePrefix, sig);
		}
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// fixedOffset: bytes at the start of `work` the fixed-capacity attempt must leave alone (the host entry point
// stages the interval arrays there)
static int accumulate_binned (gdsp_ctx* c, gdsp_layout* L, double* sig, uint64_t buffer_cells, void* work,
                              const uint32_t* d_seg, const uint32_t* d_start, const uint32_t* d_end,
                              uint64_t n, int addTo, size_t fixedOffset = 0)
	{
	int done = 0;
	const size_t workBytes = (size_t) buffer_cells * sizeof (int);
	fixedOffset = up256 (fixedOffset);
	if (fixedOffset < workBytes)
		GDSP_TRY ((accumulate_binned_fixed_t<BIN_LOG, BIN_THREADS> (c, L, sig, (char*) work + fixedOffset, workBytes - fixedOffset,
		                                                            d_seg, d_start, d_end, n, addTo, &done)));
	if (done) return GDSP_OK;
	return accumulate_binned_t<BIN_LOG, BIN_THREADS> (c, L, sig, work, d_seg, d_start, d_end, n, addTo);
	}

extern "C" int gdsp_accumulate_dev (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint64_t buffer_cells,
                                    void* work, const uint32_t* d_seg, const uint32_t* d_start,
                                    const uint32_t* d_end, const double* d_val, uint64_t n,
                                    int mode, int addTo)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && work, "gdsp_accumulate_dev: NULL argument");
	GDSP_REQUIRE (mode == GDSP_ACC_I32 || mode == GDSP_ACC_F64, "gdsp_accumulate_dev: bad mode %d", mode);
	GDSP_REQUIRE (n == 0 || (d_seg && d_start && d_end), "gdsp_accumulate_dev: NULL interval arrays");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_accumulate_dev");  GDSP_REQUIRE_ALIGNED (work, "gdsp_accumulate_dev");
	if (mode == GDSP_ACC_I32 && d_val == NULL && n > 0 && binned_fits (L, buffer_cells, n, false))
		return accumulate_binned (c, L, sig, buffer_cells, work, d_seg, d_start, d_end, n, addTo);
	AccTiles at;
	GDSP_TRY (accumulate_begin (c, L, buffer_cells, work, mode, &at));
	GDSP_TRY (accumulate_chunk (c, L, work, at, d_seg, d_start, d_end, d_val, n, mode));
	return accumulate_finish (c, L, sig, work, at, mode, addTo);
	}

// host arrays: streamed through the context's two pinned staging buffers (or
// copied directly when the caller's memory is already page-locked)
extern "C" int gdsp_accumulate_host (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint64_t buffer_cells,
                                     void* work, const uint32_t* h_seg, const uint32_t* h_start,
                                     const uint32_t* h_end, const double* h_val, uint64_t n,
                                     int mode, int addTo)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && work, "gdsp_accumulate_host: NULL argument");
	GDSP_REQUIRE (mode == GDSP_ACC_I32 || mode == GDSP_ACC_F64, "gdsp_accumulate_host: bad mode %d", mode);
	GDSP_REQUIRE (n == 0 || (h_seg && h_start && h_end), "gdsp_accumulate_host: NULL interval arrays");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_accumulate_host");  GDSP_REQUIRE_ALIGNED (work, "gdsp_accumulate_host");
	// unit-weight intervals: stage the arrays on the device (inside `work`) and take the binned path
	const bool binned = (mode == GDSP_ACC_I32 && h_val == NULL && n > 0 && binned_fits (L, buffer_cells, n, true));
	BinCarve bc;
	if (binned)
		{
		TileMap tmb;
		GDSP_TRY (gdsp_layout_tilemap (L, 1u << BIN_LOG, &tmb));
		bc = bin_carve (work, tmb.ntiles, n, true);
		}
	AccTiles at;
	if (!binned) GDSP_TRY (accumulate_begin (c, L, buffer_cells, work, mode, &at));

	const uint64_t CHUNK = 8u << 20;                        // intervals per chunk
	const size_t   rec   = 3 * sizeof (uint32_t) + (h_val ? sizeof (double) : 0);
	const uint64_t chunk = n < CHUNK ? (n ? n : 1) : CHUNK;
	const size_t   cbytes = (((size_t) chunk * rec + 64) + 255) / 256 * 256;

	// device staging: two chunks
	void* dws;
	GDSP_TRY (gdsp_ws (c, 1, 2 * cbytes, &dws));

	cudaPointerAttributes pattr;
	bool pinnedSrc = (cudaPointerGetAttributes (&pattr, h_seg) == cudaSuccess) && (pattr.type == cudaMemoryTypeHost);
	cudaGetLastError ();
	if (!pinnedSrc && c->pinned_bytes < cbytes)
		{
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		for (int i = 0; i < 2; i++)
			{
			if (c->pinned[i]) { cudaFreeHost (c->pinned[i]);  c->pinned[i] = NULL; }
			GDSP_CUDA (cudaMallocHost (&c->pinned[i], cbytes));
			}
		c->pinned_bytes = cbytes;
		}

	int buf = 0;
	for (uint64_t k0 = 0; k0 < n; k0 += chunk, buf ^= 1)
		{
		uint64_t m = (n - k0 < chunk) ? n - k0 : chunk;
		char* dbase = (char*) dws + (size_t) buf * cbytes;
		uint32_t* d_seg   = (uint32_t*) dbase;
		uint32_t* d_start = d_seg + m;
		uint32_t* d_end   = d_start + m;
		double*   d_val   = h_val ? (double*) (dbase + (((size_t) 3 * m * sizeof (uint32_t) + 15) / 16) * 16) : NULL;
		if (binned) { d_seg = bc.seg + k0;  d_start = bc.start + k0;  d_end = bc.end + k0; }
		if (pinnedSrc)
			{
			GDSP_CUDA (cudaMemcpyAsync (d_seg,   h_seg + k0,   m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaMemcpyAsync (d_start, h_start + k0, m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaMemcpyAsync (d_end,   h_end + k0,   m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
			if (h_val) GDSP_CUDA (cudaMemcpyAsync (d_val, h_val + k0, m * sizeof (double), cudaMemcpyHostToDevice, c->stream));
			}
		else
			{
			// wait until the previous use of this staging buffer has been consumed
			GDSP_CUDA (cudaEventSynchronize (c->pinned_ev[buf]));
			char* hp = (char*) c->pinned[buf];
			memcpy (hp,                                   h_seg + k0,   m * sizeof (uint32_t));
			memcpy (hp + m * sizeof (uint32_t),           h_start + k0, m * sizeof (uint32_t));
			memcpy (hp + 2 * m * sizeof (uint32_t),       h_end + k0,   m * sizeof (uint32_t));
			size_t voff = (((size_t) 3 * m * sizeof (uint32_t) + 15) / 16) * 16;
			size_t tot  = 3 * m * sizeof (uint32_t);
			if (h_val) { memcpy (hp + voff, h_val + k0, m * sizeof (double));  tot = voff + m * sizeof (double); }
			if (binned)
				{
				GDSP_CUDA (cudaMemcpyAsync (d_seg,   hp,                             m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
				GDSP_CUDA (cudaMemcpyAsync (d_start, hp + m * sizeof (uint32_t),     m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
				GDSP_CUDA (cudaMemcpyAsync (d_end,   hp + 2 * m * sizeof (uint32_t), m * sizeof (uint32_t), cudaMemcpyHostToDevice, c->stream));
				}
			else GDSP_CUDA (cudaMemcpyAsync (dbase, hp, tot, cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaEventRecord (c->pinned_ev[buf], c->stream));
			}
		if (!binned) GDSP_TRY (accumulate_chunk (c, L, work, at, d_seg, d_start, d_end, d_val, m, mode));
		}
	if (binned) return accumulate_binned (c, L, sig, buffer_cells, work, bc.seg, bc.start, bc.end, n, addTo, bc.total);
	return accumulate_finish (c, L, sig, work, at, mode, addTo);
	}

extern "C" int gdsp_cumulative_sum (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out, "gdsp_cumulative_sum: NULL argument");
	GDSP_REQUIRE_ALIGNED (in, "gdsp_cumulative_sum");  GDSP_REQUIRE_ALIGNED (out, "gdsp_cumulative_sum");
	if (c->exact_order) return gdsp_cumulative_sum_exact (c, L, in, out);
	return launch_scan (c, L, in, out, 0);
	}
