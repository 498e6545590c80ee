// gdsp_accumulate.cu -- interval accumulation (difference array + segmented scan)
// and cumulative sum.
//
// Replaces the accumulate loops of read_intervals (genodsp.c:1307-1330, sum
// overlap) and op_cumulative_sum_apply (sum.c:776-792).
//
// K1  k_diff_*      one thread per interval: +w at the interval start, -w at
//                   its end (clipped to the owned piece of the chromosome)
// K2  k_scan_tiles  segmented inclusive prefix sum, single pass with decoupled
//                   look-back; reads the difference array (int32 or fp64) and
//                   writes fp64 depth
//
// Algorithmic bytes: int32 path 4 B/bp zero + 4 B/bp read + 8 B/bp
// write + 28 B/interval; fp64 path 8+8+8 B/bp + 52 B/interval.
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
		{ double2 a = ldg_stream (p), b = ldg_stream (p + 2);  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; }
	};

template <> __device__ __forceinline__ int    warp_shfl_up<int>    (int v, int d)    { return __shfl_up_sync (0xffffffffu, v, d); }
template <> __device__ __forceinline__ double warp_shfl_up<double> (double v, int d) { return shfl_up_f64 (v, d); }
template <> __device__ __forceinline__ int    warp_shfl_idx<int>    (int v, int s)    { return __shfl_sync (0xffffffffu, v, s); }
template <> __device__ __forceinline__ double warp_shfl_idx<double> (double v, int s) { return shfl_idx_f64 (v, s); }

// MODE 0: out = scan ; MODE 1: out += scan
// CHAINED: the tile's starting value comes from the decoupled look-back (generic input);
// otherwise it is read from tilePrefix (accumulate: known from the intervals themselves)
template <typename T, int MODE, bool CHAINED>
__global__ void __launch_bounds__(SCAN_THREADS)
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
		for (int c = 0; c < 4; c++) y[c] = (double) (x[r][c] + add);
		double* o = out + t0 + off;
		if (off + 4 <= n)
			{
			if (MODE == 1)
				{
				double2 a = *reinterpret_cast<const double2*> (o), b = *reinterpret_cast<const double2*> (o + 2);
				y[0] += a.x;  y[1] += a.y;  y[2] += b.x;  y[3] += b.y;
				}
			stg_stream (o,     make_double2 (y[0], y[1]));
			stg_stream (o + 2, make_double2 (y[2], y[3]));
			}
		else
			{
			for (int c = 0; c < 4 && off + c < n; c++)
				o[c] = (MODE == 1) ? o[c] + y[c] : y[c];
			}
		}
	}

template <typename T>
static int launch_scan (gdsp_ctx* c, gdsp_layout* L, const T* in, double* out, int addTo)
	{
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SCAN_TILE, &tm));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, scan_status_bytes<T> (tm.ntiles), &ws));
	ScanStatus<T> st = scan_status_carve<T> (ws, tm.ntiles);
	GDSP_CUDA (cudaMemsetAsync (ws, 0, scan_status_clear_bytes<T> (tm.ntiles), c->stream));
	if (addTo) k_scan_tiles<T, 1, true><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st, NULL);
	else       k_scan_tiles<T, 0, true><<<(unsigned) tm.ntiles, SCAN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, st, NULL);
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

extern "C" int gdsp_accumulate_dev (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint64_t buffer_cells,
                                    void* work, const uint32_t* d_seg, const uint32_t* d_start,
                                    const uint32_t* d_end, const double* d_val, uint64_t n,
                                    int mode, int addTo)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && work, "gdsp_accumulate_dev: NULL argument");
	GDSP_REQUIRE (mode == GDSP_ACC_I32 || mode == GDSP_ACC_F64, "gdsp_accumulate_dev: bad mode %d", mode);
	GDSP_REQUIRE (n == 0 || (d_seg && d_start && d_end), "gdsp_accumulate_dev: NULL interval arrays");
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
	AccTiles at;
	GDSP_TRY (accumulate_begin (c, L, buffer_cells, work, mode, &at));

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
			GDSP_CUDA (cudaMemcpyAsync (dbase, hp, tot, cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaEventRecord (c->pinned_ev[buf], c->stream));
			}
		GDSP_TRY (accumulate_chunk (c, L, work, at, d_seg, d_start, d_end, d_val, m, mode));
		}
	return accumulate_finish (c, L, sig, work, at, mode, addTo);
	}

extern "C" int gdsp_cumulative_sum (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out, "gdsp_cumulative_sum: NULL argument");
	return launch_scan<double> (c, L, in, out, 0);
	}
