// gdsp_runs.cu -- run-length output stage.
//
// Replaces the per-base state machine of report_intervals (genodsp.c:1589-1678):
// the GPU finds the maximal runs of raw-equal values, drops runs of value 0
// unless uncovered bases are shown, and compacts (start, end, value) records;
// the host only formats text (and derives the NA gap lines).
//
//   printable(i) = show || v[i] != 0
//   head(i) = printable(i) && (first cell || !collapse || !printable(i-1) || v[i] != v[i-1])
//   tail(i) = printable(i) && (last cell  || !collapse || !printable(i+1) || v[i+1] != v[i])
// Heads and tails alternate, so the k-th head and the k-th tail bound the k-th
// run.  Three chain-free kernels:
//   k_runs_flags    every warp reads 512 cells (16 consecutive cells per lane, four 256-bit loads; the
//                   cell before and after a lane's strip come from the neighbouring lanes, across
//                   warps straight from global memory -- no block barrier at all), leaves one head
//                   bit and one tail bit per cell and adds its counts to its tile's counters
//   k_runs_offsets  one block: exclusive prefix of the per-tile counts (377 k tiles for hg38)
//   k_runs_emit     tiles that hold a head or a tail (few, for thresholded tracks) walk their bit
//                   words and write (start, value) / end records to their slots
// The first version did all of this in one kernel with a decoupled look-back across tiles: ncu showed
// 20 warps waiting at the block barrier per issuing warp (3.0 TB/s); the streaming pass below has no
// barrier and no chain.
// Algorithmic bytes: 8 B/bp read + 16 B/run written (+0.25 B/bp of flag bits).
#include "gdsp_common.cuh"

#define RUN_THREADS 512
#define RUN_PER     16
#define RUN_TILE    (RUN_THREADS * RUN_PER)      // 8192
#define RUN_WORDS   (RUN_TILE / 32)              // 256 flag words per tile

struct RunWork
	{
	uint32_t*           headW;      // ntiles * 256
	uint32_t*           tailW;
	unsigned int*       cntH;       // ntiles: heads per tile (zeroed before k_runs_flags)
	unsigned int*       cntT;
	unsigned long long* offH;       // ntiles + 1: heads before tile t
	unsigned long long* offT;
	unsigned long long* segFirst;   // nseg + 1
	};

// 256-thread blocks, two per tile: more, smaller blocks per SM keep loads in flight while others compute
#define RUN_FLAG_THREADS 256
__global__ void __launch_bounds__(RUN_FLAG_THREADS, 4)
k_runs_flags (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
              const double* __restrict__ sig, int collapse, int show, RunWork wk)
	{
	const uint64_t tile = blockIdx.x >> 1;
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * RUN_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < RUN_TILE) ? (sd.hi - t0) : RUN_TILE);

	const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) + (blockIdx.x & 1) * (RUN_FLAG_THREADS / 32);   // warp of the tile
	const uint32_t g0 = warp * 32 * RUN_PER;          // first cell of this warp's 512-cell group
	const uint32_t c0 = g0 + lane * RUN_PER;
	uint32_t heads = 0, tails = 0;

	if (g0 + 32 * RUN_PER <= n)
		{
		double myv[RUN_PER];
		const double* p = sig + t0 + c0;
		#pragma unroll
		for (int q = 0; q < RUN_PER / 4; q++) ldg_stream4 (p + 4 * q, myv[4*q], myv[4*q+1], myv[4*q+2], myv[4*q+3]);
		const bool isFirst = (t0 + c0 == sd.lo);                       // the lane's first cell opens the chromosome piece
		const bool isLast  = (t0 + c0 + RUN_PER == sd.hi);             // its last cell closes it
		double before = shfl_up_f64 (myv[RUN_PER - 1], 1), after = shfl_down_f64 (myv[0], 1);
		if (lane == 0)  before = isFirst ? 0.0 : __ldg (sig + t0 + c0 - 1);
		if (lane == 31) after  = isLast  ? 0.0 : __ldg (sig + t0 + c0 + RUN_PER);

		// pr: printable, ne: differs from the cell before
		uint32_t pr = 0, ne = 0;
		#pragma unroll
		for (int k = 0; k < RUN_PER; k++)
			{
			const double pv = (k == 0) ? before : myv[k - 1];
			if (show || myv[k] != 0) pr |= 1u << k;
			if (myv[k] != pv)        ne |= 1u << k;
			}
		const uint32_t prBefore = (!isFirst && (show || before != 0)) ? 1u : 0u;
		const uint32_t prAfter  = (!isLast  && (show || after  != 0)) ? 1u : 0u;
		const uint32_t neAfter  = (after != myv[RUN_PER - 1]) ? 1u : 0u;
		const uint32_t prPrev = ((pr << 1) | prBefore) & 0xffffu;              // printable(c-1), 0 at the piece start
		const uint32_t prNext = (pr >> 1) | (prAfter << (RUN_PER - 1));        // printable(c+1), 0 at the piece end
		const uint32_t neNext = (ne >> 1) | (neAfter << (RUN_PER - 1));        // v[c+1] != v[c]
		if (collapse)
			{
			heads = pr & (~prPrev | ne | (isFirst ? 1u : 0u));
			tails = pr & (~prNext | neNext | (isLast ? (1u << (RUN_PER - 1)) : 0u));
			}
		else heads = tails = pr;
		heads &= 0xffffu;  tails &= 0xffffu;
		}
	else if (c0 < n)
		{
		// the last, shorter group of a chromosome piece: cells and their neighbours one by one
		#pragma unroll 4
		for (int k = 0; k < RUN_PER; k++)
			{
			const uint32_t c = c0 + k;
			if (c >= n) break;
			const bool first = (t0 + c == sd.lo), last = (t0 + c + 1 == sd.hi);
			const double cur = sig[t0 + c];
			const bool pr = show || (cur != 0);
			if (pr)
				{
				const double prev = first ? 0.0 : sig[t0 + c - 1];
				const double next = last  ? 0.0 : sig[t0 + c + 1];
				const bool prPrev = !first && (show || prev != 0);
				const bool prNext = !last  && (show || next != 0);
				if (first || !collapse || !prPrev || cur != prev)  heads |= 1u << k;
				if (last  || !collapse || !prNext || next != cur)  tails |= 1u << k;
				}
			}
		}

	// one word per two lanes; every word of the tile is written (zeros past the end of the piece)
	uint32_t hw = heads << ((lane & 1) * 16), tw = tails << ((lane & 1) * 16);
	hw |= __shfl_xor_sync (0xffffffffu, hw, 1);
	tw |= __shfl_xor_sync (0xffffffffu, tw, 1);
	if ((lane & 1) == 0)
		{
		const uint64_t w = tile * RUN_WORDS + warp * 16 + (lane >> 1);
		wk.headW[w] = hw;  wk.tailW[w] = tw;
		}
	const unsigned nh = __reduce_add_sync (0xffffffffu, (unsigned) __popc (heads));
	const unsigned nt = __reduce_add_sync (0xffffffffu, (unsigned) __popc (tails));
	if (lane == 0)
		{
		if (nh) atomicAdd (&wk.cntH[tile], nh);
		if (nt) atomicAdd (&wk.cntT[tile], nt);
		}
	}

// exclusive prefix of the per-tile counts: block totals, then every block adds the totals before it
#define RUN_OFF_THREADS 1024
#define RUN_OFF_TILES   (RUN_OFF_THREADS * 4)
__device__ __forceinline__ void run_load4 (const unsigned int* __restrict__ a, uint64_t i, uint64_t n, unsigned int v[4])
	{
	if (i + 4 <= n) { const uint4 q = *reinterpret_cast<const uint4*> (a + i);  v[0] = q.x;  v[1] = q.y;  v[2] = q.z;  v[3] = q.w; }
	else for (int k = 0; k < 4; k++) v[k] = (i + k < n) ? a[i + k] : 0u;
	}

__global__ void __launch_bounds__(RUN_OFF_THREADS)
k_runs_partial (uint64_t ntiles, RunWork wk, unsigned long long* __restrict__ part)
	{
	__shared__ unsigned long long s_h[32], s_t[32];
	const uint64_t i = (uint64_t) blockIdx.x * RUN_OFF_TILES + threadIdx.x * 4;
	unsigned int a[4], b[4];
	run_load4 (wk.cntH, i, ntiles, a);  run_load4 (wk.cntT, i, ntiles, b);
	unsigned long long h = (unsigned long long) a[0] + a[1] + a[2] + a[3], t = (unsigned long long) b[0] + b[1] + b[2] + b[3];
	#pragma unroll
	for (int d = 16; d >= 1; d >>= 1) { h += __shfl_xor_sync (0xffffffffu, h, d);  t += __shfl_xor_sync (0xffffffffu, t, d); }
	if ((threadIdx.x & 31) == 0) { s_h[threadIdx.x >> 5] = h;  s_t[threadIdx.x >> 5] = t; }
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		for (int w = 1; w < 32; w++) { h += s_h[w];  t += s_t[w]; }
		part[2 * blockIdx.x] = h;  part[2 * blockIdx.x + 1] = t;
		}
	}

__global__ void __launch_bounds__(RUN_OFF_THREADS)
k_runs_offsets (uint64_t ntiles, RunWork wk, const unsigned long long* __restrict__ part)
	{
	__shared__ unsigned long long s_h[32], s_t[32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	// totals of the blocks before this one (a few dozen values)
	unsigned long long bh = 0, bt = 0;
	for (unsigned int b = lane; b < blockIdx.x; b += 32) { bh += part[2 * b];  bt += part[2 * b + 1]; }
	#pragma unroll
	for (int d = 16; d >= 1; d >>= 1) { bh += __shfl_xor_sync (0xffffffffu, bh, d);  bt += __shfl_xor_sync (0xffffffffu, bt, d); }
	const uint64_t i = (uint64_t) blockIdx.x * RUN_OFF_TILES + threadIdx.x * 4;
	unsigned int a[4], b[4];
	run_load4 (wk.cntH, i, ntiles, a);  run_load4 (wk.cntT, i, ntiles, b);
	const unsigned long long h = (unsigned long long) a[0] + a[1] + a[2] + a[3], t = (unsigned long long) b[0] + b[1] + b[2] + b[3];
	unsigned long long ih = h, it = t;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const unsigned long long uh = __shfl_up_sync (0xffffffffu, ih, d), ut = __shfl_up_sync (0xffffffffu, it, d);
		if (lane >= d) { ih += uh;  it += ut; }
		}
	if (lane == 31) { s_h[warp] = ih;  s_t[warp] = it; }
	__syncthreads ();
	unsigned long long eh = bh + ih - h, et = bt + it - t;
	for (int w = 0; w < warp; w++) { eh += s_h[w];  et += s_t[w]; }
	for (int k = 0; k < 4; k++)
		if (i + k <= ntiles)                            // entry ntiles = the totals
			{
			wk.offH[i + k] = eh;  wk.offT[i + k] = et;
			eh += a[k];  et += b[k];
			}
	}

// first run of every chromosome piece (and the total behind the last one)
__global__ void __launch_bounds__(256)
k_runs_segfirst (const uint64_t* __restrict__ base, int nseg, RunWork wk)
	{
	const int s = blockIdx.x * 256 + threadIdx.x;
	if (s <= nseg) wk.segFirst[s] = wk.offH[base[s]];
	}

#define RUN_EMIT_THREADS 256
__global__ void __launch_bounds__(RUN_EMIT_THREADS)
k_runs_emit (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
             const double* __restrict__ sig, int collapse,
             uint32_t* __restrict__ oStart, uint32_t* __restrict__ oEnd, double* __restrict__ oVal, uint64_t cap, RunWork wk)
	{
	__shared__ unsigned int s_warp[RUN_EMIT_THREADS / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
		{
		if ((wk.cntH[tile] | wk.cntT[tile]) == 0) continue;                   // block-uniform
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, tile, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * RUN_TILE;
		uint32_t heads = wk.headW[tile * RUN_WORDS + threadIdx.x], tails = wk.tailW[tile * RUN_WORDS + threadIdx.x];
		// exclusive prefix of (heads, tails) per word, packed 16:16 (a tile holds at most 8192 of each)
		const unsigned int cnt = (unsigned) __popc (heads) | ((unsigned) __popc (tails) << 16);
		unsigned int inc = cnt;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			const unsigned int up = __shfl_up_sync (0xffffffffu, inc, d);
			if (lane >= d) inc += up;
			}
		__syncthreads ();                                                      // s_warp of the previous tile has been read
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads ();
		unsigned int ex = inc - cnt;
		for (int w = 0; w < warp; w++) ex += s_warp[w];
		unsigned long long ih = wk.offH[tile] + (ex & 0xffffu), it = wk.offT[tile] + (ex >> 16);
		// walk the set bits in cell order
		uint32_t both = heads | tails;
		while (both)
			{
			const int k = __ffs (both) - 1;
			both &= both - 1;
			const uint32_t c = threadIdx.x * 32 + k;
			const uint32_t coord = sd.pos0 + (uint32_t) (t0 + c - sd.lo);
			if (heads & (1u << k))
				{
				if (ih < cap)
					{
					double v = __ldg (sig + t0 + c);
					// the reference's state machine starts with val=+0.0 (genodsp.c:1590): a collapsed
					// run of zeros that begins at base 0 reports that +0.0, not v[0]
					if (coord == 0 && collapse && v == 0) v = 0.0;
					oStart[ih] = coord;
					oVal[ih]   = v;
					}
				ih++;
				}
			// the k-th tail closes the k-th head
			if (tails & (1u << k))
				{
				if (it < cap) oEnd[it] = coord + 1;
				it++;
				}
			}
		}
	}

extern "C" int gdsp_runs (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, int collapse, int showUncovered,
                          uint32_t* d_start, uint32_t* d_end, double* d_val, uint64_t cap,
                          uint64_t* h_n_runs, uint64_t* h_seg_first)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_n_runs, "gdsp_runs: NULL argument");
	GDSP_REQUIRE (cap == 0 || (d_start && d_end && d_val), "gdsp_runs: NULL output arrays");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_runs");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, RUN_TILE, &tm));
	if (tm.ntiles == 0) { *h_n_runs = 0;  if (h_seg_first) for (int s = 0; s <= L->nseg; s++) h_seg_first[s] = 0;  return GDSP_OK; }
	// workspace: flag words, per-tile counts and offsets, per-piece first run
	const size_t wordsB = (size_t) tm.ntiles * RUN_WORDS * sizeof (uint32_t);
	const size_t cntB   = (((size_t) tm.ntiles * sizeof (unsigned int)) + 255) / 256 * 256;
	const size_t offB   = (((size_t) (tm.ntiles + 1) * sizeof (unsigned long long)) + 255) / 256 * 256;
	const size_t segB   = (((size_t) (L->nseg + 1) * sizeof (unsigned long long)) + 255) / 256 * 256;
	const size_t partB  = ((((size_t) tm.ntiles + 1) / RUN_OFF_TILES + 1) * 2 * sizeof (unsigned long long) + 255) / 256 * 256;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 0, 2 * wordsB + 2 * cntB + 2 * offB + segB + partB, &ws));
	RunWork wk;
	char* p = (char*) ws;
	wk.cntH = (unsigned int*) p;            p += cntB;
	wk.cntT = (unsigned int*) p;            p += cntB;
	wk.offH = (unsigned long long*) p;      p += offB;
	wk.offT = (unsigned long long*) p;      p += offB;
	wk.segFirst = (unsigned long long*) p;  p += segB;
	unsigned long long* part = (unsigned long long*) p;  p += partB;
	wk.headW = (uint32_t*) p;               p += wordsB;
	wk.tailW = (uint32_t*) p;
	void* wseg = wk.segFirst;
	GDSP_CUDA (cudaMemsetAsync (wk.cntH, 0, 2 * cntB, c->stream));
	k_runs_flags<<<(unsigned) (2 * tm.ntiles), RUN_FLAG_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, collapse ? 1 : 0,
	        showUncovered == 1 ? 1 : 0, wk);
	GDSP_KERNEL_CHECK ();
	const unsigned nob = (unsigned) ((tm.ntiles + 1 + RUN_OFF_TILES - 1) / RUN_OFF_TILES);      // ntiles + 1 offsets
	k_runs_partial<<<nob, RUN_OFF_THREADS, 0, c->stream>>> (tm.ntiles, wk, part);
	GDSP_KERNEL_CHECK ();
	k_runs_offsets<<<nob, RUN_OFF_THREADS, 0, c->stream>>> (tm.ntiles, wk, part);
	GDSP_KERNEL_CHECK ();
	k_runs_segfirst<<<(L->nseg + 1 + 255) / 256, 256, 0, c->stream>>> (tm.d_base, L->nseg, wk);
	GDSP_KERNEL_CHECK ();
	if (cap > 0)
		{
		uint64_t grid = (uint64_t) c->sm_count * 16;
		if (grid > tm.ntiles) grid = tm.ntiles;
		k_runs_emit<<<(unsigned) grid, RUN_EMIT_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, collapse ? 1 : 0,
		        d_start, d_end, d_val, cap, wk);
		GDSP_KERNEL_CHECK ();
		}
	std::vector<unsigned long long> sf (L->nseg + 1);
	GDSP_CUDA (cudaMemcpyAsync (sf.data (), wseg, sizeof (unsigned long long) * (L->nseg + 1), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	*h_n_runs = sf[L->nseg];
	if (h_seg_first) for (int s = 0; s <= L->nseg; s++) h_seg_first[s] = sf[s];
	if (sf[L->nseg] > cap)
		{
		gdsp_set_error ("gdsp_runs: %llu runs do not fit the caller's capacity of %llu",
		                (unsigned long long) sf[L->nseg], (unsigned long long) cap);
		return GDSP_ERR_CAPACITY;
		}
	return GDSP_OK;
	}
