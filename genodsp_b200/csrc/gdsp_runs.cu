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
// run; one exclusive scan of the head flags (single pass, decoupled look-back
// across all tiles of the launch) gives every record its slot.
// Algorithmic bytes: 8 B/bp read + 16 B/run written.
#include "gdsp_common.cuh"
#include "gdsp_scan.cuh"

#define RUN_THREADS 512
#define RUN_PER     16
#define RUN_TILE    (RUN_THREADS * RUN_PER)      // 8192

__global__ void __launch_bounds__(RUN_THREADS, 3)
k_runs (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
        const double* __restrict__ sig, int collapse, int show,
        uint32_t* __restrict__ oStart, uint32_t* __restrict__ oEnd, double* __restrict__ oVal,
        uint64_t cap, unsigned long long* __restrict__ segFirst, uint64_t ntiles,
        ScanStatus<unsigned long long> st)
	{
	__shared__ unsigned int s_warp[RUN_THREADS / 32];
	__shared__ unsigned long long s_excl;

	const uint32_t tile = scan_take_ticket (st.ticket);
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, tile, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * RUN_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < RUN_TILE) ? (sd.hi - t0) : RUN_TILE);

	const uint32_t c0 = threadIdx.x * RUN_PER;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t heads = 0, tails = 0;
	double   myv[RUN_PER];                    // FULL tiles keep their cells in registers (no staging)
	const bool fullTile = (n == RUN_TILE);

	if (fullTile)
		{
		// 16 consecutive cells per thread straight from global memory (four 256-bit loads); the cell
		// before and after the strip come from the neighbouring lanes, across warps through shared
		// memory, across the tile edge from global memory
		const double* p = sig + t0 + c0;
		#pragma unroll
		for (int q = 0; q < RUN_PER / 4; q++) ldg_stream4 (p + 4 * q, myv[4*q], myv[4*q+1], myv[4*q+2], myv[4*q+3]);
		__shared__ double s_first[RUN_THREADS / 32], s_last[RUN_THREADS / 32];
		__shared__ double s_edge[2];
		if (lane == 0)  s_first[warp] = myv[0];
		if (lane == 31) s_last[warp]  = myv[RUN_PER - 1];
		const bool segFirstCell = (t0 == sd.lo), segLastCell = (t0 + RUN_TILE == sd.hi);
		if (threadIdx.x == 0) s_edge[0] = segFirstCell ? 0.0 : sig[t0 - 1];
		if (threadIdx.x == 1) s_edge[1] = segLastCell  ? 0.0 : sig[t0 + RUN_TILE];
		__syncthreads ();
		double before = shfl_up_f64 (myv[RUN_PER - 1], 1), after = shfl_down_f64 (myv[0], 1);
		if (lane == 0)  before = (warp == 0) ? s_edge[0] : s_last[warp - 1];
		if (lane == 31) after  = (warp == RUN_THREADS / 32 - 1) ? s_edge[1] : s_first[warp + 1];
		const bool isFirst = segFirstCell && threadIdx.x == 0;                 // cell 0 of the chromosome piece
		const bool isLast  = segLastCell  && threadIdx.x == RUN_THREADS - 1;   // its last cell

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
	else
		{
		// edge tile (shorter than RUN_TILE): cells and their neighbours straight from global memory
		if (c0 < n)
			{
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
		}

	// block exclusive scan of head counts
	unsigned int cnt = __popc (heads), inc = cnt;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		unsigned int up = __shfl_up_sync (0xffffffffu, inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_warp[warp] = inc;
	__syncthreads ();
	unsigned int warpExcl = 0, tileTot = 0;
	#pragma unroll
	for (int w = 0; w < RUN_THREADS / 32; w++)
		{
		unsigned int t = s_warp[w];
		if (w < warp) warpExcl += t;
		tileTot += t;
		}
	if (threadIdx.x < 32)
		{
		unsigned long long ex = scan_lookback<unsigned long long> (st, tile, tile == 0, (unsigned long long) tileTot, 0ull,
		                            [] (unsigned long long a, unsigned long long b) { return a + b; });
		if (threadIdx.x == 0)
			{
			s_excl = ex;
			if (tis == 0) segFirst[seg] = ex;
			if (tile == ntiles - 1) segFirst[nseg] = ex + tileTot;
			}
		}
	__syncthreads ();

	unsigned long long idx = s_excl + warpExcl + (inc - cnt);      // slot of this thread's first head
	if (c0 < n && (heads | tails))
		{
		// walk the set bits in cell order; a head and a tail on the same cell: head first
		uint32_t both = heads | tails;
		while (both)
			{
			const int k = __ffs (both) - 1;
			both &= both - 1;
			const uint32_t c = c0 + k;
			const uint32_t coord = sd.pos0 + (uint32_t) (t0 + c - sd.lo);
			if (heads & (1u << k))
				{
				if (idx < cap)
					{
					double v = __ldg (sig + t0 + c);         // re-read (L2): keeping 16 cells per thread in
					                                           // registers until here halves the occupancy
					// the reference's state machine starts with val=+0.0 (genodsp.c:1590): a collapsed
					// run of zeros that begins at base 0 reports that +0.0, not v[0]
					if (coord == 0 && collapse && v == 0) v = 0.0;
					oStart[idx] = coord;
					oVal[idx]   = v;
					}
				idx++;
				}
			// the run a tail closes is the most recent head: slot idx-1
			if ((tails & (1u << k)) && idx - 1 < cap) oEnd[idx - 1] = coord + 1;
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
	void* ws;  void* wseg;
	GDSP_TRY (gdsp_ws (c, 0, scan_status_bytes<unsigned long long> (tm.ntiles), &ws));
	GDSP_TRY (gdsp_ws (c, 2, sizeof (unsigned long long) * (L->nseg + 1), &wseg));
	ScanStatus<unsigned long long> st = scan_status_carve<unsigned long long> (ws, tm.ntiles);
	GDSP_CUDA (cudaMemsetAsync (ws, 0, scan_status_clear_bytes<unsigned long long> (tm.ntiles), c->stream));
	k_runs<<<(unsigned) tm.ntiles, RUN_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, collapse ? 1 : 0,
	        showUncovered == 1 ? 1 : 0, d_start, d_end, d_val, cap, (unsigned long long*) wseg, tm.ntiles, st);
	GDSP_KERNEL_CHECK ();
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
