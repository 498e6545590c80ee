// gdsp_morph.cu -- run-length morphology on the thresholded signal.
//
// Replaces op_close_apply (morphology.c:231-319), op_open_apply (:529-605),
// op_dilate_apply (:882-1072) and op_erode_apply (:1331-1454).  The reference
// walks each chromosome with a run-length state machine; here every output cell
// is a function of the nearest "marker" cell on either side:
//
//   close / dilate : markers = cells of the set S          (S = !(v<=T))
//   open  / erode  : markers = cells NOT in the set S       (S =  (v>T))
//
//   K1 k_morph_pack   reads the signal once (8 B/bp), writes 1 marker bit per
//                     cell (warp ballots) and per-tile first/last marker
//   K2 k_morph_apply  reads only the bit words (+ the per-tile summaries of a
//                     bounded number of neighbouring tiles) and writes one/zero
//                     (8 B/bp).  Distances are exact for any length L: the
//                     in-word neighbour comes from clz/ffs, the in-tile one from
//                     a 256-word shuffle scan, the out-of-tile one from the
//                     neighbouring tiles' summaries.
// Algorithmic bytes: 16 B/bp (+ 0.16 B/bp of bit traffic).
#include "gdsp_common.cuh"
#include <algorithm>
#include <vector>

#define MO_TILE    8192
#define MO_WORDS   (MO_TILE / 32)        // 256
#define MO_THREADS 256

struct MorphWork
	{
	uint32_t* words;        // one bit per buffer cell (word = cell >> 5)
	long long* tileFirst;   // per tile: chromosome coordinate of first marker, -1 if none
	long long* tileLast;
	};

// kind: marker definition
//   0 close : !(v<=T)        1 open : !(v>T)       2 dilate : first cell (v>T), others !(v<=T)
//   3 erode : !(v>T)
__global__ void __launch_bounds__(MO_THREADS)
k_morph_pack (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
              const double* __restrict__ sig, int kind, double T, MorphWork wk)
	{
	__shared__ long long s_first[MO_THREADS / 32], s_last[MO_THREADS / 32];
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * MO_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < MO_TILE) ? (sd.hi - t0) : MO_TILE);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const bool endsChrom = ((uint64_t) sd.pos0 + (sd.hi - sd.lo) == (uint64_t) sd.chromLen);
	const bool complement = (kind == GDSP_MORPH_OPEN || kind == GDSP_MORPH_ERODE);

	// thread t owns word t of the tile: 32 consecutive cells (256 bytes) read with 256-bit loads, the
	// marker bits assembled in a register -- no warp votes, and the word store is coalesced
	long long first = -1, last = -1;
	const uint32_t c0 = threadIdx.x * 32;
	const uint32_t nRound = (n + 31u) & ~31u;
	uint32_t word = 0;
	if (c0 + 32 <= n)
		{
		const double* p = sig + t0 + c0;
		#pragma unroll
		for (int h = 0; h < 2; h++)
			{
			double v[16];
			#pragma unroll
			for (int q = 0; q < 4; q++) ldg_stream4 (p + 16 * h + 4 * q, v[4*q], v[4*q+1], v[4*q+2], v[4*q+3]);
			#pragma unroll
			for (int k = 0; k < 16; k++)
				{
				const bool mark = complement ? !(v[k] > T) : !(v[k] <= T);
				word |= (mark ? 1u : 0u) << (16 * h + k);
				}
			}
		if (kind == GDSP_MORPH_DILATE && sd.pos0 == 0 && t0 == sd.lo && threadIdx.x == 0)
			{
			const double v0 = __ldg (sig + t0);                      // the reference tests the first cell with v > T
			word = (word & ~1u) | ((v0 > T) ? 1u : 0u);
			}
		}
	else if (c0 < nRound)
		{
		for (uint32_t k = 0; k < 32; k++)
			{
			const uint32_t c = c0 + k;
			bool mark = false;
			if (c < n)
				{
				const double v = __ldg (sig + t0 + c);
				if      (kind == GDSP_MORPH_CLOSE)  mark = !(v <= T);
				else if (kind == GDSP_MORPH_DILATE) mark = (sd.pos0 == 0 && t0 + c == sd.lo) ? (v > T) : !(v <= T);
				else                                mark = !(v > T);
				}
			else if (complement && endsChrom) mark = true;        // cells past the chromosome end bound every run
			word |= (mark ? 1u : 0u) << k;
			}
		}
	if (c0 < nRound)
		{
		wk.words[(t0 >> 5) + threadIdx.x] = word;
		if (word != 0)
			{
			const long long cc = (long long) sd.pos0 + (long long) (t0 - sd.lo) + (long long) c0;
			first = cc + (__ffs (word) - 1);
			last  = cc + (31 - __clz (word));
			}
		}
	// first marker = smallest, last = largest over the warp
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		{
		const long long f2 = __shfl_xor_sync (0xffffffffu, first, d), l2 = __shfl_xor_sync (0xffffffffu, last, d);
		if (f2 >= 0 && (first < 0 || f2 < first)) first = f2;
		if (l2 > last) last = l2;
		}
	if (lane == 0) { s_first[warp] = first;  s_last[warp] = last; }
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		long long f = -1, l = -1;
		for (int w = 0; w < MO_THREADS / 32; w++)
			{
			if (s_first[w] >= 0 && (f < 0 || s_first[w] < f)) f = s_first[w];
			if (s_last[w] > l) l = s_last[w];
			}
		wk.tileFirst[blockIdx.x] = f;
		wk.tileLast[blockIdx.x]  = l;
		}
	}

// Slab pieces: marker bits of the readable halo cells [dlo,lo) and [hi,dhi) of every piece, so that
// the search for the nearest marker can continue past the piece into its neighbour's data.
// blockIdx.x = 2*segment + side.  Whole halo words are stored; the word shared with the owned tail
// (hi not a multiple of 32) is OR-ed.
__global__ void __launch_bounds__(256)
k_morph_pack_halo (const SegDev* __restrict__ segs, const double* __restrict__ sig, int kind, double T, MorphWork wk)
	{
	const SegDev sd = segs[blockIdx.x >> 1];
	const bool right = (blockIdx.x & 1) != 0;
	const uint64_t a = right ? sd.hi : sd.dlo, b = right ? sd.dhi : sd.lo;       // halo cells [a,b)
	if (a >= b) return;
	const bool complement = (kind == GDSP_MORPH_OPEN || kind == GDSP_MORPH_ERODE);
	const int lane = threadIdx.x & 31;
	const uint64_t w0 = a >> 5, w1 = (b + 31) >> 5;
	for (uint64_t w = w0 + (threadIdx.x >> 5); w < w1; w += 256 / 32)
		{
		const uint64_t cell = w * 32 + lane;
		bool mark = false;
		if (cell >= a && cell < b)
			{
			const double v = __ldg (sig + cell);
			mark = complement ? !(v > T) : !(v <= T);
			}
		const uint32_t word = __ballot_sync (0xffffffffu, mark);
		if (lane == 0)
			{
			if (word) atomicOr (wk.words + w, word);             // the array was cleared; a word may hold owned cells too
			}
		}
	}

// bits lo .. hi-1 of a word (0 <= lo, hi <= 32; empty when hi <= lo)
__device__ __forceinline__ uint32_t mo_range (int lo, int hi)
	{
	const uint32_t mh = (hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u);
	const uint32_t ml = (lo >= 32) ? 0xffffffffu : ((1u << lo) - 1u);
	return mh & ~ml;
	}
__device__ __forceinline__ int mo_clamp32 (long long x) { return (x < 0) ? 0 : ((x > 32) ? 32 : (int) x); }

__global__ void __launch_bounds__(MO_THREADS)
k_morph_apply (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
               double* __restrict__ sig, int kind, double L, long long left, long long right,
               double oneVal, double zeroVal, uint32_t maxTileSearch, MorphWork wk)
	{
	__shared__ uint32_t s_word[MO_WORDS];
	__shared__ int      s_prevW[MO_WORDS], s_nextW[MO_WORDS];
	__shared__ int      s_wtot[MO_THREADS / 32];
	__shared__ long long s_carry[2];

	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * MO_TILE;
	const uint32_t n  = (uint32_t) ((sd.hi - t0 < MO_TILE) ? (sd.hi - t0) : MO_TILE);
	const uint32_t nw = (n + 31) >> 5;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const long long coord0 = (long long) sd.pos0 + (long long) (t0 - sd.lo);      // chromosome coordinate of tile cell 0
	const uint64_t tilesInSeg = base[seg + 1] - base[seg];

	// --- tile words and nearest non-empty word on either side (256-entry shuffle scans) ---
	const uint32_t myWord = (threadIdx.x < nw) ? wk.words[(t0 >> 5) + threadIdx.x] : 0u;
	s_word[threadIdx.x] = myWord;
	int pv = myWord ? (int) threadIdx.x : -1;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		int up = __shfl_up_sync (0xffffffffu, pv, d);
		if (lane >= d && up > pv) pv = up;
		}
	if (lane == 31) s_wtot[warp] = pv;
	__syncthreads ();
	{
	int carry = -1;
	for (int w = 0; w < warp; w++) if (s_wtot[w] > carry) carry = s_wtot[w];
	s_prevW[threadIdx.x] = (pv > carry) ? pv : carry;
	}
	__syncthreads ();
	int nx = myWord ? (int) threadIdx.x : 0x7fffffff;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		int dn = __shfl_down_sync (0xffffffffu, nx, d);
		if (lane + d < 32 && dn < nx) nx = dn;
		}
	if (lane == 0) s_wtot[warp] = nx;
	__syncthreads ();
	{
	int carry = 0x7fffffff;
	for (int w = MO_THREADS / 32 - 1; w > warp; w--) if (s_wtot[w] < carry) carry = s_wtot[w];
	int r = (nx < carry) ? nx : carry;
	s_nextW[threadIdx.x] = (r == 0x7fffffff) ? -1 : r;
	}

	// --- nearest marker outside the tile: neighbouring tiles' summaries, bounded search ---
	if (warp == 0)
		{
		long long found = -1;
		for (uint64_t off = 1; off <= tis && off <= maxTileSearch && found < 0; off += 32)
			{
			uint64_t k = off + lane;
			long long v = (k <= tis && k <= maxTileSearch) ? wk.tileLast[blockIdx.x - k] : -1;
			unsigned m = __ballot_sync (0xffffffffu, v >= 0);
			if (m) found = __shfl_sync (0xffffffffu, v, __ffs (m) - 1);
			}
		// ran out of tiles of this piece: the piece's left halo (slab layouts) holds the neighbour's cells
		if (found < 0 && tis < (uint64_t) maxTileSearch && sd.dlo < sd.lo)
			{
			const long long wTop = (long long) (sd.lo >> 5) - 1, wBot = (long long) (sd.dlo >> 5);
			for (long long wb = wTop; wb >= wBot && found < 0; wb -= 32)
				{
				const long long wi = wb - lane;
				const uint32_t word = (wi >= wBot) ? wk.words[wi] : 0u;
				const unsigned m = __ballot_sync (0xffffffffu, word != 0u);
				if (m)
					{
					const int src = __ffs (m) - 1;
					const uint32_t wsel = __shfl_sync (0xffffffffu, word, src);
					const long long cell = (wb - src) * 32 + (31 - __clz (wsel));
					found = (long long) sd.pos0 - ((long long) sd.lo - cell);
					}
				}
			}
		if (lane == 0) s_carry[0] = found;
		}
	else if (warp == 1)
		{
		long long found = -1;
		const uint64_t after = tilesInSeg - 1 - tis;
		for (uint64_t off = 1; off <= after && off <= maxTileSearch && found < 0; off += 32)
			{
			uint64_t k = off + lane;
			long long v = (k <= after && k <= maxTileSearch) ? wk.tileFirst[blockIdx.x + k] : -1;
			unsigned m = __ballot_sync (0xffffffffu, v >= 0);
			if (m) found = __shfl_sync (0xffffffffu, v, __ffs (m) - 1);
			}
		if (found < 0 && after < (uint64_t) maxTileSearch && sd.dhi > sd.hi)
			{
			const long long wBot = (long long) ((sd.hi + 31) >> 5), wTop = (long long) ((sd.dhi + 31) >> 5);   // [wBot, wTop)
			for (long long wb = wBot; wb < wTop && found < 0; wb += 32)
				{
				const long long wi = wb + lane;
				const uint32_t word = (wi < wTop) ? wk.words[wi] : 0u;
				const unsigned m = __ballot_sync (0xffffffffu, word != 0u);
				if (m)
					{
					const int src = __ffs (m) - 1;
					const uint32_t wsel = __shfl_sync (0xffffffffu, word, src);
					const long long cell = (wb + src) * 32 + (__ffs (wsel) - 1);
					found = (long long) sd.pos0 + (cell - (long long) sd.lo);
					}
				}
			}
		if (lane == 0) s_carry[1] = found;
		}
	__syncthreads ();
	const long long carryPrev = s_carry[0], carryNext = s_carry[1];
	const long long chromEnd  = (long long) sd.chromLen;

	// thread t decides the 32 cells of word t: the nearest markers outside the word are looked up once,
	// the ones inside come from clz/ffs on the word; 256 bytes of results leave as 256-bit stores
	const uint32_t w = threadIdx.x;
	const uint32_t c0 = w * 32;
	if (c0 >= n) return;
	const uint32_t word = s_word[w];
	long long prevOut, nextOut;                       // nearest marker before / after this word (chromosome coordinates)
	{
	const int pw = (w > 0) ? s_prevW[w - 1] : -1;
	prevOut = (pw >= 0) ? coord0 + (long long) (pw * 32 + 31 - __clz (s_word[pw])) : carryPrev;
	const int nwd = (w + 1 < MO_WORDS) ? s_nextW[w + 1] : -1;
	nextOut = (nwd >= 0) ? coord0 + (long long) (nwd * 32 + __ffs (s_word[nwd]) - 1) : carryNext;
	}
	const long long cw = coord0 + c0;                 // coordinate of the word's first cell
	double* o = sig + t0 + c0;
	// a run length n is an integer: n > L  <=>  n > floor(L), and the compare stays on the integer pipe (the
	// s64 -> f64 conversion per cell was the most expensive instruction of this loop)
	const long long Lf = (L >= 9.0e18) ? 0x7fffffffffffffffll : ((L <= -9.0e18) ? -0x7fffffffffffffffll : (long long) floor (L));
	// Word-wise decision.  When the operator's length is at least 31 cells, what happens to a cell depends only on
	// which GAP between markers it lies in: gaps inside the word are shorter than the length (close fills them, open /
	// erode clear them, dilate covers them), and the two gaps that leave the word are settled by the nearest markers
	// outside it -- a handful of integer instructions per word instead of ~35 per cell (clz/ffs, 64-bit selects and
	// compares for every cell): ncu had this kernel at 1.9 warp instructions per cell, 82 % issue, writing 8 B/bp.
	const bool wordwise = (kind == GDSP_MORPH_CLOSE || kind == GDSP_MORPH_OPEN) ? (Lf >= 31) : (left >= 31 && right >= 31);
	if (wordwise)
		{
		const int first = word ? __ffs (word) - 1 : 32, last = word ? 31 - __clz (word) : -1;
		uint32_t bits = 0;
		if (kind == GDSP_MORPH_CLOSE)
			{
			if (word)
				{
				bits = mo_range (first, last + 1);
				if (first > 0 && prevOut >= 0 && !(cw + first - prevOut - 1 > Lf)) bits |= mo_range (0, first);
				if (last < 31 && nextOut >= 0 && !(nextOut - (cw + last) - 1 > Lf)) bits |= mo_range (last + 1, 32);
				}
			else if (prevOut >= 0 && nextOut >= 0 && !(nextOut - prevOut - 1 > Lf)) bits = 0xffffffffu;
			}
		else if (kind == GDSP_MORPH_OPEN)
			{
			const long long re = (nextOut >= 0) ? nextOut : chromEnd;
			if (word)
				{
				if (first > 0 && cw + first - (prevOut + 1) > Lf) bits |= mo_range (0, first);
				if (last < 31 && re - (cw + last + 1) > Lf) bits |= mo_range (last + 1, 32);
				}
			else if (re - (prevOut + 1) > Lf) bits = 0xffffffffu;
			}
		else if (kind == GDSP_MORPH_DILATE)
			{
			if (word) bits = 0xffffffffu;
			else
				{
				if (prevOut >= 0) bits |= mo_range (0, mo_clamp32 (prevOut + right + 1 - cw));
				if (nextOut >= 0) bits |= mo_range (mo_clamp32 (nextOut - left - cw), 32);
				}
			}
		else if (!word)                                                         // erode: a word that holds a marker keeps nothing
			{
			const long long rs = prevOut + 1, re = (nextOut >= 0) ? nextOut : chromEnd;
			bits = mo_range (mo_clamp32 (rs + right - cw), mo_clamp32 (re - left - cw));
			}
		if (c0 + 32 <= n)
			{
			#pragma unroll
			for (int g = 0; g < 8; g++)
				stg_stream4 (o + 4 * g, (bits >> (4 * g)) & 1u ? oneVal : zeroVal, (bits >> (4 * g + 1)) & 1u ? oneVal : zeroVal,
				             (bits >> (4 * g + 2)) & 1u ? oneVal : zeroVal, (bits >> (4 * g + 3)) & 1u ? oneVal : zeroVal);
			}
		else
			for (uint32_t q = 0; c0 + q < n; q++) o[q] = (bits >> q) & 1u ? oneVal : zeroVal;
		return;
		}
	#pragma unroll
	for (int g = 0; g < 8; g++)
		{
		double y[4];
		#pragma unroll
		for (int q = 0; q < 4; q++)
			{
			const int bit = 4 * g + q;
			const long long cp = cw + bit;
			uint32_t m = word & (0xffffffffu >> (31 - bit));
			const long long prevM = m ? cw + (31 - __clz (m)) : prevOut;
			m = word & (0xffffffffu << bit);
			const long long nextM = m ? cw + (__ffs (m) - 1) : nextOut;
			bool one;
			if (kind == GDSP_MORPH_DILATE)
				one = (prevM >= 0 && cp - prevM <= right) || (nextM >= 0 && nextM - cp <= left);
			else if (kind == GDSP_MORPH_CLOSE)
				one = (prevM == cp) || (prevM >= 0 && nextM >= 0 && !(nextM - prevM - 1 > Lf));
			else
				{
				// markers are the cells outside the set; a marker at cp means cp itself is outside
				if (prevM == cp) one = false;
				else
					{
					const long long rs = prevM + 1;                          // prevM == -1 -> run starts at coordinate 0
					const long long re = (nextM >= 0) ? nextM : chromEnd;
					if (kind == GDSP_MORPH_OPEN) one = (re - rs > Lf);
					else                         one = (cp >= rs + right) && (cp < re - left);
					}
				}
			y[q] = one ? oneVal : zeroVal;
			}
		const uint32_t c = c0 + 4 * g;
		if (c + 4 <= n) stg_stream4 (o + 4 * g, y[0], y[1], y[2], y[3]);
		else
			{
			#pragma unroll
			for (int q = 0; q < 4; q++) if (c + q < n) o[4 * g + q] = y[q];
			}
		}
	}

extern "C" size_t gdsp_morph_work_bytes (uint64_t buffer_cells)
	{
	uint64_t words = (buffer_cells + 31) / 32 + 64;
	uint64_t tiles = buffer_cells / MO_TILE + 4096;       // every segment adds at most one partial tile
	return (size_t) (words * 4 + 256 + 2 * (tiles * 8 + 256));
	}

extern "C" int gdsp_morphology (gdsp_ctx* c, const gdsp_layout* L_, double* sig, uint64_t buffer_cells, void* work,
                                int kind, double length, uint32_t left, uint32_t right,
                                double threshold, double oneVal, double zeroVal)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && work, "gdsp_morphology: NULL argument");
	GDSP_REQUIRE (kind >= GDSP_MORPH_CLOSE && kind <= GDSP_MORPH_ERODE, "gdsp_morphology: bad kind %d", kind);
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_morphology");
	// how far can a neighbouring marker matter?  (cells)
	double reach;
	if      (kind == GDSP_MORPH_DILATE || kind == GDSP_MORPH_ERODE) reach = (double) ((left > right) ? left : right) + 1;
	else    reach = length + 2;
	// slab pieces: the readable halo on a side where the chromosome continues must cover that reach
	bool anyHalo = false;
	for (int s = 0; s < L->nseg; s++)
		{
		const gdsp_seg& g = L->h[s];
		const bool startsChrom = (g.pos0 == 0), endsChrom = ((uint64_t) g.pos0 + (g.hi - g.lo) == (uint64_t) g.chrom_len);
		GDSP_REQUIRE (startsChrom || (double) (g.lo - g.dlo) >= reach || (uint64_t) (g.lo - g.dlo) == (uint64_t) g.pos0,
		              "gdsp_morphology: piece %d needs a left halo of %.0f cells (has %llu)", s, reach, (unsigned long long) (g.lo - g.dlo));
		GDSP_REQUIRE (endsChrom || (double) (g.dhi - g.hi) >= reach
		              || (uint64_t) (g.dhi - g.hi) == (uint64_t) g.chrom_len - ((uint64_t) g.pos0 + (g.hi - g.lo)),
		              "gdsp_morphology: piece %d needs a right halo of %.0f cells (has %llu)", s, reach, (unsigned long long) (g.dhi - g.hi));
		if (g.dlo < g.lo || g.dhi > g.hi) anyHalo = true;
		}
	if (anyHalo)
		{
		// the marker bits of two pieces must not meet in one 32-cell word
		std::vector<std::pair<uint64_t, uint64_t> > span;
		for (int s = 0; s < L->nseg; s++) span.push_back (std::make_pair ((uint64_t) L->h[s].dlo, (uint64_t) L->h[s].dhi));
		std::sort (span.begin (), span.end ());
		for (size_t k = 1; k < span.size (); k++)
			GDSP_REQUIRE ((span[k].first >> 5) >= ((span[k-1].second + 31) >> 5),
			              "gdsp_morphology: pieces %zu and %zu share a 32-cell word; leave 32 spare cells between pieces", k - 1, k);
		}
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, MO_TILE, &tm));
	uint64_t words = (buffer_cells + 31) / 32 + 64;
	uint64_t tilesCap = buffer_cells / MO_TILE + 4096;
	GDSP_REQUIRE (tm.ntiles <= tilesCap, "gdsp_morphology: layout has more tiles than the work buffer allows");
	MorphWork wk;
	char* p = (char*) work;
	wk.words = (uint32_t*) p;                  p += ((words * 4 + 255) / 256) * 256;
	wk.tileFirst = (long long*) p;             p += ((tilesCap * 8 + 255) / 256) * 256;
	wk.tileLast  = (long long*) p;

	double tilesD = reach / MO_TILE + 2;
	uint32_t maxTileSearch = (tilesD > 4.0e9) ? 0xffffffffu : (uint32_t) tilesD;

	if (anyHalo) GDSP_CUDA (cudaMemsetAsync (wk.words, 0, words * 4, c->stream));     // halo words are OR-ed in
	k_morph_pack<<<(unsigned) tm.ntiles, MO_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, kind, threshold, wk);
	GDSP_KERNEL_CHECK ();
	if (anyHalo)
		{
		k_morph_pack_halo<<<2 * L->nseg, 256, 0, c->stream>>> (L->d, sig, kind, threshold, wk);
		GDSP_KERNEL_CHECK ();
		}
	k_morph_apply<<<(unsigned) tm.ntiles, MO_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, kind, length,
	        (long long) left, (long long) right, oneVal, zeroVal, maxTileSearch, wk);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
