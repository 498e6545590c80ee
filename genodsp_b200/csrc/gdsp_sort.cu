// gdsp_sort.cu -- percentile selection and the percentile operator's post-state.
//
// Replaces op_percentile_apply (percentile.c:392-751).  The reference collects
// the qualifying samples, qsorts every chromosome and bubble-merges the
// chromosomes pairwise (percentile.c:611-651, O(m^2 n log n)); here
//
//   gdsp_percentiles : exact order statistics WITHOUT sorting the genome.
//       (1) a hashed sample of the qualifying cells is sorted and gives, for
//           every requested percentile, a value window [lo,hi] that contains
//           the wanted rank with overwhelming probability;
//       (2) ONE streaming pass (8 B/bp) counts the cells in every region
//           delimited by the window bounds and compacts the (few) cells that
//           fall strictly inside a window;
//       (3) the wanted rank is located from the exact region counts; it is
//           either a window bound itself (heavy ties) or an element of the
//           small sorted candidate set.  If a rank falls outside its window the
//           pass is repeated on the exact bracket learned from the counts, so
//           the result is always exact.
//   gdsp_sort_genome : LSD radix sort (8-bit digits, stable, single pass per
//       digit with decoupled look-back; digits on which all keys agree are
//       skipped) -- the globally sorted genome the reference leaves behind.
#include <algorithm>
#include <cmath>
#include "gdsp_common.cuh"

#define SORT_THREADS 256
#define SORT_WARPS   8
#define SORT_ROUNDS  16
#define SORT_TILE    (SORT_WARPS * SORT_ROUNDS * 32)     // 4096 keys

// output position -> buffer cell (segmented destination), or identity
struct OutMap
	{
	int             nseg;          // 0 = linear
	const uint64_t* prefix;        // nseg+1: owned cells before segment s
	const SegDev*   segs;
	};

__device__ __forceinline__ uint64_t out_cell (const OutMap& om, uint64_t pos)
	{
	if (om.nseg == 0) return pos;
	int lo = 0, hi = om.nseg - 1;
	while (lo < hi)
		{
		int mid = (lo + hi + 1) >> 1;
		if (om.prefix[mid] <= pos) lo = mid; else hi = mid - 1;
		}
	return om.segs[lo].lo + (pos - om.prefix[lo]);
	}

// ---------------------------------------------------------------------------
// bitwise OR / AND of all keys: a digit on which OR and AND agree is constant
// ---------------------------------------------------------------------------

#define ZERO_KEY 0x8000000000000000ull          // f64_key(+0.0)

// res[0] |= keys, res[1] &= keys, res[2] += #(+0.0), res[3] += #(keys below +0.0)
__global__ void __launch_bounds__(256)
k_sort_orand (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
              const double* __restrict__ in, unsigned long long* __restrict__ res)
	{
	unsigned long long o = 0ull, a = ~0ull, nz = 0ull, nneg = 0ull;
	for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
		{
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * SORT_TILE;
		uint64_t t1 = t0 + SORT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
		for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
			{
			unsigned long long k = f64_key (in[i]);
			o |= k;  a &= k;
			nz += (k == ZERO_KEY);  nneg += (k < ZERO_KEY);
			}
		}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		{
		o |= __shfl_xor_sync (0xffffffffu, o, d);
		a &= __shfl_xor_sync (0xffffffffu, a, d);
		nz   += __shfl_xor_sync (0xffffffffu, nz, d);
		nneg += __shfl_xor_sync (0xffffffffu, nneg, d);
		}
	if ((threadIdx.x & 31) == 0)
		{
		atomicOr (&res[0], o);  atomicAnd (&res[1], a);
		if (nz)   atomicAdd (&res[2], nz);
		if (nneg) atomicAdd (&res[3], nneg);
		}
	}

// copy every cell that is not +0.0 to out[] (any order; they get sorted next).
// The survivors are collected in a shared-memory buffer and flushed with ONE global atomic per
// ~1000 cells: with a global atomic per warp vote (the first version) 15 M same-address atomics
// made this 24.7 GB read take 22 ms.
#define CNZ_CAP 5120          // buffer entries; a tile adds at most SORT_TILE = 4096
__global__ void __launch_bounds__(256)
k_sort_compact_nonzero (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
                        const double* __restrict__ in, double* __restrict__ out, unsigned long long* __restrict__ counter)
	{
	__shared__ double s_buf[CNZ_CAP];
	__shared__ unsigned int s_cnt;
	__shared__ unsigned long long s_base;
	const int lane = threadIdx.x & 31;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads ();
	for (uint64_t t = blockIdx.x; ; t += gridDim.x)
		{
		// flush when the next tile might not fit (or at the end)
		const bool more = (t < ntiles);
		if (!more || s_cnt + SORT_TILE > CNZ_CAP)
			{
			const unsigned int cnt = s_cnt;
			if (threadIdx.x == 0 && cnt) s_base = atomicAdd (counter, (unsigned long long) cnt);
			__syncthreads ();
			for (unsigned int i = threadIdx.x; i < cnt; i += 256) out[s_base + i] = s_buf[i];
			__syncthreads ();
			if (threadIdx.x == 0) s_cnt = 0;
			__syncthreads ();
			}
		if (!more) break;
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * SORT_TILE;
		uint64_t t1 = t0 + SORT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
		for (uint64_t i0 = t0; i0 < t1; i0 += 256 * 4)
			{
			double v[4];  bool keep[4];
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const uint64_t i = i0 + (uint64_t) u * 256 + threadIdx.x;
				v[u] = (i < t1) ? in[i] : 0.0;
				keep[u] = (i < t1) && (f64_key (v[u]) != ZERO_KEY);
				}
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const unsigned m = __ballot_sync (0xffffffffu, keep[u]);
				if (m == 0) continue;
				unsigned int b0 = 0;
				if (lane == __ffs (m) - 1) b0 = atomicAdd (&s_cnt, (unsigned int) __popc (m));
				b0 = __shfl_sync (0xffffffffu, b0, __ffs (m) - 1);
				if (keep[u]) s_buf[b0 + __popc (m & ((1u << lane) - 1u))] = v[u];
				}
			}
		__syncthreads ();
		}
	}

// sorted genome = [negatives (sorted)] [+0.0 x nZero] [positives (sorted)]: cell at sorted position p
__global__ void __launch_bounds__(256)
k_sort_fill_with_zeros (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                        const uint64_t* __restrict__ prefix, const double* __restrict__ sorted,
                        unsigned long long nNeg, unsigned long long nZero, double* __restrict__ out)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SORT_TILE;
	uint64_t t1 = t0 + SORT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	const uint64_t p0 = prefix[seg] + (t0 - sd.lo);
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
		{
		const uint64_t p = p0 + (i - t0);
		double v = 0.0;
		if (p < nNeg) v = sorted[p];
		else if (p >= nNeg + nZero) v = sorted[p - nZero];
		out[i] = v;
		}
	}

// ---------------------------------------------------------------------------
// One stable counting-sort pass on an 8-bit digit = three chain-free kernels:
//   k_sort_hist     per-tile digit histograms (warp-private counters ranked with
//                   match.any: no same-address atomics however skewed the data)
//   k_tab_*         exclusive prefix over the digit-major table [digit][tile]:
//                   entry (d,t) becomes the output position of tile t's first key
//                   with digit d
//   k_sort_scatter  re-reads the tile, ranks its keys the same way and writes
//                   them to their final positions
// Traffic: 24 B per key per pass (+ the table, ~1 B per key).
// ---------------------------------------------------------------------------

#define SORT_HB 8      // consecutive tiles handled by one histogram block
#define SORT_SB 4      // consecutive tiles handled by one scatter block

// rank the valid keys of one round inside the warp; returns the digit
__device__ __forceinline__ unsigned sort_rank_round (unsigned long long key, int shift, bool valid, int lane,
                                                     unsigned int* warpCnt, unsigned short& rank)
	{
	const unsigned vm = __ballot_sync (0xffffffffu, valid);
	unsigned d = 0;
	if (valid)
		{
		d = (unsigned) ((key >> shift) & 255ull);
		}
	// lanes holding the same digit, from 8 votes: the cost does not depend on how many distinct digits
	// the warp holds (match.any walks the distinct values: 5x slower on uniformly random digits)
	unsigned peers = vm;
	#pragma unroll
	for (int b = 0; b < 8; b++)
		{
		const unsigned bit = (d >> b) & 1u;
		const unsigned bal = __ballot_sync (0xffffffffu, valid && bit);
		peers &= bit ? bal : ~bal;
		}
	// one full-warp shuffle (every lane reads its own group's leader); a shuffle per peer mask would be
	// issued once per distinct digit
	const int leader = valid ? (__ffs (peers) - 1) : lane;
	unsigned pre = 0;
	if (valid && lane == leader) { pre = warpCnt[d];  warpCnt[d] = pre + __popc (peers); }
	pre = __shfl_sync (0xffffffffu, pre, leader);
	if (valid) rank = (unsigned short) (pre + __popc (peers & ((1u << lane) - 1u)));
	__syncwarp ();
	return d;
	}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
             uint64_t ntilesPad, const double* __restrict__ in, int shift, unsigned int* __restrict__ tileHist)
	{
	__shared__ unsigned int s_cnt[SORT_WARPS][256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t tFirst = (uint64_t) blockIdx.x * SORT_HB;
	unsigned int mine[SORT_HB];
	#pragma unroll
	for (int k = 0; k < SORT_HB; k++)
		{
		mine[k] = 0;
		const uint64_t t = tFirst + k;
		if (t >= ntiles) continue;                           // uniform across the block
		for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
		__syncthreads ();
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * SORT_TILE;
		const uint32_t n  = (uint32_t) ((sd.hi - t0 < SORT_TILE) ? (sd.hi - t0) : SORT_TILE);
		#pragma unroll 4
		for (int r = 0; r < SORT_ROUNDS; r++)
			{
			const uint32_t idx = warp * (SORT_ROUNDS * 32) + r * 32 + lane;
			const bool valid = idx < n;
			const unsigned long long key = valid ? f64_key (in[t0 + idx]) : 0ull;
			unsigned short rk;
			sort_rank_round (key, shift, valid, lane, s_cnt[warp], rk);
			}
		__syncthreads ();
		unsigned int tot = 0;
		#pragma unroll
		for (int w = 0; w < SORT_WARPS; w++) tot += s_cnt[w][threadIdx.x];
		mine[k] = tot;
		__syncthreads ();
		}
	// thread d owns digit d: SORT_HB consecutive table entries = one 32-byte sector
	uint4* dst = reinterpret_cast<uint4*> (tileHist + (uint64_t) threadIdx.x * ntilesPad + tFirst);
	dst[0] = make_uint4 (mine[0], mine[1], mine[2], mine[3]);
	dst[1] = make_uint4 (mine[4], mine[5], mine[6], mine[7]);
	}

// ---- exclusive prefix over the table (u32 counts -> u64 positions) ------------------
#define TAB_CHUNK 8192

__global__ void __launch_bounds__(256)
k_tab_reduce (const unsigned int* __restrict__ tab, uint64_t n, unsigned long long* __restrict__ partial)
	{
	__shared__ unsigned long long s_red[8];
	const uint64_t c0 = (uint64_t) blockIdx.x * TAB_CHUNK;
	unsigned long long t = 0;
	for (uint32_t i = threadIdx.x * 4; i < TAB_CHUNK; i += 256 * 4)
		if (c0 + i < n)
			{
			uint4 v = *reinterpret_cast<const uint4*> (tab + c0 + i);        // n is a multiple of 8
			t += (unsigned long long) v.x + v.y + v.z + v.w;
			}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync (0xffffffffu, t, d);
	if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		unsigned long long a = 0;
		for (int w = 0; w < 8; w++) a += s_red[w];
		partial[blockIdx.x] = a;
		}
	}

// one block: in-place exclusive scan of the chunk totals
__global__ void __launch_bounds__(1024)
k_tab_scan_partials (unsigned long long* __restrict__ partial, uint64_t nchunks)
	{
	__shared__ unsigned long long s_w[32];
	__shared__ unsigned long long s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads ();
	for (uint64_t c0 = 0; c0 < nchunks; c0 += 1024)
		{
		const uint64_t i = c0 + threadIdx.x;
		unsigned long long v = (i < nchunks) ? partial[i] : 0ull, inc = v;
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			unsigned long long up = __shfl_up_sync (0xffffffffu, inc, d);
			if (lane >= d) inc += up;
			}
		if (lane == 31) s_w[warp] = inc;
		__syncthreads ();
		unsigned long long wex = 0, tot = 0;
		for (int w = 0; w < 32; w++) { if (w < warp) wex += s_w[w];  tot += s_w[w]; }
		const unsigned long long carry = s_carry;
		if (i < nchunks) partial[i] = carry + wex + inc - v;
		__syncthreads ();
		if (threadIdx.x == 0) s_carry = carry + tot;
		__syncthreads ();
		}
	}

__global__ void __launch_bounds__(256)
k_tab_apply (const unsigned int* __restrict__ tab, uint64_t n, const unsigned long long* __restrict__ partial,
             unsigned long long* __restrict__ off)
	{
	__shared__ unsigned long long s_w[8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t c0 = (uint64_t) blockIdx.x * TAB_CHUNK;
	// thread owns 32 consecutive entries
	const uint64_t i0 = c0 + (uint64_t) threadIdx.x * 32;
	unsigned int v[32];
	unsigned long long tot = 0;
	#pragma unroll
	for (int q = 0; q < 8; q++)
		{
		uint4 x = (i0 + q * 4 < n) ? *reinterpret_cast<const uint4*> (tab + i0 + q * 4) : make_uint4 (0, 0, 0, 0);
		v[q*4+0] = x.x;  v[q*4+1] = x.y;  v[q*4+2] = x.z;  v[q*4+3] = x.w;
		tot += (unsigned long long) x.x + x.y + x.z + x.w;
		}
	unsigned long long inc = tot;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		unsigned long long up = __shfl_up_sync (0xffffffffu, inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads ();
	unsigned long long run = partial[blockIdx.x] + inc - tot;
	for (int w = 0; w < warp; w++) run += s_w[w];
	if (i0 < n)
		{
		#pragma unroll
		for (int q = 0; q < 32; q += 2)
			{
			ulonglong2 o;
			o.x = run;  run += v[q];
			o.y = run;  run += v[q+1];
			*reinterpret_cast<ulonglong2*> (off + i0 + q) = o;
			}
		}
	}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
                uint64_t ntilesPad, const double* __restrict__ in, double* __restrict__ out, OutMap om, int shift,
                const unsigned long long* __restrict__ tileOff)
	{
	__shared__ unsigned int       s_cnt[SORT_WARPS][256];
	__shared__ unsigned long long s_off[256];
	__shared__ unsigned long long s_keys[SORT_TILE];
	__shared__ unsigned int       s_dstart[256], s_wtot[SORT_WARPS];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t tFirst = (uint64_t) blockIdx.x * SORT_SB;

	// thread d: positions of digit d for the SORT_SB tiles of this block (one 32-byte sector)
	unsigned long long myOff[SORT_SB];
	{
	const ulonglong2* src = reinterpret_cast<const ulonglong2*> (tileOff + (uint64_t) threadIdx.x * ntilesPad + tFirst);
	ulonglong2 a = src[0], b = src[1];
	myOff[0] = a.x;  myOff[1] = a.y;  myOff[2] = b.x;  myOff[3] = b.y;
	}

	#pragma unroll 1
	for (int k = 0; k < SORT_SB; k++)
		{
		const uint64_t t = tFirst + k;
		if (t >= ntiles) break;
		__syncthreads ();
		for (int i = threadIdx.x; i < SORT_WARPS * 256; i += SORT_THREADS) (&s_cnt[0][0])[i] = 0;
		s_off[threadIdx.x] = (k == 0) ? myOff[0] : (k == 1) ? myOff[1] : (k == 2) ? myOff[2] : myOff[3];
		__syncthreads ();
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * SORT_TILE;
		const uint32_t n  = (uint32_t) ((sd.hi - t0 < SORT_TILE) ? (sd.hi - t0) : SORT_TILE);

		unsigned long long key[SORT_ROUNDS];
		unsigned short     rank[SORT_ROUNDS];
		#pragma unroll
		for (int r = 0; r < SORT_ROUNDS; r++)
			{
			const uint32_t idx = warp * (SORT_ROUNDS * 32) + r * 32 + lane;
			key[r] = (idx < n) ? f64_key (in[t0 + idx]) : ~0ull;
			}
		#pragma unroll
		for (int r = 0; r < SORT_ROUNDS; r++)
			{
			const uint32_t idx = warp * (SORT_ROUNDS * 32) + r * 32 + lane;
			rank[r] = 0;
			sort_rank_round (key[r], shift, idx < n, lane, s_cnt[warp], rank[r]);
			}
		__syncthreads ();
		// per-warp exclusive offsets of every digit; thread d also learns the tile's total for digit d
		unsigned int digitTot = 0;
		{
		const int d = threadIdx.x;
		#pragma unroll
		for (int w = 0; w < SORT_WARPS; w++) { unsigned int c = s_cnt[w][d];  s_cnt[w][d] = digitTot;  digitTot += c; }
		}
		// exclusive prefix of the 256 digit totals: where digit d starts inside the tile once it is sorted
		{
		unsigned int inc = digitTot;
		#pragma unroll
		for (int dd = 1; dd < 32; dd <<= 1)
			{
			unsigned int up = __shfl_up_sync (0xffffffffu, inc, dd);
			if (lane >= dd) inc += up;
			}
		if (lane == 31) s_wtot[warp] = inc;
		__syncthreads ();
		unsigned int wex = 0;
		#pragma unroll
		for (int w = 0; w < SORT_WARPS; w++) if (w < warp) wex += s_wtot[w];
		s_dstart[threadIdx.x] = wex + inc - digitTot;
		}
		__syncthreads ();
		// keys to their place in the tile-sorted order (shared memory) ...
		#pragma unroll
		for (int r = 0; r < SORT_ROUNDS; r++)
			{
			const uint32_t idx = warp * (SORT_ROUNDS * 32) + r * 32 + lane;
			if (idx < n)
				{
				const unsigned d = (unsigned) ((key[r] >> shift) & 255ull);
				s_keys[s_dstart[d] + s_cnt[warp][d] + rank[r]] = key[r];
				}
			}
		__syncthreads ();
		// ... and from there to global memory: consecutive threads write consecutive cells of one digit's
		// run, so the stores fill whole sectors (keys written straight from the ranking loop touched one
		// 32-byte sector per 8-byte key: 4x the crossbar traffic)
		for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS)
			{
			const unsigned long long k = s_keys[i];
			const unsigned d = (unsigned) ((k >> shift) & 255ull);
			const unsigned long long pos = s_off[d] + (i - s_dstart[d]);
			out[out_cell (om, pos)] = key_f64 (k);
			}
		}
	}

// ---------------------------------------------------------------------------
// host driver of the radix sort
// ---------------------------------------------------------------------------

struct SortPlan
	{
	// input view: either the caller's layout or a linear pseudo-layout over `n` cells
	const SegDev*   segs;
	const uint64_t* base;
	int             nseg;
	uint64_t        ntiles;
	};

// device scratch for the sorter, carved from workspace slot 4
struct SortScratch
	{
	unsigned long long* orand;      // 2
	SegDev*             linSeg;     // 1 pseudo segment
	uint64_t*           linBase;    // 2
	uint64_t*           prefix;     // nseg+1 (segmented destination)
	unsigned int*       tileHist;   // 256 * ntilesPad
	unsigned long long* tileOff;    // 256 * ntilesPad
	unsigned long long* partial;    // chunks
	uint64_t            ntilesPadMax;
	};

static inline uint64_t pad8 (uint64_t x) { return (x + 7) / 8 * 8; }

static int sort_scratch (gdsp_ctx* c, uint64_t ntilesMax, int nseg, SortScratch* s)
	{
	const uint64_t tp = pad8 (ntilesMax);
	const uint64_t entries = 256 * tp;
	const uint64_t chunks = (entries + TAB_CHUNK - 1) / TAB_CHUNK + 8;
	size_t small = 4096 + ((sizeof (uint64_t) * (nseg + 8) + 255) / 256) * 256;
	size_t bytes = small + entries * 4 + 256 + entries * 8 + 256 + chunks * 8;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 4, bytes, &ws));
	char* p = (char*) ws;
	s->orand = (unsigned long long*) p;        p += 256;
	s->linSeg = (SegDev*) p;                   p += 256;
	s->linBase = (uint64_t*) p;                p += 256;
	s->prefix = (uint64_t*) p;
	p = (char*) ws + small;
	s->tileHist = (unsigned int*) p;           p += entries * 4 + 256;
	s->tileOff = (unsigned long long*) p;      p += entries * 8 + 256;
	s->partial = (unsigned long long*) p;
	s->ntilesPadMax = tp;
	return GDSP_OK;
	}

// Sort the owned cells of `src` (viewed through plan `inPlan`) ascending.
// Buffers a and b ping-pong; src may be a.  The result lands in *resultBuf
// (a or b), laid out through `finalMap` (segmented) or linearly.  Every pass but the last writes
// LINEARLY from the start of its buffer; with `lastDst` the last of two or more passes writes there
// instead (so a and b can both be scratch and only the final, segmented pass touches the signal).
static int radix_sort (gdsp_ctx* c, const SortPlan& inPlan, const SortPlan& linPlan, const double* src,
                       double* a, double* b, const OutMap& finalMap, const SortScratch& sc,
                       double** resultBuf, int* passesRun, double* lastDst = NULL)
	{
	unsigned long long init[4] = { 0ull, ~0ull, 0ull, 0ull };
	GDSP_CUDA (cudaMemcpyAsync (sc.orand, init, sizeof (init), cudaMemcpyHostToDevice, c->stream));
	int grid = c->sm_count * 8;
	if ((uint64_t) grid > inPlan.ntiles) grid = (int) inPlan.ntiles;
	k_sort_orand<<<grid, 256, 0, c->stream>>> (inPlan.segs, inPlan.base, inPlan.nseg, inPlan.ntiles, src, sc.orand);
	GDSP_KERNEL_CHECK ();
	unsigned long long oa[2];
	GDSP_CUDA (cudaMemcpyAsync (oa, sc.orand, sizeof (oa), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	int shifts[8], k = 0;
	for (int d = 0; d < 8; d++)
		if ((((oa[0] ^ oa[1]) >> (8 * d)) & 255ull) != 0) shifts[k++] = 8 * d;
	*passesRun = k;
	if (k == 0)
		{
		// all keys identical: the data is already sorted where it lies
		*resultBuf = (double*) src;
		return GDSP_OK;
		}
	const double* cur = src;
	SortPlan plan = inPlan;
	OutMap lin;  lin.nseg = 0;  lin.prefix = NULL;  lin.segs = NULL;
	for (int p = 0; p < k; p++)
		{
		const bool last = (p == k - 1);
		double* dst = (cur == a) ? b : a;
		if (last && p > 0 && lastDst != NULL) dst = lastDst;
		const uint64_t tp = pad8 (plan.ntiles);
		GDSP_REQUIRE (tp <= sc.ntilesPadMax, "radix_sort: scratch too small");
		const uint64_t entries = 256 * tp;
		const unsigned chunks = (unsigned) ((entries + TAB_CHUNK - 1) / TAB_CHUNK);
		k_sort_hist<<<(unsigned) (tp / SORT_HB), SORT_THREADS, 0, c->stream>>> (plan.segs, plan.base, plan.nseg, plan.ntiles, tp,
		        cur, shifts[p], sc.tileHist);
		GDSP_KERNEL_CHECK ();
		k_tab_reduce<<<chunks, 256, 0, c->stream>>> (sc.tileHist, entries, sc.partial);
		GDSP_KERNEL_CHECK ();
		k_tab_scan_partials<<<1, 1024, 0, c->stream>>> (sc.partial, chunks);
		GDSP_KERNEL_CHECK ();
		k_tab_apply<<<chunks, 256, 0, c->stream>>> (sc.tileHist, entries, sc.partial, sc.tileOff);
		GDSP_KERNEL_CHECK ();
		k_sort_scatter<<<(unsigned) ((plan.ntiles + SORT_SB - 1) / SORT_SB), SORT_THREADS, 0, c->stream>>> (plan.segs, plan.base,
		        plan.nseg, plan.ntiles, tp, cur, dst, last ? finalMap : lin, shifts[p], sc.tileOff);
		GDSP_KERNEL_CHECK ();
		cur = dst;
		plan = linPlan;
		}
	*resultBuf = (double*) cur;
	return GDSP_OK;
	}

static int make_lin_plan (gdsp_ctx* c, const SortScratch& sc, uint64_t n, SortPlan* out)
	{
	SegDev ls;  ls.lo = ls.dlo = 0;  ls.hi = ls.dhi = n;  ls.pos0 = 0;  ls.chromLen = (uint32_t) (n > 0xffffffffull ? 0xffffffffu : n);
	uint64_t lb[2] = { 0, (n + SORT_TILE - 1) / SORT_TILE };
	GDSP_CUDA (cudaMemcpyAsync (sc.linSeg, &ls, sizeof (ls), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (sc.linBase, lb, sizeof (lb), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	out->segs = sc.linSeg;  out->base = sc.linBase;  out->nseg = 1;  out->ntiles = lb[1];
	return GDSP_OK;
	}

extern "C" int gdsp_sort_genome (gdsp_ctx* c, const gdsp_layout* L_, double* sig, double* tmp, uint64_t buffer_cells,
                                 int* h_result_in_tmp)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && tmp && h_result_in_tmp, "gdsp_sort_genome: NULL argument");
	GDSP_REQUIRE (sig != tmp, "gdsp_sort_genome: sig and tmp must be different buffers");
	GDSP_REQUIRE (L->cells <= buffer_cells, "gdsp_sort_genome: buffer smaller than the layout");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SORT_TILE, &tm));
	SortScratch sc;
	GDSP_TRY (sort_scratch (c, tm.ntiles, L->nseg, &sc));
	std::vector<uint64_t> prefix (L->nseg + 1);
	uint64_t acc = 0;
	for (int s = 0; s < L->nseg; s++) { prefix[s] = acc;  acc += L->h[s].hi - L->h[s].lo; }
	prefix[L->nseg] = acc;
	GDSP_CUDA (cudaMemcpyAsync (sc.prefix, prefix.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	SortPlan inPlan;  inPlan.segs = L->d;  inPlan.base = tm.d_base;  inPlan.nseg = L->nseg;  inPlan.ntiles = tm.ntiles;

	// Genomic signals are often mostly zeros (uncovered bases, thresholded or peak-picked tracks).
	// One pass counts the cells that are exactly +0.0; when they are the majority only the other
	// cells are sorted and the zeros are written back as a block.
	unsigned long long init[4] = { 0ull, ~0ull, 0ull, 0ull }, st[4];
	GDSP_CUDA (cudaMemcpyAsync (sc.orand, init, sizeof (init), cudaMemcpyHostToDevice, c->stream));
	int grid = c->sm_count * 8;
	if ((uint64_t) grid > tm.ntiles) grid = (int) tm.ntiles;
	k_sort_orand<<<grid, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, sc.orand);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemcpyAsync (st, sc.orand, sizeof (st), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	const uint64_t nZero = st[2], nNeg = st[3], nOther = L->cells - nZero;
	if (nZero * 2 >= L->cells && 2 * nOther + 64 <= buffer_cells)
		{
		*h_result_in_tmp = 0;
		if (nOther == 0) return GDSP_OK;                      // all zeros: already sorted
		unsigned long long* d_counter = sc.orand + 8;
		GDSP_CUDA (cudaMemsetAsync (d_counter, 0, 8, c->stream));
		k_sort_compact_nonzero<<<grid, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, tmp, d_counter);
		GDSP_KERNEL_CHECK ();
		SortPlan lp;
		GDSP_TRY (make_lin_plan (c, sc, nOther, &lp));
		OutMap lin;  lin.nseg = 0;  lin.prefix = NULL;  lin.segs = NULL;
		double* sorted = NULL;  int passes = 0;
		double* bufB = tmp + ((nOther + 63) / 64) * 64;
		GDSP_TRY (radix_sort (c, lp, lp, tmp, tmp, bufB, lin, sc, &sorted, &passes));
		k_sort_fill_with_zeros<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sc.prefix, sorted,
		        nNeg, nZero, sig);
		GDSP_KERNEL_CHECK ();
		return GDSP_OK;
		}

	SortPlan linPlan;
	GDSP_TRY (make_lin_plan (c, sc, L->cells, &linPlan));
	OutMap fm;  fm.nseg = L->nseg;  fm.prefix = sc.prefix;  fm.segs = L->d;
	double* res = NULL;  int passes = 0;
	// The intermediate passes write linearly from the start of their buffer.  When the layout is the
	// front of the buffer (the whole genome, or the first chromosomes of it) that region of `sig` holds
	// only cells being sorted.  Any other layout (one chromosome in the middle of the genome: the
	// per-chromosome sorts of the percentile passes) must not lose its neighbours: both ping-pong
	// buffers are then halves of `tmp` and only the last, segmented pass writes into `sig`.
	bool isFront = L->h[0].lo < GDSP_ALIGN;
	for (int s = 0; s + 1 < L->nseg && isFront; s++)
		isFront = L->h[s+1].lo >= L->h[s].hi && L->h[s+1].lo - L->h[s].hi < GDSP_ALIGN;
	const uint64_t half = ((L->cells + 63) / 64) * 64;
	if (!isFront && 2 * half <= buffer_cells)
		{
		GDSP_TRY (radix_sort (c, inPlan, linPlan, sig, tmp, tmp + half, fm, sc, &res, &passes, sig));
		*h_result_in_tmp = (res == sig) ? 0 : 1;          // one pass: segmented into tmp; none: untouched; else sig
		return GDSP_OK;
		}
	GDSP_TRY (radix_sort (c, inPlan, linPlan, sig, sig, tmp, fm, sc, &res, &passes));
	*h_result_in_tmp = (res == tmp) ? 1 : 0;
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// percentile followed by binarize
//
// `percentile P = binarize --threshold=percentileP` (BASELINE configs 3 and the 5-stage pipeline) binarizes
// the SORTED genome the reference's percentile leaves behind (percentile.c:611-651): a step function.
// Sorting 3.1 G values only to threshold them is wasted work -- the step sits at
// cells - #(v > T), so one counting pass and one fill give the same bytes.  NaNs sort to the ends but
// never compare above a threshold, so a signal that holds any is left to the general path.
// ---------------------------------------------------------------------------

// res[0] += #(v > T) (v >= T with ties above), res[1] += #NaN
__global__ void __launch_bounds__(256)
k_count_above (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
               const double* __restrict__ in, double T, int tiesAbove, unsigned long long* __restrict__ res)
	{
	unsigned int above = 0, nan = 0;                   // per thread: at most ntiles/grid * 16 cells, far below 2^32
	for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
		{
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * SORT_TILE;
		uint64_t t1 = t0 + SORT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
		if (t1 - t0 == SORT_TILE)
			{
			double v[16];
			#pragma unroll
			for (int r = 0; r < 4; r++)
				ldg_stream4 (in + t0 + r * 1024 + threadIdx.x * 4, v[4 * r], v[4 * r + 1], v[4 * r + 2], v[4 * r + 3]);
			#pragma unroll
			for (int k = 0; k < 16; k++)
				{
				above += tiesAbove ? (v[k] >= T) : (v[k] > T);
				nan   += (v[k] != v[k]);
				}
			}
		else
			for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
				{
				const double v = in[i];
				above += tiesAbove ? (v >= T) : (v > T);
				nan   += (v != v);
				}
		}
	unsigned long long a = above, b = nan;
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		{
		a += __shfl_xor_sync (0xffffffffu, a, d);
		b += __shfl_xor_sync (0xffffffffu, b, d);
		}
	__shared__ unsigned long long s_a[8], s_b[8];
	if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a;  s_b[threadIdx.x >> 5] = b; }
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		for (int w = 1; w < 8; w++) { a += s_a[w];  b += s_b[w]; }
		if (a) atomicAdd (&res[0], a);
		if (b) atomicAdd (&res[1], b);
		}
	}

// cell at sorted position p (layout order) = p >= step ? one : zero
__global__ void __launch_bounds__(256)
k_fill_step (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
             const uint64_t* __restrict__ prefix, unsigned long long step, double one, double zero, double* __restrict__ out)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * SORT_TILE;
	uint64_t t1 = t0 + SORT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	const uint64_t p0 = prefix[seg] + (t0 - sd.lo);
	if (t1 - t0 == SORT_TILE)
		{
		#pragma unroll
		for (int r = 0; r < 4; r++)
			{
			const uint32_t e = r * 1024 + threadIdx.x * 4;
			const uint64_t p = p0 + e;
			stg_stream4 (out + t0 + e, (p >= step) ? one : zero, (p + 1 >= step) ? one : zero,
			             (p + 2 >= step) ? one : zero, (p + 3 >= step) ? one : zero);
			}
		}
	else
		for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256) out[i] = (p0 + (i - t0) >= step) ? one : zero;
	}

extern "C" int gdsp_sorted_binarize (gdsp_ctx* c, const gdsp_layout* L_, double* sig, double threshold, int ties_above,
                                     double one, double zero, int* h_done)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_done, "gdsp_sorted_binarize: NULL argument");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_sorted_binarize");
	*h_done = 0;
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SORT_TILE, &tm));
	SortScratch sc;
	GDSP_TRY (sort_scratch (c, tm.ntiles, L->nseg, &sc));
	std::vector<uint64_t> prefix (L->nseg + 1);
	uint64_t acc = 0;
	for (int s = 0; s < L->nseg; s++) { prefix[s] = acc;  acc += L->h[s].hi - L->h[s].lo; }
	prefix[L->nseg] = acc;
	GDSP_CUDA (cudaMemcpyAsync (sc.prefix, prefix.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	unsigned long long st[2] = { 0ull, 0ull };
	GDSP_CUDA (cudaMemsetAsync (sc.orand, 0, sizeof (st), c->stream));
	int grid = c->sm_count * 8;
	if ((uint64_t) grid > tm.ntiles) grid = (int) tm.ntiles;
	if (grid < 1) return GDSP_OK;
	k_count_above<<<grid, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, sig, threshold, ties_above ? 1 : 0, sc.orand);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemcpyAsync (st, sc.orand, sizeof (st), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (st[1] != 0) return GDSP_OK;                    // NaNs: the caller sorts and binarizes
	k_fill_step<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sc.prefix, L->cells - st[0], one, zero, sig);
	GDSP_KERNEL_CHECK ();
	*h_done = 1;
	return GDSP_OK;
	}

// The same fill for a slab-sharded genome: h_prefix[s] = position of segment s's first owned cell in
// the concatenated chromsSorted genome, `step` = the global position of the first `one` cell.
extern "C" int gdsp_fill_step (gdsp_ctx* c, const gdsp_layout* L_, double* sig, const uint64_t* h_prefix,
                               uint64_t step, double one, double zero)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_prefix, "gdsp_fill_step: NULL argument");
	GDSP_REQUIRE_ALIGNED (sig, "gdsp_fill_step");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, SORT_TILE, &tm));
	if (tm.ntiles == 0) return GDSP_OK;
	SortScratch sc;
	GDSP_TRY (sort_scratch (c, tm.ntiles, L->nseg, &sc));
	GDSP_CUDA (cudaMemcpyAsync (sc.prefix, h_prefix, sizeof (uint64_t) * L->nseg, cudaMemcpyHostToDevice, c->stream));
	k_fill_step<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sc.prefix, step, one, zero, sig);
	GDSP_KERNEL_CHECK ();                               // (the pageable h_prefix was staged before cudaMemcpyAsync returned)
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// percentile selection
// ---------------------------------------------------------------------------

#define PCT_MAXB    256                 // window bounds handled per pass
#define PCT_SAMPLES (1u << 19)      // sample sort + candidate sort are ~equal at this size (profiles/round2_stages.md)

struct SampleSpace
	{
	int             nseg;
	const uint64_t* sprefix;            // nseg+1: sample slots before segment s
	const SegDev*   segs;
	uint32_t        stride;
	};

// cell of the j-th sampled position of the genome (positions 0,stride,2*stride,.. of
// every chromosome, percentile.c:559), or ~0 if the slot is outside this piece
__device__ __forceinline__ uint64_t sample_cell (const SampleSpace& sp, uint64_t j)
	{
	int lo = 0, hi = sp.nseg - 1;
	while (lo < hi)
		{
		int mid = (lo + hi + 1) >> 1;
		if (sp.sprefix[mid] <= j) lo = mid; else hi = mid - 1;
		}
	const SegDev sd = sp.segs[lo];
	const uint64_t first = ((uint64_t) sd.pos0 + sp.stride - 1) / sp.stride * sp.stride;   // first multiple of stride >= pos0
	const uint64_t coord = first + (j - sp.sprefix[lo]) * sp.stride;
	return sd.lo + (coord - sd.pos0);
	}

__global__ void __launch_bounds__(256)
k_pct_sample (SampleSpace sp, uint64_t nslots, const double* __restrict__ sig, double mn, double mx,
              unsigned long long keyLo, unsigned long long keyHi, uint32_t m, unsigned long long seed,
              double* __restrict__ out, unsigned int* __restrict__ count)
	{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	// splitmix64 of (seed + i): slot index
	unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	const uint64_t j = __umul64hi (z, nslots);
	const double v = sig[sample_cell (sp, j)];
	const unsigned long long k = f64_key (v);
	const bool keep = !(v < mn) && !(v > mx) && k >= keyLo && k <= keyHi;
	const unsigned act = __activemask ();
	const unsigned km = __ballot_sync (act, keep);
	if (km == 0) return;
	const int lane = threadIdx.x & 31, leader = __ffs (km) - 1;
	unsigned int b0 = 0;
	if (lane == leader) b0 = atomicAdd (count, (unsigned int) __popc (km));
	b0 = __shfl_sync (act, b0, leader);
	// A first look at the whole key range only places window bounds, and any value will do as a bound: keep the top 32
	// bits of the sample (sign, exponent, 20 mantissa bits), so that the radix sort of the sample finds the four low
	// digits constant and skips them -- 4 passes instead of 8 on real-valued tracks (0.57 ms of a 5 ms selection).
	// Integers below 2^20 (depth) are unchanged, so ties still land ON a bound.  A narrowed bracket keeps full precision
	// (its samples may differ in the low bits only); non-finite samples are kept as they are.
	double w = v;
	const int hi = __double2hiint (v);
	if (keyLo == 0ull && keyHi == ~0ull && (hi & 0x7ff00000) != 0x7ff00000) w = __hiloint2double (hi, 0);
	if (keep) out[b0 + __popc (km & ((1u << lane) - 1u))] = w;
	}

struct PctBounds
	{
	int                nb;                 // number of distinct bounds (ascending keys)
	unsigned long long key[PCT_MAXB];
	unsigned char      compact[PCT_MAXB + 1];   // open region r (before bound r / after the last): compact its cells?
	};

// region of key k: 2*(#bounds < k) + (k equals a bound)
__device__ __forceinline__ int pct_region (const unsigned long long* __restrict__ b, int nb, unsigned long long k, bool& isBound)
	{
	int lo = 0, hi = nb;                       // first bound >= k
	while (lo < hi)
		{
		int mid = (lo + hi) >> 1;
		if (b[mid] < k) lo = mid + 1; else hi = mid;
		}
	isBound = (lo < nb) && (b[lo] == k);
	return 2 * lo + (isBound ? 1 : 0);
	}

#define PCT_TILE 8192

// counts[2*nb+1] region populations; cand: compacted cells of the flagged open regions
template <bool SMALL>
__global__ void __launch_bounds__(256)
k_pct_pass (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
            const double* __restrict__ sig, uint32_t stride, double mn, double mx,
            const __grid_constant__ PctBounds B, unsigned long long* __restrict__ counts,
            double* __restrict__ cand, unsigned long long cap, unsigned long long* __restrict__ ncand)
	{
	// ncand[1] receives the number of qualifying NaN cells (they sort to the ends of the key order)
	unsigned int myNan = 0;
	__shared__ unsigned long long s_key[2 * PCT_MAXB];
	// one set of region counters per warp: on tied data (integer depth) a handful of regions take nearly every cell
	// and the eight warps of a block serialised on the same shared-memory words
	__shared__ unsigned int       s_cntAll[8][2 * PCT_MAXB + 1];
	unsigned int* const s_cnt = s_cntAll[threadIdx.x >> 5];
	const int nb = B.nb, nreg = 2 * nb + 1;
	int half = 1;                                      // smallest power of two above nb, halved: first step of the search
	while (2 * half <= nb) half *= 2;
	for (int i = threadIdx.x; i < 2 * half; i += 256) s_key[i] = (i < nb) ? B.key[i] : ~0ull;
	for (int i = threadIdx.x; i < 8 * (2 * PCT_MAXB + 1); i += 256) (&s_cntAll[0][0])[i] = 0;
	__syncthreads ();
	const int lane = threadIdx.x & 31;

	unsigned int mine[5] = { 0, 0, 0, 0, 0 };         // SMALL: nb <= 2 -> at most 5 regions, counted in registers

	for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
		{
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * PCT_TILE;
		uint64_t t1 = t0 + PCT_TILE;  if (t1 > sd.hi) t1 = sd.hi;
		for (uint64_t i0 = t0; i0 < t1; i0 += 256 * 4)
			{
			// four independent loads per thread before any of the warp votes below (lanes hold NEIGHBOURING cells: a
			// 256-bit load per lane would put them 4 cells apart and shorten the runs the counting shares: measured slower)
			double vv[4];  bool qq[4];
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const uint64_t i = i0 + (uint64_t) u * 256 + threadIdx.x;
				qq[u] = (i < t1);
				if (qq[u] && stride > 1) qq[u] = (((uint64_t) sd.pos0 + (i - sd.lo)) % stride) == 0;
				vv[u] = qq[u] ? sig[i] : 0.0;
				}
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const double v = vv[u];
				bool q = qq[u] && !(v < mn) && !(v > mx);
				myNan += q && (v != v);
				int reg = 0;  bool isB = false;
				if (SMALL) { if (q) reg = pct_region (s_key, nb, f64_key (v), isB); }
				else
					{
					// branch-free search of the padded table (s_key[nb..] = the largest key): lo = number of bounds below k
					const unsigned long long k = f64_key (v);
					int lo = 0;
					for (int step = half; step > 0; step >>= 1)
						if (s_key[lo + step - 1] < k) lo += step;
					isB = q && (lo < nb) && (s_key[lo] == k);
					reg = q ? 2 * lo + (isB ? 1 : 0) : 0;
					}
				if (SMALL)
					{
					if (q)
						{
						#pragma unroll
						for (int r = 0; r < 5; r++) mine[r] += (reg == r);
						}
					}
				else
					{
					// neighbouring lanes hold neighbouring cells: runs of equal regions share one shared-memory atomic
					// (MATCH.ANY is a slow MIO instruction; a run per lane is what it found on noisy signals anyway)
					const unsigned qm = __ballot_sync (0xffffffffu, q);
					const int left = __shfl_up_sync (0xffffffffu, reg, 1);
					const bool head = q && (lane == 0 || ((qm >> (lane - 1)) & 1u) == 0 || left != reg);
					const unsigned heads = __ballot_sync (0xffffffffu, head);
					if (head)
						{
						const unsigned above = (lane == 31) ? 0u : ~((2u << lane) - 1u);
						const unsigned stop = (heads | ~qm) & above;
						const int end = stop ? __ffs (stop) - 1 : 32;
						atomicAdd (&s_cnt[reg], (unsigned int) (end - lane));
						}
					}
				const bool c = q && !isB && B.compact[reg >> 1];
				const unsigned cm = __ballot_sync (0xffffffffu, c);
				if (cm)
					{
					unsigned long long b0 = 0;
					if (lane == __ffs (cm) - 1) b0 = atomicAdd (ncand, (unsigned long long) __popc (cm));
					b0 = __shfl_sync (0xffffffffu, b0, __ffs (cm) - 1);
					if (c)
						{
						const unsigned long long slot = b0 + __popc (cm & ((1u << lane) - 1u));
						if (slot < cap) cand[slot] = v;
						}
					}
				}
			}
		}
	if (SMALL)
		{
		#pragma unroll
		for (int r = 0; r < 5; r++)
			{
			unsigned int x = mine[r];
			#pragma unroll
			for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync (0xffffffffu, x, d);
			if (lane == 0 && x && r < nreg) atomicAdd (&s_cnt[r], x);
			}
		}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) myNan += __shfl_xor_sync (0xffffffffu, myNan, d);
	if (lane == 0 && myNan) atomicAdd (&ncand[1], (unsigned long long) myNan);
	__syncthreads ();
	for (int i = threadIdx.x; i < nreg; i += 256)
		{
		unsigned long long t = 0;
		#pragma unroll
		for (int w = 0; w < 8; w++) t += s_cntAll[w][i];
		if (t) atomicAdd (&counts[i], t);
		}
	}

// ---------------------------------------------------------------------------
// k_pct_pass_small<NB>: the same counting / compaction pass for NB <= 2 bounds (one open
// percentile, the common case), written for issue rate: a lane reads 4 consecutive cells with one
// 256-bit load, the region counts are kept as four CUMULATIVE per-thread counters (key < k0,
// key <= k0, key < k1, key <= k1 -- two integer compares and a predicated add each) and the
// compaction branch is taken only by warps that actually hold a candidate.  Edge tiles (shorter
// than PCT_TILE) and stride > 1 take a scalar path.
// ---------------------------------------------------------------------------

struct PctSmall
	{
	unsigned long long k0, k1;
	double             d0, d1;         // the same bounds as doubles (FAST path)
	unsigned int       cmask;          // bit o: compact open region o
	int                limits;         // 0: every value qualifies (no --min/--max)
	};

template <int NB>
struct PctAcc
	{
	unsigned int tot, lt0, le0, lt1, le1, nan, negz;
	__device__ __forceinline__ void clear () { tot = lt0 = le0 = lt1 = le1 = nan = negz = 0; }
	// returns true when the cell must be compacted
	__device__ __forceinline__ bool add (const PctSmall& P, double v, bool q)
		{
		const unsigned long long k = f64_key (v);
		tot += q;
		nan += q && (v != v);
		bool a0 = false, b0 = false, a1 = false, b1 = false;
		if (NB >= 1) { a0 = q && (k < P.k0);  b0 = q && (k <= P.k0);  lt0 += a0;  le0 += b0; }
		if (NB >= 2) { a1 = q && (k < P.k1);  b1 = q && (k <= P.k1);  lt1 += a1;  le1 += b1; }
		int  o   = 0;
		bool isB = false;
		if (NB >= 1) { o = b0 ? 0 : 1;  isB = b0 && !a0; }
		if (NB >= 2) { o += b1 ? 0 : 1; isB = isB || (b1 && !a1); }
		return q && !isB && ((P.cmask >> o) & 1u);
		}
	// FAST: every cell qualifies, the bounds are finite and non-zero and the signal is expected to be finite,
	// so a < b <=> key(a) < key(b) and the compares can run on the FP64 pipe (DSETP) instead of 64-bit
	// integer compare pairs -- the integer pipe was what bound this kernel (ncu: ALU 73 %, 39 instructions
	// per cell; profiles/r2_ncu_full_scale4_summaries.txt).  Cells that are NOT finite are counted in `nan`
	// (one 32-bit test of the exponent field); if there are any the host repeats the pass on the key path.
	// ZB: one of the bounds is a zero.  -0.0 and +0.0 are equal to DSETP but not by key, so the cells holding
	// -0.0 are counted and the host moves them to the side of the bound their key puts them on.
	template <bool ZB, bool COUNT_NONFINITE = true>
	__device__ __forceinline__ bool add_fast (const PctSmall& P, double v)
		{
		if (COUNT_NONFINITE) nan += ((unsigned int) __double2hiint (v) & 0x7ff00000u) == 0x7ff00000u;
		if (ZB) negz += (__double_as_longlong (v) == (long long) 0x8000000000000000ull);
		bool a0 = false, b0 = false, a1 = false, b1 = false;
		if (NB >= 1) { a0 = v < P.d0;  b0 = v <= P.d0;  lt0 += a0;  le0 += b0; }
		if (NB >= 2) { a1 = v < P.d1;  b1 = v <= P.d1;  lt1 += a1;  le1 += b1; }
		if (NB == 0) return (P.cmask & 1u) != 0;
		int  o   = 0;
		bool isB = false;
		if (NB >= 1) { o = b0 ? 0 : 1;  isB = b0 && !a0; }
		if (NB >= 2) { o += b1 ? 0 : 1; isB = isB || (b1 && !a1); }
		return !isB && ((P.cmask >> o) & 1u);
		}
	// The cells the settled-side variants (MODE 3 / 4) did not settle, two bounds.  Written out instead of calling
	// add_fast: nvcc 12.9 folded add_fast's compares with the caller's `!(v <= d0)` into a compaction test that
	// only cells BELOW the first bound could pass (SASS: the cmask test predicated on !(v >= d0)), so every
	// candidate of a percentile-99 window was counted and none was compacted.
	// add_above_first: v > d0 (or NaN, counted by the caller: the pass is then repeated on keys)
	__device__ __forceinline__ bool add_above_first (const PctSmall& P, double v)
		{
		const bool a1 = v < P.d1, e1 = (v == P.d1);
		lt1 += a1;  le1 += (a1 || e1);
		const unsigned int bit = a1 ? 2u : 4u;                 // open region 1 (inside the window) or 2 (above it)
		return !e1 && (P.cmask & bit) != 0;
		}
	// add_below_last: v < d1 (or NaN)
	__device__ __forceinline__ bool add_below_last (const PctSmall& P, double v)
		{
		const bool a0 = v < P.d0, e0 = (v == P.d0), in1 = v > P.d0;
		lt0 += a0;  le0 += (a0 || e0);
		lt1 += (a0 || e0 || in1);  le1 += (a0 || e0 || in1);   // every ordered cell here lies below the last bound
		const unsigned int bit = a0 ? 1u : 2u;                 // open region 0 (below the window) or 1 (inside it)
		return (a0 || in1) && (P.cmask & bit) != 0;
		}
	};

__device__ __forceinline__ void pct_compact (bool c, double v, double* __restrict__ cand, unsigned long long cap,
                                             unsigned long long* __restrict__ ncand)
	{
	const int lane = threadIdx.x & 31;
	const unsigned cm = __ballot_sync (0xffffffffu, c);
	if (cm)
		{
		unsigned long long b0 = 0;
		if (lane == __ffs (cm) - 1) b0 = atomicAdd (ncand, (unsigned long long) __popc (cm));
		b0 = __shfl_sync (0xffffffffu, b0, __ffs (cm) - 1);
		if (c)
			{
			const unsigned long long slot = b0 + __popc (cm & ((1u << lane) - 1u));
			if (slot < cap) cand[slot] = v;
			}
		}
	}

// Candidates are staged per block in shared memory and leave with ONE global atomic per tile: with a
// warp-aggregated atomic per candidate-holding warp, the 3.7 M candidates of a percentile-99 window on a
// smoothed hg38 track all hit the single counter `ncand` and the pass took 7.0 ms instead of 4.0
// (profiles/round2_stages.md).  A tile that holds more than PCT_STAGE candidates spills the rest directly.
#define PCT_STAGE 1024
__device__ __forceinline__ void pct_stage (bool c, double v, double* __restrict__ s_stage, unsigned int* __restrict__ s_n,
                                           double* __restrict__ cand, unsigned long long cap, unsigned long long* __restrict__ ncand)
	{
	const int lane = threadIdx.x & 31;
	const unsigned cm = __ballot_sync (0xffffffffu, c);
	if (cm)
		{
		unsigned int b0 = 0;
		if (lane == __ffs (cm) - 1) b0 = atomicAdd (s_n, (unsigned int) __popc (cm));
		b0 = __shfl_sync (0xffffffffu, b0, __ffs (cm) - 1);
		if (c)
			{
			const unsigned int slot = b0 + __popc (cm & ((1u << lane) - 1u));
			if (slot < PCT_STAGE) s_stage[slot] = v;
			else
				{
				const unsigned long long g = atomicAdd (ncand, 1ull);      // rare: more than PCT_STAGE candidates in one tile
				if (g < cap) cand[g] = v;
				}
			}
		}
	}

// MODE 0: key compares (any limits, any bounds); 1: FAST; 2: FAST with a zero bound;
// 3 / 4: FAST for two bounds when nearly every cell lies below (3) / above (4) the window -- the percentile 99
// and percentile 1 shape: one DSETP settles those cells, only the few others take the four compares
// (with all four on every cell the FP64 pipe, not HBM, set the pace: 9.6 ms against 4.0 ms for one bound)
template <int NB, int MODE>
__global__ void __launch_bounds__(256)
k_pct_pass_small (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
                  const double* __restrict__ sig, uint32_t stride, double mn, double mx,
                  const PctSmall P, unsigned long long* __restrict__ counts,
                  double* __restrict__ cand, unsigned long long cap, unsigned long long* __restrict__ ncand)
	{
	__shared__ unsigned int s_cnt[7];
	__shared__ double s_stage[PCT_STAGE];
	__shared__ unsigned int s_nstage;
	__shared__ unsigned long long s_base;
	if (threadIdx.x < 7) s_cnt[threadIdx.x] = 0;
	if (threadIdx.x == 0) s_nstage = 0;
	__syncthreads ();
	PctAcc<NB> A;  A.clear ();
	unsigned int sLt = 0, sLe = 0;                       // MODE 3 / 4: cells settled by the first compare
	const bool compacting = (P.cmask != 0);

	for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
		{
		int seg;  uint64_t tis;
		tile_to_seg (base, nseg, t, seg, tis);
		const SegDev sd = segs[seg];
		const uint64_t t0 = sd.lo + tis * PCT_TILE;
		const uint32_t n  = (uint32_t) ((sd.hi - t0 < PCT_TILE) ? (sd.hi - t0) : PCT_TILE);
		const double* p = sig + t0;
		if (n == PCT_TILE && stride == 1)
			{
			// 8 rounds of 256 lanes x 4 cells, two rounds in flight
			#pragma unroll 1
			for (uint32_t j0 = 0; j0 < PCT_TILE; j0 += 2048)
				{
				double v[8];
				ldg_stream4 (p + j0 + 4 * threadIdx.x,        v[0], v[1], v[2], v[3]);
				ldg_stream4 (p + j0 + 1024 + 4 * threadIdx.x, v[4], v[5], v[6], v[7]);
				bool cc[8];  bool any = false;
				if (MODE == 3 || MODE == 4)
					{
					bool rest[8];  bool anyRest = false;
					#pragma unroll
					for (int u = 0; u < 8; u++)
						{
						A.nan += ((unsigned int) __double2hiint (v[u]) & 0x7ff00000u) == 0x7ff00000u;
						// settled: at or beyond the bound on the crowded side (heavy ties sit ON the bound: 99 % zeros
						// under a percentile-99 window that starts at 0.0); two compares tell "beyond" from "on"
						bool settled;
						if (MODE == 3)
							{
							settled = (v[u] <= P.d0);  sLe += settled;  sLt += (v[u] < P.d0);
							A.negz += (__double_as_longlong (v[u]) == (long long) 0x8000000000000000ull);
							}
						else { settled = (v[u] >= P.d1);  sLe += settled;  sLt += (v[u] > P.d1); }
						rest[u] = !settled;  anyRest = anyRest || rest[u];
						cc[u] = false;
						}
					if (__any_sync (0xffffffffu, anyRest))
						{
						#pragma unroll
						for (int u = 0; u < 8; u++)
							if (rest[u]) { cc[u] = (MODE == 3) ? A.add_above_first (P, v[u]) : A.add_below_last (P, v[u]);  any = any || cc[u]; }
						}
					}
				else
				#pragma unroll
				for (int u = 0; u < 8; u++)
					{
					if (MODE == 1) cc[u] = A.template add_fast<false> (P, v[u]);
					else if (MODE == 2) cc[u] = A.template add_fast<true> (P, v[u]);
					else
						{
						const bool q = P.limits ? (!(v[u] < mn) && !(v[u] > mx)) : true;
						cc[u] = A.add (P, v[u], q);
						}
					any = any || cc[u];
					}
				if (MODE != 0) A.tot += 8;
				if (__any_sync (0xffffffffu, any))
					{
					#pragma unroll
					for (int u = 0; u < 8; u++) pct_stage (cc[u], v[u], s_stage, &s_nstage, cand, cap, ncand);
					}
				}
						if (compacting)
				{
				// the tile's candidates leave together: one global atomic, coalesced stores
				__syncthreads ();
				const unsigned int ns = (s_nstage < PCT_STAGE) ? s_nstage : PCT_STAGE;
				if (ns)
					{
					if (threadIdx.x == 0) s_base = atomicAdd (ncand, (unsigned long long) ns);
					__syncthreads ();
					for (unsigned int i = threadIdx.x; i < ns; i += 256)
						{ const unsigned long long g = s_base + i;  if (g < cap) cand[g] = s_stage[i]; }
					__syncthreads ();
					if (threadIdx.x == 0) s_nstage = 0;
					__syncthreads ();
					}
				}
}
		else
			{
			for (uint32_t j0 = 0; j0 < n; j0 += 256)
				{
				const uint32_t j = j0 + threadIdx.x;
				bool q = (j < n);
				if (q && stride > 1) q = (((uint64_t) sd.pos0 + (t0 - sd.lo) + j) % stride) == 0;
				const double v = q ? p[j] : 0.0;
				q = q && !(v < mn) && !(v > mx);
				const bool c = A.add (P, v, q);
				pct_compact (c, v, cand, cap, ncand);
				}
			}
		}

	if (MODE == 3) { A.lt0 += sLt;  A.le0 += sLe;  A.lt1 += sLe;  A.le1 += sLe; }   // at or below the first bound: below the second
	if (MODE == 4) { A.le1 += sLe - sLt; }                                          // on the last bound; beyond it: in no cumulative count
	unsigned int x[7] = { A.tot, A.lt0, A.le0, A.lt1, A.le1, A.nan, A.negz };
	#pragma unroll
	for (int r = 0; r < 7; r++)
		{
		#pragma unroll
		for (int d = 16; d > 0; d >>= 1) x[r] += __shfl_xor_sync (0xffffffffu, x[r], d);
		if ((threadIdx.x & 31) == 0 && x[r]) atomicAdd (&s_cnt[r], x[r]);
		}
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		// cumulative counts -> region populations (region 2i+1 = "equals bound i")
		const unsigned long long tot = s_cnt[0], lt0 = s_cnt[1], le0 = s_cnt[2], lt1 = s_cnt[3], le1 = s_cnt[4];
		if (s_cnt[5]) atomicAdd (&ncand[1], (unsigned long long) s_cnt[5]);      // qualifying NaN (FAST: non-finite) cells
		if (s_cnt[6]) atomicAdd (&ncand[3], (unsigned long long) s_cnt[6]);      // FAST with a zero bound: cells holding -0.0
		if (NB == 0) { if (tot) atomicAdd (&counts[0], tot); }
		if (NB == 1)
			{
			if (lt0)       atomicAdd (&counts[0], lt0);
			if (le0 - lt0) atomicAdd (&counts[1], le0 - lt0);
			if (tot - le0) atomicAdd (&counts[2], tot - le0);
			}
		if (NB == 2)
			{
			if (lt0)       atomicAdd (&counts[0], lt0);
			if (le0 - lt0) atomicAdd (&counts[1], le0 - lt0);
			if (lt1 - le0) atomicAdd (&counts[2], lt1 - le0);
			if (le1 - lt1) atomicAdd (&counts[3], le1 - lt1);
			if (tot - le1) atomicAdd (&counts[4], tot - le1);
			}
		}
	}

// one counting / compaction pass over the layout (d_counts, d_ncand already zeroed)
static inline double host_unkey (unsigned long long k);

static int pct_launch_pass (gdsp_ctx* c, gdsp_layout* L, const TileMap& tmPct, const double* sig, uint32_t stride,
                            double mn, double mx, const PctBounds& B, unsigned long long* d_counts,
                            double* d_cand, unsigned long long cap, unsigned long long* d_ncand,
                            bool allowFast = false, bool* usedFast = NULL, int sideHint = 0)
	{
	// sideHint > 0: nearly every cell lies below the bounds, < 0: above (two-bound FAST pass)
	if (usedFast) *usedFast = false;
	int grid = c->sm_count * 8;
	if ((uint64_t) grid > tmPct.ntiles) grid = (int) tmPct.ntiles;
	if (B.nb <= 2 && (((uintptr_t) sig) & 31u) == 0)
		{
		PctSmall P;
		P.k0 = (B.nb >= 1) ? B.key[0] : 0;  P.k1 = (B.nb >= 2) ? B.key[1] : 0;
		P.cmask = 0;
		for (int r = 0; r <= B.nb; r++) if (B.compact[r]) P.cmask |= 1u << r;
		P.limits = !(mn == -HUGE_VAL && mx == HUGE_VAL);
		// persistent grid: exactly the blocks that are resident at once (a larger grid runs a ragged second wave)
		int perSM = 0;
		GDSP_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&perSM, k_pct_pass_small<2, 0>, 256, 0));
		grid = c->sm_count * (perSM > 0 ? perSM : 1);
		if ((uint64_t) grid > tmPct.ntiles) grid = (int) tmPct.ntiles;
		// FAST: the default limits (-DBL_MAX..DBL_MAX, which only exclude +-inf) and finite non-zero bounds
		P.d0 = (B.nb >= 1) ? host_unkey (B.key[0]) : 0.0;  P.d1 = (B.nb >= 2) ? host_unkey (B.key[1]) : 0.0;
		bool fast = allowFast && stride == 1 && mn <= -DBL_MAX && mx >= DBL_MAX;
		int zeroBounds = 0;
		for (int r = 0; r < B.nb; r++)
			{
			const double d = host_unkey (B.key[r]);
			if (!(d == d) || d > DBL_MAX || d < -DBL_MAX) fast = false;
			if (d == 0.0) zeroBounds++;
			}
		// a zero bound: -0.0 and +0.0 are one value to DSETP and two keys.  The counts are repaired on the host from the
		// number of -0.0 cells; the compaction is only right when the zeros on the far side of the bound's key are
		// not in a compacted region: bound +0.0 -> the region below it, bound -0.0 -> the region above it
		int zeroAt = -1;
		for (int r = 0; r < B.nb; r++) if (host_unkey (B.key[r]) == 0.0) zeroAt = r;
		if (zeroBounds > 1) fast = false;
		if (zeroBounds == 1)
			{
			const bool negZero = std::signbit (host_unkey (B.key[zeroAt]));
			if (!negZero && ((P.cmask >> zeroAt) & 1u)) fast = false;
			if (negZero && ((P.cmask >> (zeroAt + 1)) & 1u)) fast = false;
			}
		if (usedFast) *usedFast = fast;
		if (sideHint > 0 && (P.cmask & 1u)) sideHint = 0;   // the settled side must not be a region that is compacted
		if (sideHint < 0 && (P.cmask & 4u)) sideHint = 0;
		if (sideHint < 0 && zeroBounds != 0) sideHint = 0;
		if (sideHint > 0 && zeroBounds == 1 && zeroAt != 0) sideHint = 0;      // (the settled-side variant repairs a zero FIRST bound only)
		if (fast && B.nb == 2 && sideHint != 0)
			{
			if (sideHint > 0) k_pct_pass_small<2, 3><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			else              k_pct_pass_small<2, 4><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			}
		else if (fast && zeroBounds == 0)
			{
			if (B.nb == 0)      k_pct_pass_small<0, 1><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			else if (B.nb == 1) k_pct_pass_small<1, 1><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			else                k_pct_pass_small<2, 1><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			}
		else if (fast)
			{
			if (B.nb == 1) k_pct_pass_small<1, 2><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			else           k_pct_pass_small<2, 2><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
			}
		else if (B.nb == 0) k_pct_pass_small<0, 0><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
		else if (B.nb == 1) k_pct_pass_small<1, 0><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
		else                k_pct_pass_small<2, 0><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, P, d_counts, d_cand, cap, d_ncand);
		}
	else if (B.nb <= 2)
		k_pct_pass<true><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, B, d_counts, d_cand, cap, d_ncand);
	else
		k_pct_pass<false><<<grid, 256, 0, c->stream>>> (L->d, tmPct.d_base, L->nseg, tmPct.ntiles, sig, stride, mn, mx, B, d_counts, d_cand, cap, d_ncand);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

static inline unsigned long long host_key (double v)
	{
	unsigned long long b;  memcpy (&b, &v, 8);
	return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
	}
static inline double host_unkey (unsigned long long k)
	{
	unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
	double d;  memcpy (&d, &b, 8);  return d;
	}

// [lower, upper) positions of key(v) inside the ascending (key order) array a[0..n): how many candidates
// lie below the selected value and how many equal it
__global__ void k_equal_range (const double* __restrict__ a, unsigned long long n, const double* __restrict__ vals, int nvals,
                               const unsigned long long* __restrict__ lo0, const unsigned long long* __restrict__ hi0,
                               unsigned long long* __restrict__ out)
	{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nvals) return;
	const unsigned long long k = f64_key (vals[i]);
	unsigned long long lo = lo0[i], hi = hi0[i];
	while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1;  if (f64_key (a[mid]) < k) lo = mid + 1; else hi = mid; }
	out[2 * i] = lo;
	hi = hi0[i];
	unsigned long long l2 = lo;
	while (l2 < hi) { const unsigned long long mid = (l2 + hi) >> 1;  if (f64_key (a[mid]) <= k) l2 = mid + 1; else hi = mid; }
	out[2 * i + 1] = l2;
	}

// One counting / compaction pass with its results on the host.  The FAST kernel is tried first; it also
// counts the cells that are not finite, and if there are any the pass is repeated on the key path (which
// treats +-inf as the limits say and counts the NaNs).
static int pct_run_pass (gdsp_ctx* c, gdsp_layout* L, const TileMap& tmPct, const double* sig, uint32_t stride,
                         double mn, double mx, const PctBounds& B, unsigned long long* d_counts,
                         double* d_cand, unsigned long long cap, unsigned long long* d_ncand,
                         std::vector<unsigned long long>& counts, unsigned long long* ncand, unsigned long long* nnan,
                         int sideHint = 0)
	{
	const int nreg = 2 * B.nb + 1;
	counts.assign (nreg, 0);
	for (int attempt = 0; attempt < 2; attempt++)
		{
		bool usedFast = false;
		GDSP_CUDA (cudaMemsetAsync (d_counts, 0, sizeof (unsigned long long) * nreg, c->stream));
		GDSP_CUDA (cudaMemsetAsync (d_ncand, 0, 32, c->stream));
		GDSP_TRY (pct_launch_pass (c, L, tmPct, sig, stride, mn, mx, B, d_counts, d_cand, cap, d_ncand, attempt == 0, &usedFast, sideHint));
		unsigned long long four[4] = { 0, 0, 0, 0 };
		GDSP_CUDA (cudaMemcpyAsync (counts.data (), d_counts, sizeof (unsigned long long) * nreg, cudaMemcpyDeviceToHost, c->stream));
		GDSP_CUDA (cudaMemcpyAsync (four, d_ncand, 32, cudaMemcpyDeviceToHost, c->stream));
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		*ncand = four[0];  *nnan = four[1];
		if (usedFast && four[1] != 0) continue;         // non-finite cells met by the FAST kernel: once more, on keys
		if (usedFast)
			for (int r = 0; r < B.nb; r++)
				{
				// DSETP saw -0.0 == +0.0; by key -0.0 lies below +0.0: move the -0.0 cells where their key puts them
				const double d = host_unkey (B.key[r]);
				if (d != 0.0) continue;
				const unsigned long long negz = four[3];
				if (!std::signbit (d)) { counts[2 * r] += negz;  counts[2 * r + 1] -= negz; }                   // bound +0.0
				else { counts[2 * r + 2] += counts[2 * r + 1] - negz;  counts[2 * r + 1] = negz; }             // bound -0.0
				}
		break;
		}
	return GDSP_OK;
	}

struct PctJob
	{
	unsigned long long nBelow, nEqual;     // samples with a key below / equal to the selected value (exact)
	uint32_t pMilli;
	bool     done;
	double   value;
	// exact bracket learned so far: the wanted element has key in [keyLo,keyHi];
	// `below` samples have keys < keyLo and `inside` samples lie in the bracket
	// (both exact, known once a pass has run: haveCounts)
	unsigned long long keyLo, keyHi;
	unsigned long long below, inside;
	bool     haveCounts;
	};

static int percentiles_impl (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double* tmp,
                             uint64_t buffer_cells, uint32_t stride, double mn, double mx,
                             const uint32_t* h_p_milli, int np, double* h_values, uint64_t* h_num_samples,
                             uint64_t* h_below, uint64_t* h_equal, uint64_t* h_nan);

extern "C" int gdsp_percentiles (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double* tmp,
                                 uint64_t buffer_cells, uint32_t stride, double mn, double mx,
                                 const uint32_t* h_p_milli, int np, double* h_values, uint64_t* h_num_samples)
	{
	return percentiles_impl (c, L_, sig, tmp, buffer_cells, stride, mn, mx, h_p_milli, np, h_values, h_num_samples, NULL, NULL, NULL);
	}

extern "C" int gdsp_percentiles_ranked (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double* tmp,
                                        uint64_t buffer_cells, uint32_t stride, double mn, double mx,
                                        const uint32_t* h_p_milli, int np, double* h_values, uint64_t* h_num_samples,
                                        uint64_t* h_below, uint64_t* h_equal, uint64_t* h_nan)
	{
	GDSP_REQUIRE (h_below && h_equal && h_nan, "gdsp_percentiles_ranked: NULL argument");
	return percentiles_impl (c, L_, sig, tmp, buffer_cells, stride, mn, mx, h_p_milli, np, h_values, h_num_samples, h_below, h_equal, h_nan);
	}

static int percentiles_impl (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double* tmp,
                             uint64_t buffer_cells, uint32_t stride, double mn, double mx,
                             const uint32_t* h_p_milli, int np, double* h_values, uint64_t* h_num_samples,
                             uint64_t* h_below, uint64_t* h_equal, uint64_t* h_nan)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && tmp && h_num_samples, "gdsp_percentiles: NULL argument");
	GDSP_REQUIRE (np == 0 || (h_p_milli && h_values), "gdsp_percentiles: NULL percentile arrays");
	GDSP_REQUIRE (sig != tmp, "gdsp_percentiles: sig and tmp must be different buffers");
	if (stride == 0) stride = 1;
	GDSP_REQUIRE (buffer_cells >= 1024, "gdsp_percentiles: tmp buffer must hold at least 1024 cells");

	// sample-slot prefix (every stride-th chromosome coordinate owned by this piece)
	std::vector<uint64_t> sprefix (L->nseg + 1);
	uint64_t nslots = 0;
	for (int s = 0; s < L->nseg; s++)
		{
		const gdsp_seg& g = L->h[s];
		uint64_t p0 = g.pos0, p1 = p0 + (g.hi - g.lo);
		uint64_t first = (p0 + stride - 1) / stride * stride;
		sprefix[s] = nslots;
		if (first < p1) nslots += (p1 - 1 - first) / stride + 1;
		}
	sprefix[L->nseg] = nslots;

	TileMap tmSort, tmPct;
	GDSP_TRY (gdsp_layout_tilemap (L, SORT_TILE, &tmSort));
	GDSP_TRY (gdsp_layout_tilemap (L, PCT_TILE, &tmPct));
	const uint64_t candCapTotal = buffer_cells / 2 - 64;         // tmp = [candidates | sort ping-pong]
	SortScratch sc;
	GDSP_TRY (sort_scratch (c, (candCapTotal + SORT_TILE - 1) / SORT_TILE + 1, L->nseg, &sc));
	void* wsv;
	GDSP_TRY (gdsp_ws (c, 5, 64 + sizeof (unsigned long long) * (2 * PCT_MAXB + 8) + sizeof (uint64_t) * (L->nseg + 1), &wsv));
	unsigned long long* d_ncand  = (unsigned long long*) wsv;
	unsigned int*       d_scount = (unsigned int*) ((char*) wsv + 16);
	unsigned long long* d_counts = (unsigned long long*) ((char*) wsv + 64);
	uint64_t*           d_sprefix = (uint64_t*) (d_counts + 2 * PCT_MAXB + 8);
	GDSP_CUDA (cudaMemcpyAsync (d_sprefix, sprefix.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));

	double* bufA = tmp;                                  // candidates / samples
	double* bufB = tmp + buffer_cells / 2;               // sort ping-pong partner
	SampleSpace sp;  sp.nseg = L->nseg;  sp.sprefix = d_sprefix;  sp.segs = L->d;  sp.stride = stride;

	std::vector<PctJob> jobs (np);
	for (int i = 0; i < np; i++) { jobs[i].nBelow = jobs[i].nEqual = 0;  jobs[i].pMilli = h_p_milli[i];  jobs[i].done = false;  jobs[i].value = 0;  jobs[i].keyLo = 0;  jobs[i].keyHi = ~0ull;  jobs[i].below = jobs[i].inside = 0;  jobs[i].haveCounts = false; }

	uint64_t numSamples = 0, numNan = 0;
	bool haveCount = false;
	if (h_nan) *h_nan = 0;
	if (nslots == 0) { *h_num_samples = 0;  return GDSP_OK; }

	double* hs = NULL;                                   // sorted sample (page-locked host copy)
	size_t  nhs = 0;
	{ void* hp;  GDSP_TRY (gdsp_host_scratch (c, sizeof (double) * PCT_SAMPLES, &hp));  hs = (double*) hp; }
	for (int iter = 0; iter < 80; iter++)
		{
		// jobs still open in this iteration (at most PCT_MAXB/2 at a time)
		std::vector<int> open;
		for (int i = 0; i < np && (int) open.size () < PCT_MAXB / 2; i++) if (!jobs[i].done) open.push_back (i);
		if (open.empty () && haveCount) break;

		// ---- (1) sample inside the union bracket of the open jobs, sort it
		unsigned long long bLo = ~0ull, bHi = 0ull;
		for (int i : open) { if (jobs[i].keyLo < bLo) bLo = jobs[i].keyLo;  if (jobs[i].keyHi > bHi) bHi = jobs[i].keyHi; }
		if (open.empty ()) { bLo = 0;  bHi = ~0ull; }
		GDSP_CUDA (cudaMemsetAsync (d_scount, 0, 4, c->stream));
		const uint32_t m = (uint32_t) std::min<uint64_t> (PCT_SAMPLES, std::max<uint64_t> (64, buffer_cells / 8));
		k_pct_sample<<<(m + 255) / 256, 256, 0, c->stream>>> (sp, nslots, sig, mn, mx, bLo, bHi, m,
		        0x243F6A8885A308D3ull + 0x9E37ull * iter, bufA, d_scount);
		GDSP_KERNEL_CHECK ();
		unsigned int scount = 0;
		GDSP_CUDA (cudaMemcpyAsync (&scount, d_scount, 4, cudaMemcpyDeviceToHost, c->stream));
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		nhs = scount;
		if (scount > 0)
			{
			SortPlan lp;
			GDSP_TRY (make_lin_plan (c, sc, scount, &lp));
			OutMap lin;  lin.nseg = 0;  lin.prefix = NULL;  lin.segs = NULL;
			double* res = NULL;  int passes = 0;
			GDSP_TRY (radix_sort (c, lp, lp, bufA, bufA, bufB, lin, sc, &res, &passes));
			GDSP_CUDA (cudaMemcpyAsync (hs, res, sizeof (double) * scount, cudaMemcpyDeviceToHost, c->stream));
			GDSP_CUDA (cudaStreamSynchronize (c->stream));
			}

		// ---- (2) window bounds for every open job: the sample quantiles around the fraction of the
		// job's bracket that lies below the wanted rank (+-6 sigma of the binomial sampling error)
		PctBounds B;  memset (&B, 0, sizeof (B));
		std::vector<unsigned long long> bounds;
		std::vector<std::pair<unsigned long long, unsigned long long> > win (np);
		double fMin = 2.0, fMax = -1.0;                  // where the wanted ranks sit inside the sampled population
		for (int i : open)
			{
			unsigned long long lo = jobs[i].keyLo, hi = jobs[i].keyHi;
			// the part of the sorted sample that lies inside this job's bracket
			size_t a = std::lower_bound (hs, hs + nhs, lo,
			               [] (double x, unsigned long long k) { return host_key (x) < k; }) - hs;
			size_t b = std::upper_bound (hs, hs + nhs, hi,
			               [] (unsigned long long k, double x) { return k < host_key (x); }) - hs;
			const size_t ns = b - a;
			double f = -1;
			if (!jobs[i].haveCounts) f = (jobs[i].pMilli >= 100000) ? 1.0 : jobs[i].pMilli / 100000.0;
			else if (jobs[i].inside > 0)
				{
				const uint32_t nv = (uint32_t) numSamples;
				unsigned long long rank = (jobs[i].pMilli >= 100000) ? numSamples - 1
				                        : (unsigned long long) (uint32_t) (((uint64_t) nv) * jobs[i].pMilli / (100.0 * 1000));
				f = ((double) (rank - jobs[i].below) + 0.5) / (double) jobs[i].inside;
				}
			if (f >= 0 && !jobs[i].haveCounts) { if (f < fMin) fMin = f;  if (f > fMax) fMax = f; }
			else { fMin = -1.0;  fMax = 2.0; }            // a narrowed bracket: the window sits anywhere inside it
			if (ns >= 64 && f >= 0)
				{
				if (f > 1) f = 1;
				double sdv = sqrt (f * (1 - f) / ns);
				double dl = 6 * sdv + 2.0 / ns;
				long long il = (long long) floor ((f - dl) * ns) - 1, ih = (long long) ceil ((f + dl) * ns) + 1;
				if (il >= 0 && il < (long long) ns) { unsigned long long k = host_key (hs[a + il]);  if (k > lo) lo = k; }
				if (ih >= 0 && ih < (long long) ns) { unsigned long long k = host_key (hs[a + ih]);  if (k < hi) hi = k; }
				}
			win[i] = std::make_pair (lo, hi);
			bounds.push_back (lo);  bounds.push_back (hi);
			}
		std::sort (bounds.begin (), bounds.end ());
		bounds.erase (std::unique (bounds.begin (), bounds.end ()), bounds.end ());
		B.nb = (int) bounds.size ();
		for (int k = 0; k < B.nb; k++) B.key[k] = bounds[k];
		// open region r lies between bound r-1 and bound r; compact it if some job's window spans it
		for (int i : open)
			for (int r = 1; r < B.nb; r++)
				if (bounds[r - 1] >= win[i].first && bounds[r] <= win[i].second) B.compact[r] = 1;

		// ---- (3) the counting / compaction pass
		const int nreg = 2 * B.nb + 1;
		std::vector<unsigned long long> counts;
		unsigned long long ncand = 0;
		{
		unsigned long long nn = 0;
		const int sideHint = (fMin >= 0.75 && fMax <= 1.0) ? 1 : ((fMax <= 0.25 && fMin >= 0.0) ? -1 : 0);
		GDSP_TRY (pct_run_pass (c, L, tmPct, sig, stride, mn, mx, B, d_counts, bufA, candCapTotal, d_ncand, counts, &ncand, &nn, sideHint));
		numNan = nn;
		}
		unsigned long long total = 0;
		for (int r = 0; r < nreg; r++) total += counts[r];
		numSamples = total;  haveCount = true;
		if (total == 0 || np == 0) break;

		// sort the candidates (if they fit)
		const bool candOk = (ncand <= candCapTotal);
		double* sortedCand = NULL;
		if (candOk && ncand > 0)
			{
			SortPlan lp;
			GDSP_TRY (make_lin_plan (c, sc, ncand, &lp));
			OutMap lin;  lin.nseg = 0;  lin.prefix = NULL;  lin.segs = NULL;
			int passes = 0;
			GDSP_TRY (radix_sort (c, lp, lp, bufA, bufA, bufB, lin, sc, &sortedCand, &passes));
			}
		// candidate offset of each compacted open region
		std::vector<unsigned long long> candBefore (B.nb + 2, 0);
		{
		unsigned long long acc = 0;
		for (int r = 0; r <= B.nb; r++) { candBefore[r] = acc;  if (B.compact[r]) acc += counts[2 * r]; }
		}

		// ---- (4) locate every open job's rank
		std::vector<int> fromCand;                       // jobs answered from the sorted candidates in this iteration
		std::vector<unsigned long long> fcLo, fcHi, fcBase;
		for (int i : open)
			{
			// rank exactly as percentile.c:588/:686 computes it (numValues is a u32 there)
			const uint32_t nv = (uint32_t) total;
			unsigned long long rank;
			if (jobs[i].pMilli >= 100000) rank = total - 1;
			else
				{
				rank = (uint32_t) (((uint64_t) nv) * jobs[i].pMilli / (100.0 * 1000));
				if (rank >= total) rank = total - 1;
				}
			unsigned long long cum = 0;
			int reg = 0;
			for (reg = 0; reg < nreg; reg++) { if (rank < cum + counts[reg]) break;  cum += counts[reg]; }
			if (reg & 1)
				{
				jobs[i].value = host_unkey (bounds[reg >> 1]);  jobs[i].done = true;
				jobs[i].nBelow = cum;  jobs[i].nEqual = counts[reg];
				continue;
				}
			const int r = reg >> 1;
			if (B.compact[r] && candOk)
				{
				double v;
				GDSP_CUDA (cudaMemcpyAsync (&v, sortedCand + candBefore[r] + (rank - cum), 8, cudaMemcpyDeviceToHost, c->stream));
				GDSP_CUDA (cudaStreamSynchronize (c->stream));
				jobs[i].value = v;  jobs[i].done = true;
				if (h_below != NULL)
					{
					fromCand.push_back (i);
					fcLo.push_back (candBefore[r]);  fcHi.push_back (candBefore[r] + counts[reg]);  fcBase.push_back (cum);
					}
				continue;
				}
			// missed: the wanted key lies strictly inside open region r; if the region is small
			// enough, the next pass compacts all of it
			jobs[i].keyLo = (r > 0) ? bounds[r - 1] + 1 : 0ull;
			jobs[i].keyHi = (r < B.nb) ? bounds[r] - 1 : ~0ull;
			jobs[i].below = cum;  jobs[i].inside = counts[reg];  jobs[i].haveCounts = true;
			}
		if (!fromCand.empty ())
			{
			// how many cells lie below / equal each value found among the candidates: one small kernel
			// (a binary search per job in the sorted candidates), results through the unused half of bufB
			const int nf = (int) fromCand.size ();
			std::vector<double> fv (nf);
			for (int k = 0; k < nf; k++) fv[k] = jobs[fromCand[k]].value;
			void* wq;
			GDSP_TRY (gdsp_ws (c, 6, (size_t) nf * 48, &wq));
			double* d_v = (double*) wq;
			unsigned long long* d_lo = (unsigned long long*) (d_v + nf);
			unsigned long long* d_hi = d_lo + nf;
			unsigned long long* d_out = d_hi + nf;
			GDSP_CUDA (cudaMemcpyAsync (d_v, fv.data (), 8 * nf, cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaMemcpyAsync (d_lo, fcLo.data (), 8 * nf, cudaMemcpyHostToDevice, c->stream));
			GDSP_CUDA (cudaMemcpyAsync (d_hi, fcHi.data (), 8 * nf, cudaMemcpyHostToDevice, c->stream));
			k_equal_range<<<(nf + 63) / 64, 64, 0, c->stream>>> (sortedCand, ncand, d_v, nf, d_lo, d_hi, d_out);
			GDSP_KERNEL_CHECK ();
			std::vector<unsigned long long> er (2 * nf);
			GDSP_CUDA (cudaMemcpyAsync (er.data (), d_out, 16 * nf, cudaMemcpyDeviceToHost, c->stream));
			GDSP_CUDA (cudaStreamSynchronize (c->stream));
			for (int k = 0; k < nf; k++)
				{
				jobs[fromCand[k]].nBelow = fcBase[k] + (er[2 * k] - fcLo[k]);
				jobs[fromCand[k]].nEqual = er[2 * k + 1] - er[2 * k];
				}
			}
		}
	if (h_nan) *h_nan = numNan;
	for (int i = 0; i < np; i++)
		{
		if (h_below) { h_below[i] = jobs[i].nBelow;  h_equal[i] = jobs[i].nEqual; }
		GDSP_REQUIRE (jobs[i].done || numSamples == 0, "gdsp_percentiles: selection did not converge for percentile %u", jobs[i].pMilli);
		h_values[i] = jobs[i].value;
		}
	*h_num_samples = numSamples;
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// Building blocks of the percentile selection, for slab-sharded runs: every rank
// samples, counts and compacts its own slab; the host layer (genodsp_b200/slab.py)
// combines samples, region counts and candidates across ranks with all-gather /
// all-reduce and repeats gdsp_percentiles' decision logic.
// ---------------------------------------------------------------------------

static int pct_sample_space (gdsp_ctx* c, gdsp_layout* L, uint32_t stride, uint64_t** d_sprefix, uint64_t* nslots)
	{
	std::vector<uint64_t> sprefix (L->nseg + 1);
	uint64_t n = 0;
	for (int s = 0; s < L->nseg; s++)
		{
		const gdsp_seg& g = L->h[s];
		uint64_t p0 = g.pos0, p1 = p0 + (g.hi - g.lo);
		uint64_t first = (p0 + stride - 1) / stride * stride;
		sprefix[s] = n;
		if (first < p1) n += (p1 - 1 - first) / stride + 1;
		}
	sprefix[L->nseg] = n;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 5, 64 + sizeof (unsigned long long) * (2 * PCT_MAXB + 8) + sizeof (uint64_t) * (L->nseg + 1), &ws));
	*d_sprefix = (uint64_t*) ((char*) ws + 64 + sizeof (unsigned long long) * (2 * PCT_MAXB + 8));
	GDSP_CUDA (cudaMemcpyAsync (*d_sprefix, sprefix.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	*nslots = n;
	return GDSP_OK;
	}

extern "C" int gdsp_pct_sample (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                                double mn, double mx, uint64_t key_lo, uint64_t key_hi, uint32_t m, uint64_t seed,
                                double* d_out, uint32_t* h_count, uint64_t* h_slots)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && d_out && h_count, "gdsp_pct_sample: NULL argument");
	if (stride == 0) stride = 1;
	uint64_t* d_sprefix;  uint64_t nslots;
	GDSP_TRY (pct_sample_space (c, L, stride, &d_sprefix, &nslots));
	if (h_slots) *h_slots = nslots;
	*h_count = 0;
	if (nslots == 0 || m == 0) return GDSP_OK;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 5, 64, &ws));
	unsigned int* d_scount = (unsigned int*) ((char*) ws + 16);
	GDSP_CUDA (cudaMemsetAsync (d_scount, 0, 4, c->stream));
	SampleSpace sp;  sp.nseg = L->nseg;  sp.sprefix = d_sprefix;  sp.segs = L->d;  sp.stride = stride;
	k_pct_sample<<<(m + 255) / 256, 256, 0, c->stream>>> (sp, nslots, sig, mn, mx, key_lo, key_hi, m, seed, d_out, d_scount);
	GDSP_KERNEL_CHECK ();
	GDSP_CUDA (cudaMemcpyAsync (h_count, d_scount, 4, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// The collect pass of op_percentile_apply (percentile.c:547-580) as a permutation kernel.
// The reference walks the qualifying samples (every stride-th chromosome coordinate with
// min <= v <= max) in chromsSorted order and swaps the j-th of them with position j of the concatenated
// genome.  Closed form (tests/percentile_model.py pins it to the reference): with q_j the position of the
// j-th qualifying sample, n their number and rank(p) the number of qualifying positions before p,
//     out[j]   = in[q_j]                       for j < n
//     out[q_j] = in[chase(j)]                  for q_j >= n, chase(j) = j if position j does not qualify,
//                                              else chase(rank(j))   (depth ~ log_stride N)
//     out[p]   = in[p]                         for every other position p >= n
// Three small kernels build the qualifying bit per cell and its exclusive prefix count (one u32 per
// 32-cell word, so rank() is one load + a popcount), the fourth moves the values.
// ---------------------------------------------------------------------------
#define COL_TILE 8192

__global__ void __launch_bounds__(256)
k_collect_flags (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                 const double* __restrict__ sig, uint32_t stride, double mn, double mx,
                 uint32_t* __restrict__ bits, uint32_t* __restrict__ wpre, uint32_t* __restrict__ tileCount)
	{
	__shared__ uint32_t s_w[8];
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * COL_TILE;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t w0 = t0 + (uint64_t) warp * 1024;               // this warp's 32 words
	uint32_t mine = 0;
	for (int it = 0; it < 32; it++)
		{
		const uint64_t i = w0 + (uint64_t) it * 32 + lane;
		bool q = (i < sd.hi);
		if (q && stride > 1) q = (((uint64_t) sd.pos0 + (i - sd.lo)) % stride) == 0;
		if (q) { const double v = sig[i];  q = !(v < mn) && !(v > mx); }
		const uint32_t w = __ballot_sync (0xffffffffu, q);
		if (lane == it) mine = w;
		}
	uint32_t inc = __popc (mine);
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const uint32_t up = __shfl_up_sync (0xffffffffu, inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads ();
	uint32_t wex = 0, tot = 0;
	#pragma unroll
	for (int w = 0; w < 8; w++) { const uint32_t t = s_w[w];  if (w < warp) wex += t;  tot += t; }
	const uint64_t word = w0 / 32 + lane;
	if (w0 + (uint64_t) lane * 32 < sd.hi)                         // (the words after a short last tile belong to the next segment)
		{
		bits[word] = mine;
		wpre[word] = wex + inc - __popc (mine);                    // exclusive inside the tile; the tile's prefix is added later
		}
	if (threadIdx.x == 0) tileCount[blockIdx.x] = tot;
	}

// one block: exclusive prefix of the tile counts (in place), total -> *n
__global__ void __launch_bounds__(1024)
k_collect_tilescan (uint32_t* __restrict__ tileCount, uint64_t ntiles, unsigned long long* __restrict__ n)
	{
	__shared__ unsigned long long s_w[32];
	const uint64_t chunk = (ntiles + 1023) / 1024;
	const uint64_t lo = (threadIdx.x * chunk < ntiles) ? threadIdx.x * chunk : ntiles;
	const uint64_t hi = (lo + chunk < ntiles) ? lo + chunk : ntiles;
	unsigned long long sum = 0;
	for (uint64_t t = lo; t < hi; t++) sum += tileCount[t];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long inc = sum;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		const unsigned long long up = __shfl_up_sync (0xffffffffu, inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads ();
	unsigned long long wex = 0, tot = 0;
	for (int w = 0; w < 32; w++) { if (w < warp) wex += s_w[w];  tot += s_w[w]; }
	unsigned long long run = wex + inc - sum;
	for (uint64_t t = lo; t < hi; t++) { const uint32_t c = tileCount[t];  tileCount[t] = (uint32_t) run;  run += c; }
	if (threadIdx.x == 0) *n = tot;
	}

__global__ void __launch_bounds__(256)
k_collect_addprefix (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                     const uint32_t* __restrict__ tilePre, uint32_t* __restrict__ wpre)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const uint64_t t0 = segs[seg].lo + tis * COL_TILE;
	if (t0 + (uint64_t) threadIdx.x * 32 < segs[seg].hi) wpre[t0 / 32 + threadIdx.x] += tilePre[blockIdx.x];
	}

struct ColMap
	{
	int             nseg;
	const uint64_t* segpos;      // nseg+1: position of every segment's first cell in the concatenated genome
	const SegDev*   segs;
	};

__device__ __forceinline__ uint64_t col_cell (const ColMap& m, uint64_t pos)
	{
	int lo = 0, hi = m.nseg - 1;
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1;  if (m.segpos[mid] <= pos) lo = mid; else hi = mid - 1; }
	return m.segs[lo].lo + (pos - m.segpos[lo]);
	}
__device__ __forceinline__ bool col_bit (const uint32_t* __restrict__ bits, uint64_t cell) { return (bits[cell >> 5] >> (cell & 31)) & 1u; }
__device__ __forceinline__ uint64_t col_rank (const uint32_t* __restrict__ bits, const uint32_t* __restrict__ wpre, uint64_t cell)
	{ return (uint64_t) wpre[cell >> 5] + __popc (bits[cell >> 5] & ((1u << (cell & 31)) - 1u)); }

__global__ void __launch_bounds__(256)
k_collect_move (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, ColMap cm,
                const double* __restrict__ in, double* __restrict__ out,
                const uint32_t* __restrict__ bits, const uint32_t* __restrict__ wpre, const unsigned long long* __restrict__ nPtr)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * COL_TILE;
	uint64_t t1 = t0 + COL_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	const unsigned long long n = *nPtr;
	const uint64_t pos0 = cm.segpos[seg] + (t0 - sd.lo);
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
		{
		const uint64_t pos = pos0 + (i - t0);
		const bool q = col_bit (bits, i);
		const double v = in[i];
		if (q) out[col_cell (cm, col_rank (bits, wpre, i))] = v;   // front: the rank-th qualifying sample
		if (pos >= n)
			{
			if (!q) out[i] = v;
			else
				{
				uint64_t j = col_rank (bits, wpre, i), cj = col_cell (cm, j);
				while (col_bit (bits, cj)) { j = col_rank (bits, wpre, cj);  cj = col_cell (cm, j); }
				out[i] = in[cj];
				}
			}
		}
	}

extern "C" size_t gdsp_percentile_collect_work_bytes (uint64_t buffer_cells)
	{
	return (size_t) ((buffer_cells / 32 + 512) * 8 + (buffer_cells / COL_TILE + 65536 + 64) * 4 + 4096);
	}

extern "C" int gdsp_percentile_collect (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double* out,
                                        uint64_t buffer_cells, void* work, uint32_t stride, double mn, double mx,
                                        uint64_t* h_n)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && out && work && h_n, "gdsp_percentile_collect: NULL argument");
	GDSP_REQUIRE (sig != out, "gdsp_percentile_collect: sig and out must be different buffers");
	if (stride == 0) stride = 1;
	for (int s = 0; s < L->nseg; s++)
		GDSP_REQUIRE (L->h[s].pos0 == 0 && L->h[s].hi - L->h[s].lo == L->h[s].chrom_len,
		              "gdsp_percentile_collect: whole chromosomes only");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, COL_TILE, &tm));
	GDSP_REQUIRE (L->cells < 0xffffffffull, "gdsp_percentile_collect: more than 2^32-1 cells (the reference's numValues is a u32)");
	const uint64_t nwords = buffer_cells / 32 + 512;
	char* p = (char*) work;
	uint32_t* bits = (uint32_t*) p;       p += nwords * 4;
	uint32_t* wpre = (uint32_t*) p;       p += nwords * 4;
	uint32_t* tcnt = (uint32_t*) p;       p += ((tm.ntiles + 63) / 64) * 64 * 4;
	unsigned long long* d_n = (unsigned long long*) p;   p += 64;
	uint64_t* d_segpos = (uint64_t*) p;
	GDSP_REQUIRE ((size_t) (p - (char*) work) + (L->nseg + 1) * 8 <= gdsp_percentile_collect_work_bytes (buffer_cells),
	              "gdsp_percentile_collect: too many segments for the work buffer");
	std::vector<uint64_t> segpos (L->nseg + 1);
	uint64_t acc = 0;
	for (int s = 0; s < L->nseg; s++) { segpos[s] = acc;  acc += L->h[s].hi - L->h[s].lo; }
	segpos[L->nseg] = acc;
	GDSP_CUDA (cudaMemcpyAsync (d_segpos, segpos.data (), sizeof (uint64_t) * (L->nseg + 1), cudaMemcpyHostToDevice, c->stream));
	k_collect_flags<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, stride, mn, mx, bits, wpre, tcnt);
	GDSP_KERNEL_CHECK ();
	k_collect_tilescan<<<1, 1024, 0, c->stream>>> (tcnt, tm.ntiles, d_n);
	GDSP_KERNEL_CHECK ();
	k_collect_addprefix<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tcnt, wpre);
	GDSP_KERNEL_CHECK ();
	ColMap cm;  cm.nseg = L->nseg;  cm.segpos = d_segpos;  cm.segs = L->d;
	k_collect_move<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, cm, sig, out, bits, wpre, d_n);
	GDSP_KERNEL_CHECK ();
	unsigned long long n = 0;
	GDSP_CUDA (cudaMemcpyAsync (&n, d_n, 8, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	*h_n = n;
	return GDSP_OK;
	}

// positions [lo, hi) of the cells whose key equals key(value) in an array sorted by gdsp_sort_array
extern "C" int gdsp_equal_range (gdsp_ctx* c, const double* d_sorted, uint64_t n, double value, uint64_t* h_lo, uint64_t* h_hi)
	{
	GDSP_REQUIRE (c && d_sorted && h_lo && h_hi, "gdsp_equal_range: NULL argument");
	void* wq;
	GDSP_TRY (gdsp_ws (c, 6, 48, &wq));
	double* d_v = (double*) wq;
	unsigned long long* d_lo = (unsigned long long*) (d_v + 1);
	unsigned long long* d_hi = d_lo + 1;
	unsigned long long* d_out = d_hi + 1;
	const unsigned long long zero = 0, nn = n;
	GDSP_CUDA (cudaMemcpyAsync (d_v, &value, 8, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (d_lo, &zero, 8, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (d_hi, &nn, 8, cudaMemcpyHostToDevice, c->stream));
	k_equal_range<<<1, 64, 0, c->stream>>> (d_sorted, n, d_v, 1, d_lo, d_hi, d_out);
	GDSP_KERNEL_CHECK ();
	unsigned long long er[2];
	GDSP_CUDA (cudaMemcpyAsync (er, d_out, 16, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	*h_lo = er[0];  *h_hi = er[1];
	return GDSP_OK;
	}

extern "C" int gdsp_sort_array (gdsp_ctx* c, double* d_a, double* d_b, uint64_t n, int* h_result_in_b)
	{
	GDSP_REQUIRE (c && d_a && d_b && h_result_in_b, "gdsp_sort_array: NULL argument");
	*h_result_in_b = 0;
	if (n < 2) return GDSP_OK;
	SortScratch sc;
	GDSP_TRY (sort_scratch (c, (n + SORT_TILE - 1) / SORT_TILE + 1, 1, &sc));
	SortPlan lp;
	GDSP_TRY (make_lin_plan (c, sc, n, &lp));
	OutMap lin;  lin.nseg = 0;  lin.prefix = NULL;  lin.segs = NULL;
	double* res = NULL;  int passes = 0;
	GDSP_TRY (radix_sort (c, lp, lp, d_a, d_a, d_b, lin, sc, &res, &passes));
	*h_result_in_b = (res == d_b) ? 1 : 0;
	return GDSP_OK;
	}

static int pct_count_impl (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                           double mn, double mx, const uint64_t* h_bound_keys, int nb, const uint8_t* h_compact,
                           uint64_t* h_counts, double* d_cand, uint64_t cap, uint64_t* h_ncand, uint64_t* h_nan);

extern "C" int gdsp_pct_count (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                               double mn, double mx, const uint64_t* h_bound_keys, int nb, const uint8_t* h_compact,
                               uint64_t* h_counts, double* d_cand, uint64_t cap, uint64_t* h_ncand)
	{
	return pct_count_impl (c, L_, sig, stride, mn, mx, h_bound_keys, nb, h_compact, h_counts, d_cand, cap, h_ncand, NULL);
	}

extern "C" int gdsp_pct_count_nan (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                                   double mn, double mx, const uint64_t* h_bound_keys, int nb, const uint8_t* h_compact,
                                   uint64_t* h_counts, double* d_cand, uint64_t cap, uint64_t* h_ncand, uint64_t* h_nan)
	{
	GDSP_REQUIRE (h_nan, "gdsp_pct_count_nan: NULL argument");
	return pct_count_impl (c, L_, sig, stride, mn, mx, h_bound_keys, nb, h_compact, h_counts, d_cand, cap, h_ncand, h_nan);
	}

static int pct_count_impl (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                           double mn, double mx, const uint64_t* h_bound_keys, int nb, const uint8_t* h_compact,
                           uint64_t* h_counts, double* d_cand, uint64_t cap, uint64_t* h_ncand, uint64_t* h_nan)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_counts && h_ncand, "gdsp_pct_count: NULL argument");
	GDSP_REQUIRE (nb >= 0 && nb <= PCT_MAXB, "gdsp_pct_count: at most %d bounds per pass", PCT_MAXB);
	GDSP_REQUIRE (nb == 0 || (h_bound_keys && h_compact), "gdsp_pct_count: NULL bounds");
	if (stride == 0) stride = 1;
	PctBounds B;  memset (&B, 0, sizeof (B));
	B.nb = nb;
	for (int k = 0; k < nb; k++)
		{
		GDSP_REQUIRE (k == 0 || h_bound_keys[k] > h_bound_keys[k-1], "gdsp_pct_count: bounds must be strictly ascending");
		B.key[k] = h_bound_keys[k];
		}
	for (int k = 0; k <= nb; k++) B.compact[k] = (h_compact != NULL && h_compact[k]) ? 1 : 0;
	TileMap tmPct;
	GDSP_TRY (gdsp_layout_tilemap (L, PCT_TILE, &tmPct));
	void* ws;
	GDSP_TRY (gdsp_ws (c, 5, 64 + sizeof (unsigned long long) * (2 * PCT_MAXB + 8), &ws));
	unsigned long long* d_ncand  = (unsigned long long*) ws;
	unsigned long long* d_counts = (unsigned long long*) ((char*) ws + 64);
	const int nreg = 2 * nb + 1;
	std::vector<unsigned long long> counts;
	unsigned long long ncandNan[2] = { 0, 0 };
	GDSP_TRY (pct_run_pass (c, L, tmPct, sig, stride, mn, mx, B, d_counts, d_cand, cap, d_ncand, counts, &ncandNan[0], &ncandNan[1]));
	for (int r = 0; r < nreg; r++) h_counts[r] = counts[r];
	*h_ncand = ncandNan[0];
	if (h_nan) *h_nan = ncandNan[1];
	return GDSP_OK;
	}
