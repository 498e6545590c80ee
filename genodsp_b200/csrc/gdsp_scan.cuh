// gdsp_scan.cuh -- single-pass chained scan ("decoupled look-back") plumbing
// shared by the accumulate, cumulative-sum, clump and run-length kernels.
//
// Tiles are numbered globally (TileMap); the tiles of one segment are
// consecutive, and a scan never crosses a segment: the first tile of a segment
// publishes its inclusive value immediately, which terminates every look-back
// that reaches it.  Tile ids are handed out by an atomic ticket so that every
// tile a block waits on belongs to a block that has already started.
//
// Each tile owns ONE 16-byte status record {value, flag}, written and read with
// single 128-bit volatile accesses, so a reader sees a consistent pair without
// fences.  The look-back is warp-cooperative and inspects 128 predecessors per
// step (4 per lane, all loads in flight together): with ~1k tiles in flight the
// walk to the nearest tile that already knows its inclusive value is a handful
// of L2 round trips instead of hundreds.
#pragma once
#include "gdsp_common.cuh"

#define SCAN_FLAG_EMPTY 0ull
#define SCAN_FLAG_AGG   1ull
#define SCAN_FLAG_INCL  2ull

struct __align__(16) ScanRec { unsigned long long bits, flag; };

template <typename T>
struct ScanStatus
	{
	uint32_t* ticket;   // one counter
	ScanRec*  rec;      // per tile
	};

template <typename T>
static inline size_t scan_status_bytes (uint64_t ntiles) { return 256 + ntiles * sizeof (ScanRec); }

template <typename T>
static inline ScanStatus<T> scan_status_carve (void* ws, uint64_t ntiles)
	{
	(void) ntiles;
	ScanStatus<T> s;
	s.ticket = (uint32_t*) ws;
	s.rec    = (ScanRec*) ((char*) ws + 256);
	return s;
	}

// bytes that must be zeroed before each launch (ticket + records)
template <typename T>
static inline size_t scan_status_clear_bytes (uint64_t ntiles) { return 256 + ntiles * sizeof (ScanRec); }

#ifdef __CUDACC__

template <typename T> struct ScanBits;
template <> struct ScanBits<double>
	{ static __device__ __forceinline__ unsigned long long to (double v) { return (unsigned long long) __double_as_longlong (v); }
	  static __device__ __forceinline__ double from (unsigned long long b) { return __longlong_as_double ((long long) b); } };
template <> struct ScanBits<int>
	{ static __device__ __forceinline__ unsigned long long to (int v) { return (unsigned long long) (unsigned int) v; }
	  static __device__ __forceinline__ int from (unsigned long long b) { return (int) (unsigned int) b; } };
template <> struct ScanBits<unsigned long long>
	{ static __device__ __forceinline__ unsigned long long to (unsigned long long v) { return v; }
	  static __device__ __forceinline__ unsigned long long from (unsigned long long b) { return b; } };

__device__ __forceinline__ void scan_rec_store (ScanRec* p, unsigned long long bits, unsigned long long flag)
	{
	asm volatile ("st.volatile.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(bits), "l"(flag) : "memory");
	}

__device__ __forceinline__ void scan_rec_load (const ScanRec* p, unsigned long long& bits, unsigned long long& flag)
	{
	asm volatile ("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(bits), "=l"(flag) : "l"(p) : "memory");
	}

__device__ __forceinline__ unsigned long long shfl_down_u64 (unsigned long long v, int d)
	{ return __shfl_down_sync (0xffffffffu, v, d); }

#define SCAN_SUBW 1        // sub-windows of 32 tiles inspected per look-back step

// Warp-cooperative look-back: called by ALL 32 lanes of ONE warp of the block
// (every lane passes the same arguments).  `myAgg` is this tile's aggregate; the
// function publishes it, folds the aggregates of the preceding tiles in tile
// order until a tile with a published inclusive value is met, publishes this
// tile's inclusive value and returns the aggregate of all earlier tiles of the
// segment (identity for the first tile) in every lane.  op(a,b) combines an
// earlier aggregate a with a later aggregate b and may be non-commutative.
template <typename T, typename Op>
__device__ T scan_lookback (const ScanStatus<T>& st, uint64_t tile, bool firstOfSeg,
                            T myAgg, T identity, Op op)
	{
	const int lane = threadIdx.x & 31;
	if (firstOfSeg)
		{
		if (lane == 0) scan_rec_store (&st.rec[tile], ScanBits<T>::to (myAgg), SCAN_FLAG_INCL);
		return identity;
		}
	if (lane == 0) scan_rec_store (&st.rec[tile], ScanBits<T>::to (myAgg), SCAN_FLAG_AGG);

	T excl = identity;
	bool haveExcl = false;
	uint64_t j0 = tile;                       // this step covers tiles j0-1 .. j0-128
	while (true)
		{
		unsigned long long bits[SCAN_SUBW], flag[SCAN_SUBW];
		bool inRange[SCAN_SUBW];
		unsigned inclMask[SCAN_SUBW];
		int  firstK = -1;                     // first sub-window that contains an inclusive value
		bool again;
		do  {
			#pragma unroll
			for (int k = 0; k < SCAN_SUBW; k++)
				{
				const uint64_t back = (uint64_t) k * 32 + lane + 1;
				inRange[k] = (j0 >= back);
				flag[k] = SCAN_FLAG_EMPTY;  bits[k] = 0;
				if (inRange[k]) scan_rec_load (&st.rec[j0 - back], bits[k], flag[k]);
				}
			again = false;  firstK = -1;
			#pragma unroll
			for (int k = 0; k < SCAN_SUBW; k++)
				{
				inclMask[k] = __ballot_sync (0xffffffffu, inRange[k] && flag[k] == SCAN_FLAG_INCL);
				unsigned empty = __ballot_sync (0xffffffffu, inRange[k] && flag[k] == SCAN_FLAG_EMPTY);
				if (firstK < 0)
					{
					// only tiles NEARER than the first inclusive one have to be ready
					if (inclMask[k]) { firstK = k;  empty &= (1u << (__ffs (inclMask[k]) - 1)) - 1u; }
					if (empty) again = true;
					}
				}
			} while (again);

		const int lastK = (firstK < 0) ? SCAN_SUBW - 1 : firstK;
		#pragma unroll
		for (int k = 0; k < SCAN_SUBW; k++)
			{
			if (k > lastK) break;
			const int last = (k == firstK) ? (__ffs (inclMask[k]) - 1) : 31;      // farthest lane that takes part
			T v = identity;
			if (inRange[k] && lane <= last) v = ScanBits<T>::from (bits[k]);
			// ordered fold: lane l+d holds an EARLIER tile than lane l
			#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
				{
				T o = ScanBits<T>::from (shfl_down_u64 (ScanBits<T>::to (v), d));
				if (lane + d <= last && (inRange[k])) v = op (o, v);
				}
			const T w = ScanBits<T>::from (__shfl_sync (0xffffffffu, ScanBits<T>::to (v), 0));
			excl = haveExcl ? op (w, excl) : w;
			haveExcl = true;
			}
		if (firstK >= 0) break;
		j0 -= 32 * SCAN_SUBW;
		}
	if (lane == 0) scan_rec_store (&st.rec[tile], ScanBits<T>::to (op (excl, myAgg)), SCAN_FLAG_INCL);
	return excl;
	}

// block-wide ticket: every thread of the block gets the same tile id
__device__ __forceinline__ uint32_t scan_take_ticket (uint32_t* ticket)
	{
	__shared__ uint32_t s_ticket;
	if (threadIdx.x == 0) s_ticket = atomicAdd (ticket, 1u);
	__syncthreads ();
	return s_ticket;
	}

#endif
