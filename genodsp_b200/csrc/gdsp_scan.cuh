// gdsp_scan.cuh -- single-pass chained scan ("decoupled look-back") plumbing
// shared by the accumulate, cumulative-sum, clump and run-length kernels.
//
// Tiles are numbered globally (TileMap); the tiles of one segment are
// consecutive, and a scan never crosses a segment: the first tile of a segment
// publishes its inclusive value immediately, which terminates every look-back
// that reaches it.  Tile ids are handed out by an atomic ticket so that every
// tile a block waits on belongs to a block that has already started.
#pragma once
#include "gdsp_common.cuh"

#define SCAN_FLAG_EMPTY 0u
#define SCAN_FLAG_AGG   1u
#define SCAN_FLAG_INCL  2u

template <typename T>
struct ScanStatus
	{
	uint32_t* ticket;   // one counter
	uint32_t* flag;     // per tile
	T*        agg;      // per tile: aggregate of this tile alone
	T*        incl;     // per tile: aggregate of all tiles of the segment up to and including this one
	};

// bytes of workspace needed for ntiles tiles (ticket + flags + 2 arrays of T)
template <typename T>
static inline size_t scan_status_bytes (uint64_t ntiles)
	{
	size_t fl = ((ntiles * sizeof (uint32_t) + 255) / 256) * 256;
	size_t ar = ((ntiles * sizeof (T) + 255) / 256) * 256;
	return 256 + fl + 2 * ar;
	}

template <typename T>
static inline ScanStatus<T> scan_status_carve (void* ws, uint64_t ntiles)
	{
	ScanStatus<T> s;
	char* p = (char*) ws;
	size_t fl = ((ntiles * sizeof (uint32_t) + 255) / 256) * 256;
	size_t ar = ((ntiles * sizeof (T) + 255) / 256) * 256;
	s.ticket = (uint32_t*) p;            p += 256;
	s.flag   = (uint32_t*) p;            p += fl;
	s.agg    = (T*) p;                   p += ar;
	s.incl   = (T*) p;
	return s;
	}

// bytes that must be zeroed before each launch (ticket + flags)
template <typename T>
static inline size_t scan_status_clear_bytes (uint64_t ntiles)
	{ return 256 + ((ntiles * sizeof (uint32_t) + 255) / 256) * 256; }

#ifdef __CUDACC__

template <typename T> __device__ __forceinline__ T ld_vol (const T* p) { return *(const volatile T*) p; }
template <typename T> __device__ __forceinline__ void st_vol (T* p, T v) { *(volatile T*) p = v; }

// Warp-cooperative look-back: called by ALL 32 lanes of ONE warp of the block
// (every lane passes the same arguments).  `myAgg` is this tile's aggregate; the
// function publishes it, inspects the 32 preceding tiles at a time (one per lane),
// folds their aggregates in tile order until a tile with a published inclusive
// value is met, publishes this tile's inclusive value and returns the aggregate of
// all earlier tiles of the segment (identity for the first tile) in every lane.
// op(a,b) combines an earlier aggregate a with a later aggregate b and may be
// non-commutative.
template <typename T> struct LbShfl;
template <> struct LbShfl<double>
	{ static __device__ __forceinline__ double down (double v, int d) { return shfl_down_f64 (v, d); }
	  static __device__ __forceinline__ double idx  (double v, int s) { return shfl_idx_f64 (v, s); } };
template <> struct LbShfl<int>
	{ static __device__ __forceinline__ int down (int v, int d) { return __shfl_down_sync (0xffffffffu, v, d); }
	  static __device__ __forceinline__ int idx  (int v, int s) { return __shfl_sync (0xffffffffu, v, s); } };
template <> struct LbShfl<unsigned long long>
	{ static __device__ __forceinline__ unsigned long long down (unsigned long long v, int d) { return __shfl_down_sync (0xffffffffu, v, d); }
	  static __device__ __forceinline__ unsigned long long idx  (unsigned long long v, int s) { return __shfl_sync (0xffffffffu, v, s); } };

template <typename T, typename Op>
__device__ T scan_lookback (const ScanStatus<T>& st, uint64_t tile, bool firstOfSeg,
                            T myAgg, T identity, Op op)
	{
	const int lane = threadIdx.x & 31;
	if (firstOfSeg)
		{
		if (lane == 0)
			{
			st_vol (&st.incl[tile], myAgg);
			__threadfence ();
			st_vol (&st.flag[tile], SCAN_FLAG_INCL);
			}
		return identity;
		}
	if (lane == 0)
		{
		st_vol (&st.agg[tile], myAgg);
		__threadfence ();
		st_vol (&st.flag[tile], SCAN_FLAG_AGG);
		}

	T excl = identity;
	bool haveExcl = false;
	uint64_t j0 = tile;                       // the window covers tiles j0-1 .. j0-32
	while (true)
		{
		// lane l looks at tile j0-1-l; tiles before the segment's first tile are never reached
		// because that first tile always publishes an inclusive value
		const bool inRange = (j0 >= (uint64_t) lane + 1);
		const uint64_t j = inRange ? j0 - 1 - lane : 0;
		uint32_t f = SCAN_FLAG_EMPTY;
		unsigned inclMask, emptyMask;
		do  {
			if (inRange) f = ld_vol (&st.flag[j]);
			inclMask  = __ballot_sync (0xffffffffu, inRange && f == SCAN_FLAG_INCL);
			emptyMask = __ballot_sync (0xffffffffu, inRange && f == SCAN_FLAG_EMPTY);
			// spin only while a tile NEARER than the first inclusive one is still empty
			const unsigned need = inclMask ? ((1u << (__ffs (inclMask) - 1)) - 1u) | (1u << (__ffs (inclMask) - 1)) : 0xffffffffu;
			emptyMask &= need;
			} while (emptyMask != 0);
		__threadfence ();
		const int last = inclMask ? (__ffs (inclMask) - 1) : 31;          // farthest lane that takes part
		T v = identity;
		const bool part = inRange && lane <= last;
		if (part) v = (f == SCAN_FLAG_INCL) ? ld_vol (&st.incl[j]) : ld_vol (&st.agg[j]);
		// ordered fold: lane l+d holds an EARLIER tile than lane l
		#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
			{
			T o = LbShfl<T>::down (v, d);
			if (lane + d <= last) v = op (o, v);
			}
		const T w = LbShfl<T>::idx (v, 0);
		excl = haveExcl ? op (w, excl) : w;
		haveExcl = true;
		if (inclMask) break;
		j0 -= 32;
		}
	if (lane == 0)
		{
		st_vol (&st.incl[tile], op (excl, myAgg));
		__threadfence ();
		st_vol (&st.flag[tile], SCAN_FLAG_INCL);
		}
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
