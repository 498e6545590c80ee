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

// Called by ONE thread of the block.  `myAgg` is this tile's aggregate; returns
// the aggregate of all earlier tiles of the same segment (identity for the
// first tile) and publishes this tile's inclusive value.  op(a,b) combines an
// earlier aggregate a with a later aggregate b.
template <typename T, typename Op>
__device__ T scan_lookback (const ScanStatus<T>& st, uint64_t tile, bool firstOfSeg,
                            T myAgg, T identity, Op op)
	{
	if (firstOfSeg)
		{
		st_vol (&st.incl[tile], myAgg);
		__threadfence ();
		st_vol (&st.flag[tile], SCAN_FLAG_INCL);
		return identity;
		}
	st_vol (&st.agg[tile], myAgg);
	__threadfence ();
	st_vol (&st.flag[tile], SCAN_FLAG_AGG);

	T excl = identity;
	for (uint64_t j = tile - 1; ; j--)
		{
		uint32_t f;
		do { f = ld_vol (&st.flag[j]); } while (f == SCAN_FLAG_EMPTY);
		__threadfence ();
		if (f == SCAN_FLAG_INCL) { excl = op (ld_vol (&st.incl[j]), excl);  break; }
		excl = op (ld_vol (&st.agg[j]), excl);
		}
	st_vol (&st.incl[tile], op (excl, myAgg));
	__threadfence ();
	st_vol (&st.flag[tile], SCAN_FLAG_INCL);
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
