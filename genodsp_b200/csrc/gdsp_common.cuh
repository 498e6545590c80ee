// gdsp_common.cuh -- shared device/host plumbing of the sm_100a operator library.
//
// Compiled with -fmad=false: the reference is an x86-64 baseline build without
// FMA contraction, so every product and sum here rounds
// separately, which is what makes smooth() bit-exact.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <float.h>
#include <map>
#include <vector>

#include "gdsp_b200.h"

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------

void gdsp_set_error (const char* fmt, ...);

#define GDSP_CUDA(call)                                                        \
	do {                                                                       \
		cudaError_t e__ = (call);                                              \
		if (e__ != cudaSuccess)                                                \
			{                                                                  \
			gdsp_set_error ("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, \
			                cudaGetErrorString (e__));                         \
			return GDSP_ERR_CUDA;                                              \
			}                                                                  \
	} while (0)

// every kernel launch of the library goes through this macro; the count is what bench.py reports
extern unsigned long long g_gdsp_launches;
#define GDSP_KERNEL_CHECK()                                                    \
	do { __atomic_fetch_add (&g_gdsp_launches, 1ull, __ATOMIC_RELAXED);  GDSP_CUDA (cudaGetLastError ()); } while (0)

#define GDSP_REQUIRE(cond, ...)                                                \
	do {                                                                       \
		if (!(cond)) { gdsp_set_error (__VA_ARGS__); return GDSP_ERR_ARG; }    \
	} while (0)

// signal buffers are accessed with 256-bit vector loads/stores
#define GDSP_REQUIRE_ALIGNED(p, what)                                          \
	GDSP_REQUIRE ((((uintptr_t) (p)) & 31u) == 0, what ": signal buffers must be 32-byte aligned")

#define GDSP_TRY(call)                                                         \
	do { int s__ = (call); if (s__ != GDSP_OK) return s__; } while (0)

// ---------------------------------------------------------------------------
// segment table on the device
// ---------------------------------------------------------------------------

struct SegDev
	{
	uint64_t lo, hi;      // owned cells
	uint64_t dlo, dhi;    // readable cells (chromosome clip bounds)
	uint32_t pos0;        // chromosome coordinate of cell lo
	uint32_t chromLen;
	};

// per-tile-size prefix of tile counts: tile t of the launch belongs to the
// segment s with base[s] <= t < base[s+1]
struct TileMap
	{
	uint32_t  tile;
	uint64_t  ntiles;
	uint64_t* d_base;     // nseg+1 entries
	};

struct gdsp_layout
	{
	gdsp_ctx*             ctx;
	int                   nseg;
	std::vector<gdsp_seg> h;
	SegDev*               d;
	uint64_t              cells;      // sum (hi-lo)
	uint64_t              span_lo, span_hi;   // min lo, max hi
	uint32_t              max_len;    // longest owned segment
	std::map<uint32_t, TileMap> tiles;
	};

int gdsp_layout_tilemap (gdsp_layout* lay, uint32_t tile, TileMap* out);

#define GDSP_NUM_WS 8

struct gdsp_ctx
	{
	int          device;
	cudaStream_t stream;
	bool         owns_stream;
	int          sm_count;
	int          cc_major, cc_minor;
	size_t       smem_optin;          // max dynamic smem per block
	void*        ws[GDSP_NUM_WS];     // grow-only device workspaces
	size_t       ws_bytes[GDSP_NUM_WS];
	void*        pinned[2];           // staging buffers for *_host entry points
	size_t       pinned_bytes;
	cudaEvent_t  pinned_ev[2];
	cudaEvent_t  t0, t1;
	double*      taps_host;           // last taps uploaded by gdsp_smooth (host copy) ...
	double*      taps_dev;            // ... and their device copy, reused while unchanged
	uint32_t     taps_n;
	void*        host_small;          // page-locked scratch for small device->host results (percentile samples)
	size_t       host_small_bytes;
	int          exact_order;         // gdsp_ctx_set_exact_order: sequential-order slidingsum / cumulativesum / clump
	int          smooth_direct;       // gdsp_ctx_set_smooth_direct: always the direct FIR (k_smooth_ct), for A/B measurements
	};

// gdsp_smooth_sym.cu
int gdsp_smooth_sym_plan (const gdsp_layout* L, const double* in, uint32_t W, const double* h_taps);
int gdsp_smooth_sym_launch (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, uint32_t W, const double* h_taps);

// gdsp_exact.cu
int gdsp_cumulative_sum_exact (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out);
int gdsp_sliding_sum_exact (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, uint32_t W, double denom);
int gdsp_clump_prefix_exact (gdsp_ctx* c, gdsp_layout* L, const double* sig, double T, int above,
                             double* P, double* M, int* segAllNeg);

// grow-only device scratch (slot 0..GDSP_NUM_WS-1); contents undefined
int gdsp_ws (gdsp_ctx* ctx, int slot, size_t bytes, void** out);
// grow-only page-locked host scratch; contents undefined
int gdsp_host_scratch (gdsp_ctx* ctx, size_t bytes, void** out);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------

#ifdef __CUDACC__

// Which segment does tile `t` belong to, and which tile of that segment is it?
// Warp-cooperative: all 32 lanes of the calling warp must call it with the same t.
// Each lane inspects one entry of the tile-count prefix (one coalesced load for up
// to 32 segments) and a vote finds the last entry <= t -- one memory round trip
// instead of a chain of dependent bisection loads at the head of every block.
__device__ __forceinline__ void tile_to_seg (const uint64_t* __restrict__ base, int nseg,
                                             uint64_t t, int& seg, uint64_t& tileInSeg)
	{
	const int lane = threadIdx.x & 31;
	int s = 0;
	uint64_t b = 0;
	for (int s0 = 0; s0 < nseg; s0 += 32)
		{
		const int idx = s0 + lane;
		const uint64_t v = (idx < nseg) ? __ldg (base + idx) : ~0ull;
		const unsigned m = __ballot_sync (0xffffffffu, v <= t);
		if (m == 0) break;
		const int cnt = __popc (m);
		s = s0 + cnt - 1;
		b = __shfl_sync (0xffffffffu, v, cnt - 1);
		if (cnt < 32) break;
		}
	seg = s;
	tileInSeg = t - b;
	}

__device__ __forceinline__ double shfl_up_f64 (double v, int delta)
	{
	int lo = __double2loint (v), hi = __double2hiint (v);
	lo = __shfl_up_sync (0xffffffffu, lo, delta);
	hi = __shfl_up_sync (0xffffffffu, hi, delta);
	return __hiloint2double (hi, lo);
	}

__device__ __forceinline__ double shfl_down_f64 (double v, int delta)
	{
	int lo = __double2loint (v), hi = __double2hiint (v);
	lo = __shfl_down_sync (0xffffffffu, lo, delta);
	hi = __shfl_down_sync (0xffffffffu, hi, delta);
	return __hiloint2double (hi, lo);
	}

__device__ __forceinline__ double shfl_idx_f64 (double v, int src)
	{
	int lo = __double2loint (v), hi = __double2hiint (v);
	lo = __shfl_sync (0xffffffffu, lo, src);
	hi = __shfl_sync (0xffffffffu, hi, src);
	return __hiloint2double (hi, lo);
	}

__device__ __forceinline__ double shfl_xor_f64 (double v, int mask)
	{
	int lo = __double2loint (v), hi = __double2hiint (v);
	lo = __shfl_xor_sync (0xffffffffu, lo, mask);
	hi = __shfl_xor_sync (0xffffffffu, hi, mask);
	return __hiloint2double (hi, lo);
	}

// streaming 128-bit global accesses (data touched once per kernel)
__device__ __forceinline__ double2 ldg_stream (const double* p)
	{
	double2 r;
	asm volatile ("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
	              : "=d"(r.x), "=d"(r.y) : "l"(p));
	return r;
	}

__device__ __forceinline__ void stg_stream (double* p, double2 v)
	{
	asm volatile ("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};"
	              :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
	}

// 256-bit accesses (sm_100): one lane moves a whole 32-byte sector.  A lane that owns 4 consecutive
// cells must NOT use two 128-bit accesses instead: each of them touches half of every sector, and
// the L1 forwards every half-filled sector to the crossbar separately (ncu: l1tex2xbar write bytes
// = 2x the stored bytes, the crossbar port -- not HBM -- then bounds the kernel).
// p must be 32-byte aligned.
__device__ __forceinline__ void stg_stream4 (double* p, double a, double b, double c, double d)
	{
	asm volatile ("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
	              :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
	}

__device__ __forceinline__ void ldg_stream4 (const double* p, double& a, double& b, double& c, double& d)
	{
	asm volatile ("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
	              : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
	}

// Stage cells [g0, g0+count) of `in` into shared memory (index j -> smem[j + (j >> PADSHIFT)], or
// smem[j] when PADSHIFT == 0); cells outside the readable range [dlo,dhi) get `neutral`.
// Interior tiles (the whole range readable -- all but the first/last tile of a chromosome) take a
// branch-free path with 128-bit loads; edge tiles check every cell.
template <int PADSHIFT>
__device__ __forceinline__ uint32_t stage_idx (uint32_t j) { return PADSHIFT ? j + (j >> PADSHIFT) : j; }

// DEEP: four 128-bit loads in flight per thread before the first store.  Measured on hg38 (profiles/round2_stages.md):
// localmax 10.83 -> 10.12 ms, bestmax 14.96 -> 13.77 ms, but slidingsum 9.84 -> 10.16 ms -- so it is a per-kernel choice.
template <int PADSHIFT, bool DEEP = false>
__device__ __forceinline__ void stage_tile (double* smem, const double* __restrict__ in, int64_t g0, uint32_t count,
                                            uint64_t dlo, uint64_t dhi, double neutral)
	{
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	if (g0 >= (int64_t) dlo && g0 + (int64_t) count <= (int64_t) dhi)
		{
		const double* p = in + g0;
		const uint32_t j0 = (uint32_t) (g0 & 1);                 // first cell on a 16-byte boundary
		const uint32_t npair = (count - j0) >> 1;
		if (j0 && tid == 0) smem[stage_idx<PADSHIFT> (0)] = __ldg (p);
		if (DEEP)
		for (uint32_t q = tid; q < npair; q += 4 * nt)
			{
			double2 v[4];
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const uint32_t qq = q + u * nt;
				if (qq < npair) v[u] = __ldg (reinterpret_cast<const double2*> (p + j0 + 2 * qq));
				}
			#pragma unroll
			for (int u = 0; u < 4; u++)
				{
				const uint32_t qq = q + u * nt;
				if (qq < npair)
					{
					const uint32_t j = j0 + 2 * qq;
					smem[stage_idx<PADSHIFT> (j)]     = v[u].x;
					smem[stage_idx<PADSHIFT> (j + 1)] = v[u].y;
					}
				}
			}
		else
		for (uint32_t q = tid; q < npair; q += nt)
			{
			const uint32_t j = j0 + 2 * q;
			const double2 v = __ldg (reinterpret_cast<const double2*> (p + j));
			smem[stage_idx<PADSHIFT> (j)]     = v.x;
			smem[stage_idx<PADSHIFT> (j + 1)] = v.y;
			}
		if (((count - j0) & 1) && tid == nt - 1) smem[stage_idx<PADSHIFT> (count - 1)] = __ldg (p + count - 1);
		}
	else
		{
		for (uint32_t j = tid; j < count; j += nt)
			{
			const int64_t g = g0 + (int64_t) j;
			double v = neutral;
			if (g >= (int64_t) dlo && g < (int64_t) dhi) v = __ldg (in + g);
			smem[stage_idx<PADSHIFT> (j)] = v;
			}
		}
	}

// Asynchronous variant: the cells go to shared memory with 8-byte cp.async (no registers in between), so that a
// persistent block can have its NEXT tile in flight while it works on the current one.  The caller commits the
// group, and waits + __syncthreads() before reading.  Tiles that leave the readable range are staged with plain
// loads and stores (they are the first / last tile of a chromosome).
__device__ __forceinline__ void cp_async_commit () { asm volatile ("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait () { asm volatile ("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

template <int PADSHIFT>
__device__ __forceinline__ void stage_tile_async (double* smem, const double* __restrict__ in, int64_t g0, uint32_t count,
                                                  uint64_t dlo, uint64_t dhi, double neutral)
	{
	const uint32_t tid = threadIdx.x, nt = blockDim.x;
	if (g0 >= (int64_t) dlo && g0 + (int64_t) count <= (int64_t) dhi)
		{
		const double* p = in + g0;
		const unsigned int s0 = (unsigned int) __cvta_generic_to_shared (smem);
		for (uint32_t j = tid; j < count; j += nt)
			asm volatile ("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s0 + 8u * stage_idx<PADSHIFT> (j)), "l"(p + j) : "memory");
		}
	else
		{
		for (uint32_t j = tid; j < count; j += nt)
			{
			const int64_t g = g0 + (int64_t) j;
			double v = neutral;
			if (g >= (int64_t) dlo && g < (int64_t) dhi) v = __ldg (in + g);
			smem[stage_idx<PADSHIFT> (j)] = v;
			}
		}
	}

// exact int32 -> double without the I2F.F64 conversion unit (a 16-per-clock-per-SM path that caps a
// whole-genome kernel at ~5 ms): build 2^52 + (x + 2^31) from its bit pattern and subtract the bias
// with one FP64 add (exact: both operands and the result are integers below 2^53)
__device__ __forceinline__ double i32_to_f64 (int x)
	{
	return __dadd_rn (__hiloint2double (0x43300000, x ^ (int) 0x80000000), -4503601774854144.0);
	}
__device__ __forceinline__ double to_f64 (int x)    { return i32_to_f64 (x); }
__device__ __forceinline__ double to_f64 (double x) { return x; }

// order-preserving 64-bit key of a double: a<b  <=>  key(a)<key(b) for all
// non-NaN a,b (with -0.0 just below +0.0)
__device__ __forceinline__ uint64_t f64_key (double v)
	{
	uint64_t b = (uint64_t) __double_as_longlong (v);
	return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
	}

__device__ __forceinline__ double key_f64 (uint64_t k)
	{
	uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
	return __longlong_as_double ((long long) b);
	}

// ----- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------

__device__ __forceinline__ uint32_t smem_u32 (const void* p)
	{ return (uint32_t) __cvta_generic_to_shared (p); }

__device__ __forceinline__ void mbar_init (uint64_t* bar, uint32_t count)
	{
	asm volatile ("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32 (bar)), "r"(count));
	}

__device__ __forceinline__ void mbar_fence_init ()
	{ asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, uint32_t bytes)
	{
	asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
	              :: "r"(smem_u32 (bar)), "r"(bytes) : "memory");
	}

__device__ __forceinline__ void mbar_wait (uint64_t* bar, uint32_t phase)
	{
	asm volatile (
		"{\n"
		".reg .pred p;\n"
		"WAIT_%=:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra DONE_%=;\n"
		"bra WAIT_%=;\n"
		"DONE_%=:\n"
		"}\n" :: "r"(smem_u32 (bar)), "r"(phase) : "memory");
	}

// bytes must be a multiple of 16; src/dst 16-byte aligned
__device__ __forceinline__ void bulk_g2s (void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
	{
	asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	              :: "r"(smem_u32 (dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smem_u32 (bar)) : "memory");
	}

#endif // __CUDACC__
