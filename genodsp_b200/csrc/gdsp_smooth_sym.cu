// gdsp_smooth_sym.cu -- smooth (Hann FIR, sum.c:616-676) with every product of a symmetric tap
// pair computed ONCE, still in the reference's summation order.
//
// The direct FIR (k_smooth_ct, gdsp_window.cu) spends 2*W separately rounded FP64 instructions per base
// and runs at the FP64 issue limit.  The window is symmetric bit for bit (sum.c:641 stores the same double
// into window[k] and window[W-1-k], :650 divides both by the same sum), so the product w[k]*in[j] is needed
// twice: by output j+h-k at tap k and by output j-h+k at tap W-1-k (h = (W-1)/2).  The two outputs are
// 2(h-k) cells apart and need the product at different points of their ascending-tap folds, which is why a
// thread that owns a few consecutive outputs cannot share it.
//
// This kernel turns the loop inside out.  A thread streams along the INPUT cells of a strip, one cell per
// step, and holds every output that is still being folded in registers: h "young" outputs that are at taps
// 0..h-1 and h "old" ones at the mirrored taps W-1..h+1.  One step multiplies the new cell by the h pair taps
// (h DMUL) and adds each product to one young and one old accumulator (2h DADD); the output leaving the young
// side takes the centre tap (DMUL + DADD) and enters the old side; what leaves the old side is finished.  For
// every output the products arrive k = 0, 1, ..., W-1 in turn, each rounded like the reference's (DMUL, then
// DADD) -- the result is bit-identical with 3h+2 instead of 4h+2 FP64 instructions per base (W = 101: 152
// against 202).  Cells outside the chromosome are zeros, as in the direct kernel (products +0.0 leave an
// accumulator that started at +0.0 unchanged -- the reference skips those taps, sum.c:655-662).
//
// Every accumulator moves one tap per step.  The step loop is unrolled 4 times with the accumulators at fixed
// registers inside a block and one register shift by 4 per block (2(h-1) 64-bit moves per 4 steps on the
// integer pipe); unrolling by h instead would need no moves and 125 KB of code.  Taps are read from the
// constant bank (the kernel parameter), the thread's 4 input cells per block arrive with one 256-bit load
// (next block in flight), its 4 outputs leave with one 256-bit store: no shared memory, no barriers, no
// shuffles.  scripts/smooth_sym_model.py is a CPU model of the accumulator schedule.
//
// (A first design dealt the pairs out to T lanes per strip with shuffles handing the accumulators on; on
// B200 it never beat the direct FIR: 8-byte loads/stores per lane saturated the L1 data pipe, staging through
// shared memory cost more instructions than the multiplies it saved.  profiles/r2_smooth_sym.md.)

#include "gdsp_common.cuh"
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#define SY_KMAX    50                              // pairs held by one thread: W <= 101
#define SY_KUNI(K) ((K) > 44 ? 30 : ((K) > 39 ? 34 : (K)))   // taps kept in uniform registers (63 of them hold 31 doubles; the spills of a few more still fit)
#define SY_THREADS 128
struct SymTaps { double w[SY_KMAX + 1]; };         // w[k], k < h: pair k (= tap k = tap W-1-k); w[h]: centre

__device__ __forceinline__ int sym_strip_seg (const uint64_t* __restrict__ base, int nseg, uint64_t t)
	{
	int lo = 0, hi = nseg - 1;                     // last s with base[s] <= t
	while (lo < hi) { const int mid = (lo + hi + 1) >> 1;  if (__ldg (base + mid) <= t) lo = mid; else hi = mid - 1; }
	return lo;
	}

// The U (4 or 2) input cells of steps n .. n+U-1 go to this thread's slot of the prefetch ring in shared memory
// with 16-byte cp.async (8-byte ones with zero fill where the block leaves the readable range [nLo, nHi)).
#define SY_PD 4                                    // blocks in flight per thread
template <int U>
__device__ __forceinline__ void sym_fetch (const double* __restrict__ pin, uint32_t n, uint32_t nLo, uint32_t nHi, unsigned int slot)
	{
	if (n >= nLo && n + U <= nHi)
		{
		#pragma unroll
		for (int u = 0; u < U; u += 2)
			asm volatile ("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(slot + 8u * u), "l"(pin + n + u) : "memory");
		}
	else
		{
		#pragma unroll
		for (int u = 0; u < U; u++)
			{
			const bool valid = (n + u >= nLo && n + u < nHi);
			const int bytes = valid ? 8 : 0;           // 0: nothing is read, the slot is zero-filled
			asm volatile ("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(slot + 8u * u), "l"(valid ? pin + n + u : pin + nLo), "r"(bytes) : "memory");
			}
		}
	asm volatile ("cp.async.commit_group;" ::: "memory");
	}

// Narrow windows are HBM-bound and their threads have registers to spare: one 256-bit load straight into registers,
// the next block's in flight (hg38 W = 31: 10.9 ms against 12.6 ms through the ring)
__device__ __forceinline__ void sym_load4 (const double* __restrict__ pin, uint32_t n, uint32_t nLo, uint32_t nHi, double (&v)[4])
	{
	if (n >= nLo && n + 4 <= nHi) ldg_stream4 (pin + n, v[0], v[1], v[2], v[3]);
	else
		{
		#pragma unroll
		for (int u = 0; u < 4; u++) v[u] = (n + u >= nLo && n + u < nHi) ? __ldg (pin + n + u) : 0.0;
		}
	}

template <int U>
__device__ __forceinline__ void sym_store (double* __restrict__ pout, uint32_t n, uint32_t out0, uint32_t out1, const double (&e)[U])
	{
	if (n >= out0 && n + U <= out1)
		{
		if (U == 4) stg_stream4 (pout + n, e[0], e[1], e[2], e[U - 1]);
		else asm volatile ("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(pout + n), "d"(e[0]), "d"(e[1]) : "memory");
		}
	else if (n + U > out0 && n < out1)
		{
		#pragma unroll
		for (int u = 0; u < U; u++) if (n + u >= out0 && n + u < out1) pout[n + u] = e[u];
		}
	}

// U = steps per block = cells per vector access.  The accumulator arrays hold K+U-1 doubles each; at 50 pairs the
// register file only has room for U = 2.
template <int K>
__global__ void __launch_bounds__(SY_THREADS, 2)
k_smooth_sym (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t nstrips,
              const double* __restrict__ in, double* __restrict__ out, uint32_t S,
              const __grid_constant__ SymTaps tp)
	{
	constexpr int U     = (K > 40) ? 2 : 4;
	constexpr int LEAD  = (K + 3) / 4 * 4;         // the stream starts LEAD cells before the strip: a multiple of 4, like x0
	constexpr int DELAY = LEAD - K;                // a finished output waits DELAY steps so that stores are aligned too
	constexpr int NH    = (DELAY > 0) ? DELAY : 1;
	constexpr int KUNI  = SY_KUNI (K);
	// taps beyond the uniform register file are read from shared memory (a broadcast LDS per use): ptxas keeps every
	// tap of the loop in a uniform register and, past 63 of them, spills those into the registers the accumulators need
	__shared__ double s_w[(K > KUNI) ? K - KUNI : 1];
	constexpr bool RING = (K > 24);                // input through the cp.async ring (else straight into registers)
	__shared__ __align__(16) double s_ring[RING ? SY_PD : 1][RING ? SY_THREADS : 1][U];
	if (K > KUNI)
		{
		for (int i = threadIdx.x; i < K - KUNI; i += SY_THREADS) s_w[i] = tp.w[KUNI + i];
		__syncthreads ();
		}
	const unsigned int swBase = (unsigned int) __cvta_generic_to_shared (s_w);
	const uint64_t strip = (uint64_t) blockIdx.x * SY_THREADS + threadIdx.x;
	if (strip >= nstrips) return;

	// this thread's strip: outputs [x0, x0+len); step n reads cell x0-LEAD+n and emits output x0-2*LEAD+n
	const int seg = sym_strip_seg (base, nseg, strip);
	const SegDev sd = segs[seg];
	const uint64_t x0  = sd.lo + (strip - __ldg (base + seg)) * S;
	const uint32_t len = (uint32_t) ((sd.hi - x0 < S) ? (sd.hi - x0) : S);
	const int64_t  j0  = (int64_t) x0 - LEAD;
	const int64_t  lo  = ((int64_t) sd.dlo > j0) ? (int64_t) sd.dlo - j0 : 0;
	const int64_t  hi  = (int64_t) sd.dhi - j0;
	const uint32_t total = len + 2 * LEAD;
	const uint32_t nCap = total + (SY_PD + 1) * U;                 // the prefetch runs SY_PD blocks ahead
	const uint32_t nLo = (lo < (int64_t) nCap) ? (uint32_t) lo : nCap;
	const uint32_t nHi = (hi < (int64_t) nCap) ? (uint32_t) ((hi > 0) ? hi : 0) : nCap;
	const double* pin  = in + j0;
	double*       pout = out + (j0 - LEAD);
	const uint32_t out0 = 2 * LEAD, out1 = 2 * LEAD + len;

	double A[K + U - 1], B[K + U - 1];             // young / old accumulators; slot k of sub-step u lives at [k + U-1-u] / [K-1-k + U-1-u]
	#pragma unroll
	for (int i = 0; i < K + U - 1; i++) { A[i] = 0.0;  B[i] = 0.0; }
	double hist[NH];                               // the last DELAY finished outputs
	#pragma unroll
	for (int i = 0; i < NH; i++) hist[i] = 0.0;
	double exitLow = 0.0, cPrev = 0.0;
	const double wc = tp.w[K];

	// prefetch ring: SY_PD blocks in flight, slot (block mod SY_PD) of this thread
	const unsigned int ring = (unsigned int) __cvta_generic_to_shared (&s_ring[0][threadIdx.x][0]);
	constexpr unsigned int SLOT = SY_THREADS * U * 8;
	double cur[4], nxt[4];                         // (U = 4 whenever the ring is not used)
	if (RING)
		{
		#pragma unroll
		for (int q = 0; q < SY_PD; q++) sym_fetch<U> (pin, q * U, nLo, nHi, ring + q * SLOT);
		}
	else sym_load4 (pin, 0, nLo, nHi, cur);
	unsigned int q = 0;
	for (uint32_t n = 0; n < total; n += U)
		{
		if (RING)
			{
			asm volatile ("cp.async.wait_group %0;" :: "n"(SY_PD - 1) : "memory");
			if (U == 4)
				{
				asm volatile ("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cur[0]), "=d"(cur[1]) : "r"(ring + q * SLOT) : "memory");
				asm volatile ("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cur[2]), "=d"(cur[3]) : "r"(ring + q * SLOT + 16u) : "memory");
				}
			else asm volatile ("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(cur[0]), "=d"(cur[1]) : "r"(ring + q * SLOT) : "memory");
			sym_fetch<U> (pin, n + SY_PD * U, nLo, nHi, ring + q * SLOT);
			q = (q + 1) & (SY_PD - 1);
			}
		else sym_load4 (pin, n + U, nLo, nHi, nxt);
		double fin[U];
		#pragma unroll
		for (int u = 0; u < U; u++)
			{
			const double v = cur[u];
			// the output that left the young side one step ago takes the centre tap now and enters the old side next step
			const double cNew = __dadd_rn (exitLow, __dmul_rn (wc, v));
			A[U - 1 - u] = 0.0;
			B[U - 1 - u] = cPrev;
			cPrev = cNew;
			#pragma unroll
			for (int k = 0; k < K; k++)
				{
				double wk;
				if (k < KUNI) wk = tp.w[k];
				else asm volatile ("ld.shared.f64 %0, [%1];" : "=d"(wk) : "r"(swBase + 8u * (unsigned int) ((k < KUNI) ? 0 : k - KUNI)));
				const double p = __dmul_rn (wk, v);
				A[k + U - 1 - u]         = __dadd_rn (A[k + U - 1 - u], p);
				B[K - 1 - k + U - 1 - u] = __dadd_rn (B[K - 1 - k + U - 1 - u], p);
				}
			exitLow = A[K - 1 + U - 1 - u];
			fin[u]  = B[K - 1 + U - 1 - u];
			}
		// every accumulator is one tap further per step: shift by U (the moves fold into the adds' destinations)
		#pragma unroll
		for (int i = K + U - 2; i >= U; i--) { A[i] = A[i - U];  B[i] = B[i - U]; }

		// emit what finished DELAY steps ago: U outputs = one aligned vector store
		double em[U];
		#pragma unroll
		for (int u = 0; u < U; u++) em[u] = (u >= DELAY) ? fin[(u >= DELAY) ? u - DELAY : 0] : hist[(u < NH) ? NH - DELAY + u : 0];
		// hist keeps the last DELAY finished outputs, oldest first
		if (DELAY > 0)
			{
			double nh[NH];
			#pragma unroll
			for (int i = 0; i < DELAY; i++)
				{
				// finished at step n + U - DELAY + i: from this block if that index is >= 0, else from the history
				const int f = U - DELAY + i;
				nh[i] = (f >= 0) ? fin[(f >= 0) ? f : 0] : hist[(f < 0) ? i + U : 0];
				}
			#pragma unroll
			for (int i = 0; i < DELAY; i++) hist[i] = nh[i];
			}
		sym_store<U> (pout, n, out0, out1, em);
		if (!RING)
			{
			#pragma unroll
			for (int u = 0; u < 4; u++) cur[u] = nxt[u];
			}
		}
	}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

extern "C" int gdsp_ctx_set_smooth_direct (gdsp_ctx* c, int on)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_ctx_set_smooth_direct: NULL context");
	c->smooth_direct = on ? 1 : 0;
	return GDSP_OK;
	}

typedef void (*sym_kernel_t) (const SegDev*, const uint64_t*, int, uint64_t, const double*, double*, uint32_t, const SymTaps);

template <int K> struct SymTable
	{
	static void fill (sym_kernel_t* t) { t[K] = k_smooth_sym<K>;  SymTable<K - 1>::fill (t); }
	};
template <> struct SymTable<0> { static void fill (sym_kernel_t*) {} };

// Does this window suit the shared-product kernel?  Odd, at most 2*SY_KMAX+1 wide, symmetric bit for bit;
// below 9 taps the direct FIR is at the HBM floor anyway.
int gdsp_smooth_sym_plan (const gdsp_layout* L, const double* in, uint32_t W, const double* h_taps)
	{
	if (W < 9 || (W & 1u) == 0) return 0;
	const uint32_t h = (W - 1) / 2;
	if (h > SY_KMAX) return 0;
	for (uint32_t k = 0; k < h; k++)
		if (memcmp (&h_taps[k], &h_taps[W - 1 - k], sizeof (double)) != 0) return 0;
	if ((((uintptr_t) in) & 31u) != 0) return 0;           // 256-bit loads
	for (int s = 0; s < L->nseg; s++) if (L->h[s].lo & 3u) return 0;
	return 1;
	}

int gdsp_smooth_sym_launch (gdsp_ctx* c, gdsp_layout* L, const double* in, double* out, uint32_t W, const double* h_taps)
	{
	static sym_kernel_t table[SY_KMAX + 1];
	static bool filled = false;
	if (!filled) { SymTable<SY_KMAX>::fill (table);  filled = true; }
	const int K = (int) ((W - 1) / 2);
	SymTaps tp;
	memset (&tp, 0, sizeof (tp));
	memcpy (tp.w, h_taps, sizeof (double) * (K + 1));
	// Strip length: one thread per strip, 2 CTAs of 128 threads per SM at once; the 2*LEAD ramp steps of a strip are
	// its overhead.  Whole waves of strips of about 8192 cells (ramp 1.3 % at W = 101); short genomes get short strips.
	const uint64_t slots = (uint64_t) c->sm_count * 2 * SY_THREADS;
	uint64_t waves = (L->cells + slots * 8192 / 2) / (slots * 8192);
	if (waves < 1) waves = 1;
	uint64_t S = (L->cells + slots * waves - 1) / (slots * waves);
	S = (S + 63) / 64 * 64;
	if (S < 512) S = 512;
	const char* forceS = getenv ("GDSP_SYM_STRIP");          // measurement override
	if (forceS != NULL && atoi (forceS) >= 64) S = (uint64_t) atoi (forceS) / 64 * 64;
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, (uint32_t) S, &tm));
	const uint32_t S32 = (uint32_t) S;
	if (tm.ntiles == 0) return GDSP_OK;                          // a layout without cells
	const uint64_t blocks = (tm.ntiles + SY_THREADS - 1) / SY_THREADS;
	sym_kernel_t kern = table[K];
	void* args[] = { (void*) &L->d, (void*) &tm.d_base, (void*) &L->nseg, (void*) &tm.ntiles, (void*) &in, (void*) &out,
	                 (void*) &S32, (void*) &tp };
	GDSP_CUDA (cudaLaunchKernel ((const void*) kern, dim3 ((unsigned) blocks), dim3 (SY_THREADS), args, 0, c->stream));
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
