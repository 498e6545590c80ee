// gdsp_merge.cu -- one step of the percentile operator's "bubble" passes as two merges.
//
// Replaces combine_sorted_vectors (percentile.c:820-864): two chromosomes C and D, each sorted
// ascending, exchange the m smallest cells of D with the m largest of C so that every cell of C is
// <= every cell of D, and both are sorted again.  The reference finds m with a linear scan, swaps, and
// runs qsort on both vectors; sorting the pair jointly gives the same bytes and is what the host
// executor did until now (one radix sort of both chromosomes per step, ~190 B per cell moved).
// Both inputs ARE sorted, so:
//   k_mx_split  m = first k with !(D[k] < C[cLen-1-k])      (the reference's scan, by bisection)
//   k_merge     C' = merge (C[0,cLen-m), D[0,m)),  D' = merge (C[cLen-m,cLen), D[m,dLen))
// each output tile of a merge finds its own corner of the merge path by bisection in global memory,
// stages its two input pieces as keys in shared memory, and every thread merges MG_PER consecutive
// outputs.  16 B per cell moved and no pass over digits.  Order is the total order of f64_key, the
// same as gdsp_sort_genome's (so -0.0 sorts before +0.0 and results equal the joint sort bit for bit).
#include "gdsp_common.cuh"

#define MG_THREADS 256
#define MG_PER     16
#define MG_TILE    (MG_THREADS * MG_PER)

__global__ void k_mx_split (const double* __restrict__ C, uint64_t cLen, const double* __restrict__ D, uint64_t dLen,
                            unsigned long long* __restrict__ out)
	{
	uint64_t lo = 0, hi = (cLen < dLen) ? cLen : dLen;
	while (lo < hi)
		{
		const uint64_t mid = (lo + hi) >> 1;
		if (f64_key (D[mid]) < f64_key (C[cLen - 1 - mid])) lo = mid + 1; else hi = mid;
		}
	*out = lo;
	}

// how many of the first d cells of merge (A, B) come from A (equal keys: A first)
__device__ __forceinline__ uint64_t merge_corner (const double* __restrict__ A, uint64_t na,
                                                  const double* __restrict__ B, uint64_t nb, uint64_t d)
	{
	uint64_t lo = (d > nb) ? d - nb : 0, hi = (d < na) ? d : na;
	while (lo < hi)
		{
		const uint64_t mid = (lo + hi) >> 1;
		if (!(f64_key (B[d - 1 - mid]) < f64_key (A[mid]))) lo = mid + 1; else hi = mid;
		}
	return lo;
	}

__global__ void __launch_bounds__(MG_THREADS)
k_merge (const double* __restrict__ A, uint64_t na, const double* __restrict__ B, uint64_t nb, double* __restrict__ out)
	{
	__shared__ uint64_t s_key[MG_TILE];               // keys of the tile's piece of A, then of B
	__shared__ uint64_t s_corner[2];
	const uint64_t n  = na + nb;
	const uint64_t d0 = (uint64_t) blockIdx.x * MG_TILE;
	const uint64_t d1 = (d0 + MG_TILE < n) ? d0 + MG_TILE : n;
	if (threadIdx.x == 0)  s_corner[0] = merge_corner (A, na, B, nb, d0);
	if (threadIdx.x == 32) s_corner[1] = merge_corner (A, na, B, nb, d1);
	__syncthreads ();
	const uint64_t a0 = s_corner[0], a1 = s_corner[1], b0 = d0 - a0, b1 = d1 - a1;
	const uint32_t la = (uint32_t) (a1 - a0), lb = (uint32_t) (b1 - b0), total = la + lb;
	for (uint32_t j = threadIdx.x; j < la; j += MG_THREADS) s_key[j]      = f64_key (A[a0 + j]);
	for (uint32_t j = threadIdx.x; j < lb; j += MG_THREADS) s_key[la + j] = f64_key (B[b0 + j]);
	__syncthreads ();
	const uint64_t* sA = s_key;
	const uint64_t* sB = s_key + la;

	uint32_t ld = threadIdx.x * MG_PER;
	if (ld > total) ld = total;
	uint32_t lo = (ld > lb) ? ld - lb : 0, hi = (ld < la) ? ld : la;
	while (lo < hi)
		{
		const uint32_t mid = (lo + hi) >> 1;
		if (!(sB[ld - 1 - mid] < sA[mid])) lo = mid + 1; else hi = mid;
		}
	uint32_t ai = lo, bi = ld - lo;
	#pragma unroll
	for (int k = 0; k < MG_PER; k++)
		{
		const uint32_t o = ld + k;
		if (o >= total) break;
		const bool takeA = (bi >= lb) || (ai < la && !(sB[bi] < sA[ai]));
		const uint64_t key = takeA ? sA[ai] : sB[bi];
		if (takeA) ai++; else bi++;
		out[d0 + o] = key_f64 (key);
		}
	}

static int merge_launch (gdsp_ctx* c, const double* A, uint64_t na, const double* B, uint64_t nb, double* out)
	{
	const uint64_t n = na + nb;
	if (n == 0) return GDSP_OK;
	const uint64_t blocks = (n + MG_TILE - 1) / MG_TILE;
	GDSP_REQUIRE (blocks <= 0x7fffffffull, "gdsp_merge_exchange: too many cells for one launch");
	k_merge<<<(unsigned) blocks, MG_THREADS, 0, c->stream>>> (A, na, B, nb, out);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

extern "C" int gdsp_merge_exchange (gdsp_ctx* c, double* sig, double* tmp, uint64_t c_lo, uint64_t c_len,
                                    uint64_t d_lo, uint64_t d_len, uint64_t* h_moved)
	{
	GDSP_REQUIRE (c && sig && tmp && sig != tmp, "gdsp_merge_exchange: NULL or aliased buffers");
	GDSP_REQUIRE (c_lo + c_len <= d_lo || d_lo + d_len <= c_lo, "gdsp_merge_exchange: the two ranges overlap");
	if (h_moved) *h_moved = 0;
	if (c_len == 0 || d_len == 0) return GDSP_OK;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 2, 64, &ws));
	const double* C = sig + c_lo;
	const double* D = sig + d_lo;
	k_mx_split<<<1, 1, 0, c->stream>>> (C, c_len, D, d_len, (unsigned long long*) ws);
	GDSP_KERNEL_CHECK ();
	unsigned long long m = 0;
	GDSP_CUDA (cudaMemcpyAsync (&m, ws, sizeof (m), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (h_moved) *h_moved = m;
	if (m == 0) return GDSP_OK;                           // every cell of C already <= every cell of D
	GDSP_TRY (merge_launch (c, C, c_len - m, D, m, tmp + c_lo));
	GDSP_TRY (merge_launch (c, C + (c_len - m), m, D + m, d_len - m, tmp + d_lo));
	GDSP_CUDA (cudaMemcpyAsync (sig + c_lo, tmp + c_lo, c_len * sizeof (double), cudaMemcpyDeviceToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (sig + d_lo, tmp + d_lo, d_len * sizeof (double), cudaMemcpyDeviceToDevice, c->stream));
	return GDSP_OK;
	}
