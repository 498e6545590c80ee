// gdsp_pointwise.cu -- fused pointwise operator programs, interval tables and
// the min/max reduction.
//
// One launch applies a whole chain of consecutive pointwise operators per cell
// (binarize, addconst, abs, clip, erase, invert and the interval-file operators
// add/subtract/multiply/divide/mask/masknot/or/and), so a chain of any length
// costs one 8-byte read and one 8-byte write per base instead of one pass per
// operator as in the reference's executor loop (genodsp.c:909-921).
// Reference loops replaced: logical.c:250-263, add.c:736-739, add.c:1046-1047,
// mask.c:893-913, mask.c:1182-1228, add.c:936, logical.c:466-473, add.c:280-281,
// multiply.c:312-329, multiply.c:705-722, mask.c:295-296, mask.c:591-599,
// logical.c:or/and loops.
#include "gdsp_common.cuh"

struct gdsp_ivl_table
	{
	gdsp_ctx* ctx;
	uint64_t  n;
	uint64_t* d_start;      // buffer cell index, ascending
	uint64_t* d_end;
	double*   d_val;
	};

struct PwOpDev
	{
	int32_t  code;
	uint32_t flags;
	double   a, b, c;
	const uint64_t* start;
	const uint64_t* end;
	const double*   val;
	uint64_t        n;
	const uint2*    tix;     // per tile of this launch: x = first table entry that can touch the tile, y = how many
	};

struct PwProgram
	{
	int     nops;
	int     hasIvl;
	PwOpDev ops[GDSP_MAX_POINTWISE];
	};

#define PW_THREADS 256
#define PW_TILE    4096


// One plain (table-free) operator on N cells held in registers.  The opcode dispatch is a warp-uniform
// switch paid once per N cells instead of once per cell.
template <int N>
__device__ __forceinline__ void pw_plain_op (const PwOpDev& op, double (&v)[N])
	{
	const double a = op.a, b = op.b, c = op.c;
	switch (op.code)
		{
		case GDSP_PW_BINARIZE_GT:
			#pragma unroll
			for (int e = 0; e < N; e++) v[e] = (v[e] >  a) ? b : c;
			break;
		case GDSP_PW_BINARIZE_GE:
			#pragma unroll
			for (int e = 0; e < N; e++) v[e] = (v[e] >= a) ? b : c;
			break;
		case GDSP_PW_ADDCONST:
			#pragma unroll
			for (int e = 0; e < N; e++) v[e] = __dadd_rn (v[e], a);
			break;
		case GDSP_PW_ABS:
			#pragma unroll
			for (int e = 0; e < N; e++) if (v[e] < 0) v[e] = -v[e];
			break;
		case GDSP_PW_CLIP_MIN:
			#pragma unroll
			for (int e = 0; e < N; e++) if (v[e] < a) v[e] = a;
			break;
		case GDSP_PW_CLIP_MAX:
			#pragma unroll
			for (int e = 0; e < N; e++) if (v[e] > a) v[e] = a;
			break;
		case GDSP_PW_CLIP_BOTH:
			#pragma unroll
			for (int e = 0; e < N; e++) { if (v[e] < a) v[e] = a; else if (v[e] > b) v[e] = b; }
			break;
		case GDSP_PW_ERASE:
			{
			const bool hmin = op.flags & GDSP_PW_ERASE_HAVE_MIN, hmax = op.flags & GDSP_PW_ERASE_HAVE_MAX;
			const bool keepIn = op.flags & GDSP_PW_ERASE_KEEP_INSIDE;
			#pragma unroll
			for (int e = 0; e < N; e++)
				{
				bool kill;
				if (keepIn) kill = (hmin && v[e] < a) || (hmax && v[e] > b);
				else        kill = (!hmin || v[e] >= a) && (!hmax || v[e] <= b);
				if (kill) v[e] = c;
				}
			break;
			}
		case GDSP_PW_INVERT:
			#pragma unroll
			for (int e = 0; e < N; e++) v[e] = __dsub_rn (a, v[e]);
			break;
		case GDSP_PW_NONZERO_TO_ONE:
			#pragma unroll
			for (int e = 0; e < N; e++) if (v[e] != 0.0) v[e] = 1.0;
			break;
		default: break;
		}
	}

#define PW_VEC 8

// programs without interval-table operators (few registers: these chains run at the HBM rate)
__global__ void __launch_bounds__(PW_THREADS, 3)
k_pointwise (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
             const double* __restrict__ in, double* __restrict__ out, const __grid_constant__ PwProgram P)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * PW_TILE;
	uint64_t t1 = t0 + PW_TILE;  if (t1 > sd.hi) t1 = sd.hi;

	// PW_TILE = PW_THREADS * 16: every thread owns PW_VEC/2 pairs per half tile, pair p of the
	// warp-wide access q at cell t0 + 2*(q*PW_THREADS + tid): 128-bit coalesced accesses
	#pragma unroll 1
	for (uint32_t half = 0; half < PW_TILE / (PW_THREADS * PW_VEC); half++)
		{
		double v[PW_VEC];
		const uint64_t h0 = t0 + (uint64_t) half * (PW_THREADS * PW_VEC);
		if (h0 >= t1) break;
		#pragma unroll
		for (int q = 0; q < PW_VEC / 2; q++)
			{
			const uint64_t i = h0 + 2 * ((uint64_t) q * PW_THREADS + threadIdx.x);
			if (i + 1 < t1) { double2 x = ldg_stream (in + i);  v[2*q] = x.x;  v[2*q+1] = x.y; }
			else            { v[2*q] = (i < t1) ? in[i] : 0.0;  v[2*q+1] = 0.0; }
			}
		for (int i = 0; i < P.nops; i++) pw_plain_op<PW_VEC> (P.ops[i], v);
		#pragma unroll
		for (int q = 0; q < PW_VEC / 2; q++)
			{
			const uint64_t i = h0 + 2 * ((uint64_t) q * PW_THREADS + threadIdx.x);
			if (i + 1 < t1) stg_stream (out + i, make_double2 (v[2*q], v[2*q+1]));
			else if (i < t1) out[i] = v[2*q];
			}
		}
	}

// ---------------------------------------------------------------------------
// Programs with interval-table operators (add/subtract/multiply/divide/mask/masknot/or/and/...).
// The first version searched the table once per cell (two dependent global loads per step, then end[k]
// and val[k]): a five-operator chain ran at a fifth of the plain chains' rate.  Now
//   k_ivl_tile_index  finds, once per launch and table, the slice of table entries that can touch
//                     every tile (all tiles in parallel, so the search latency is hidden), and
//   k_pointwise_ivl   PAINTS the slice into a shared-memory array of one 16-bit entry index per cell
//                     (warp per interval for sparse slices, thread per interval for dense ones) and
//                     then reads one entry per cell -- no search; consecutive operators on the same
//                     table (add B = multiply B = and B) share the painted array.
// ---------------------------------------------------------------------------

// index of the last interval with start <= g inside [lo,hi), or lo-1
__device__ __forceinline__ int64_t ivl_find (const uint64_t* __restrict__ start, int64_t lo, int64_t hi, uint64_t g)
	{
	while (lo < hi)
		{
		int64_t mid = (lo + hi) >> 1;
		if (start[mid] <= g) lo = mid + 1; else hi = mid;
		}
	return lo - 1;
	}

__global__ void __launch_bounds__(256)
k_ivl_tile_index (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, uint64_t ntiles,
                  const uint64_t* __restrict__ start, uint64_t n, uint2* __restrict__ tix)
	{
	const uint64_t tile = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (tile >= ntiles) return;
	int lo = 0, hi = nseg - 1;                        // segment of the tile (plain bisection: one thread per tile)
	while (lo < hi)
		{
		const int mid = (lo + hi + 1) >> 1;
		if (base[mid] <= tile) lo = mid; else hi = mid - 1;
		}
	const SegDev sd = segs[lo];
	const uint64_t t0 = sd.lo + (tile - base[lo]) * PW_TILE;
	uint64_t t1 = t0 + PW_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	// the last entry starting at or before t0 (it may cover t0) up to the last entry starting before t1
	int64_t a = ivl_find (start, 0, (int64_t) n, t0);
	const int64_t b = ivl_find (start, 0, (int64_t) n, t1 - 1);
	if (a < 0) a = 0;
	tix[tile] = make_uint2 ((unsigned) a, (unsigned) (b + 1 - a));
	}

#define PW_EARLY     64         // slice entries of the first table staged early
#define PW_IVL_CELLS 16         // cells per thread: the whole tile stays in registers across the program

__global__ void __launch_bounds__(PW_THREADS, 3)
k_pointwise_ivl (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                 const double* __restrict__ in, double* __restrict__ out, const __grid_constant__ PwProgram P)
	{
	__shared__ __align__(16) short s_idx[PW_TILE];    // table entry (relative to the slice) covering the cell, or -1
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * PW_TILE;
	uint64_t t1 = t0 + PW_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

	// The slice of the program's FIRST table is fetched while the signal loads are in flight: its tile
	// index entry first, then (sparse slices) the clipped bounds of its intervals into shared memory --
	// otherwise two dependent global round trips sit between the barriers of the paint step of every tile
	__shared__ uint32_t s_ca[PW_EARLY], s_cb[PW_EARLY];
	int firstIvl = 0;
	while (firstIvl < P.nops && P.ops[firstIvl].code < GDSP_PW_IVL_ADD) firstIvl++;       // < nops: the program has one
	const uint2 tx0 = P.ops[firstIvl].tix[blockIdx.x];

	// pair q of the thread at cell t0 + 2*(q*PW_THREADS + tid): 128-bit coalesced accesses
	double v[PW_IVL_CELLS];
	#pragma unroll
	for (int q = 0; q < PW_IVL_CELLS / 2; q++)
		{
		const uint64_t i = t0 + 2 * ((uint64_t) q * PW_THREADS + threadIdx.x);
		if (i + 1 < t1) { double2 x = ldg_stream (in + i);  v[2*q] = x.x;  v[2*q+1] = x.y; }
		else            { v[2*q] = (i < t1) ? in[i] : 0.0;  v[2*q+1] = 0.0; }
		}
	if (tx0.y <= PW_EARLY && threadIdx.x < tx0.y)
		{
		const uint64_t st = P.ops[firstIvl].start[tx0.x + threadIdx.x], en = P.ops[firstIvl].end[tx0.x + threadIdx.x];
		s_ca[threadIdx.x] = (uint32_t) ((st > t0 ? st : t0) - t0);
		s_cb[threadIdx.x] = (en <= t0) ? 0u : (uint32_t) ((en < t1 ? en : t1) - t0);
		}

	const uint64_t* painted = NULL;                   // table whose slice is in s_idx
	uint32_t klo = 0;
	for (int i = 0; i < P.nops; i++)
		{
		const PwOpDev& op = P.ops[i];
		if (op.code < GDSP_PW_IVL_ADD) { pw_plain_op<PW_IVL_CELLS> (op, v);  continue; }
		if (op.start != painted)
			{
			painted = op.start;
			__syncthreads ();                         // readers of the previous table are done
			#pragma unroll
			for (int j = 0; j < PW_TILE / 8 / PW_THREADS; j++)
				reinterpret_cast<uint4*> (s_idx)[j * PW_THREADS + threadIdx.x] = make_uint4 (~0u, ~0u, ~0u, ~0u);
			__syncthreads ();
			const uint2 tx = (i == firstIvl) ? tx0 : op.tix[blockIdx.x];
			klo = tx.x;
			const uint32_t cnt = tx.y;
			if (i == firstIvl && cnt <= PW_EARLY)
				{
				// bounds already in shared memory (the barrier above ordered their stores)
				for (uint32_t j = warp; j < cnt; j += PW_THREADS / 32)
					{
					const uint32_t ca = s_ca[j], cb = s_cb[j];
					for (uint32_t cc = ca + lane; cc < cb; cc += 32) s_idx[cc] = (short) j;
					}
				}
			else if (cnt > 128)
				{
				// dense slice (short intervals): one thread per entry
				for (uint32_t j = threadIdx.x; j < cnt; j += PW_THREADS)
					{
					const uint64_t st = op.start[klo + j], en = op.end[klo + j];
					const uint32_t ca = (uint32_t) ((st > t0 ? st : t0) - t0);
					const uint32_t cb = (en <= t0) ? 0u : (uint32_t) ((en < t1 ? en : t1) - t0);
					for (uint32_t cc = ca; cc < cb; cc++) s_idx[cc] = (short) j;
					}
				}
			else
				{
				for (uint32_t j = warp; j < cnt; j += PW_THREADS / 32)
					{
					const uint64_t st = op.start[klo + j], en = op.end[klo + j];
					const uint32_t ca = (uint32_t) ((st > t0 ? st : t0) - t0);
					const uint32_t cb = (en <= t0) ? 0u : (uint32_t) ((en < t1 ? en : t1) - t0);
					for (uint32_t cc = ca + lane; cc < cb; cc += 32) s_idx[cc] = (short) j;
					}
				}
			__syncthreads ();
			}
		const double a = op.a;
		const double* __restrict__ val = op.val + klo;
		// entry index of each cell (-1 outside every interval): one 32-bit shared-memory load per pair of cells
		#define PW_EACH(body) _Pragma("unroll") for (int q = 0; q < PW_IVL_CELLS / 2; q++) \
			{ const uint32_t pairIdx = reinterpret_cast<const uint32_t*> (s_idx)[q * PW_THREADS + threadIdx.x]; \
			  _Pragma("unroll") for (int h = 0; h < 2; h++) \
				{ const int e = 2 * q + h;  const int j = (int) (short) (pairIdx >> (16 * h));  const bool inside = (j >= 0);  body } }
		switch (op.code)
			{
			case GDSP_PW_IVL_ADD: PW_EACH (if (inside) v[e] = __dadd_rn (v[e], val[j]);)  break;
			case GDSP_PW_IVL_SUB: PW_EACH (if (inside) v[e] = __dsub_rn (v[e], val[j]);)  break;
			case GDSP_PW_IVL_MUL: PW_EACH (v[e] = inside ? __dmul_rn (v[e], val[j]) : a;)  break;
			case GDSP_PW_IVL_DIV: PW_EACH (v[e] = inside ? __ddiv_rn (v[e], val[j]) : ((v[e] >= 0) ? a : -a);)  break;
			case GDSP_PW_IVL_SET: PW_EACH (if (inside) v[e] = a;)  break;
			case GDSP_PW_IVL_SET_OUTSIDE: PW_EACH (if (!inside) v[e] = a;)  break;
			case GDSP_PW_IVL_ASSIGN: PW_EACH (if (inside) v[e] = val[j];)  break;
			case GDSP_PW_IVL_MIN: PW_EACH (if (inside) { const double w = val[j];  if (w < v[e]) v[e] = w; })  break;
			case GDSP_PW_IVL_MAX: PW_EACH (if (inside) { const double w = val[j];  if (w > v[e]) v[e] = w; })  break;
			case GDSP_PW_IVL_KEEP_AT:
				PW_EACH (const uint64_t g = t0 + 2 * ((uint64_t) (e >> 1) * PW_THREADS + threadIdx.x) + (e & 1);
				         if (!(inside && (double) g == val[j])) v[e] = a;)
				break;
			case GDSP_PW_IVL_ACCUM_CLEAR: PW_EACH (if (inside) v[e] = (v[e] == a) ? val[j] : __dadd_rn (v[e], val[j]);)  break;
			default: break;
			}
		#undef PW_EACH
		}

	#pragma unroll
	for (int q = 0; q < PW_IVL_CELLS / 2; q++)
		{
		const uint64_t i = t0 + 2 * ((uint64_t) q * PW_THREADS + threadIdx.x);
		if (i + 1 < t1) stg_stream (out + i, make_double2 (v[2*q], v[2*q+1]));
		else if (i < t1) out[i] = v[2*q];
		}
	}

// ---------------------------------------------------------------------------
// min / max / count reduction (strided, range-filtered)
// ---------------------------------------------------------------------------

#define MMR_TILE 8192

__global__ void __launch_bounds__(256)
k_minmax (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
          const double* __restrict__ sig, uint32_t stride, double mnAllowed, double mxAllowed,
          unsigned long long* __restrict__ res /* [0]=minKey [1]=maxKey [2]=count */)
	{
	__shared__ unsigned long long s_mn[8], s_mx[8], s_ct[8];
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * MMR_TILE;
	uint64_t t1 = t0 + MMR_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	unsigned long long mn = ~0ull, mx = 0ull, ct = 0ull;
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
		{
		if (stride > 1 && ((uint64_t) sd.pos0 + (i - sd.lo)) % stride != 0) continue;
		double v = sig[i];
		if (v < mnAllowed) continue;
		if (v > mxAllowed) continue;
		ct++;
		if (v != v) continue;
		unsigned long long k = f64_key (v);
		if (k < mn) mn = k;
		if (k > mx) mx = k;
		}
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1)
		{
		unsigned long long a = __shfl_xor_sync (0xffffffffu, mn, d);  if (a < mn) mn = a;
		unsigned long long b = __shfl_xor_sync (0xffffffffu, mx, d);  if (b > mx) mx = b;
		ct += __shfl_xor_sync (0xffffffffu, ct, d);
		}
	if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn;  s_mx[threadIdx.x >> 5] = mx;  s_ct[threadIdx.x >> 5] = ct; }
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		for (int w = 1; w < 8; w++)
			{
			if (s_mn[w] < mn) mn = s_mn[w];
			if (s_mx[w] > mx) mx = s_mx[w];
			ct += s_ct[w];
			}
		if (ct != 0)
			{
			atomicMin (&res[0], mn);
			atomicMax (&res[1], mx);
			atomicAdd (&res[2], ct);
			}
		}
	}

// cells that are NOT integers of magnitude <= limit (NaN and infinities count): when there are none,
// adding integer interval values in any order gives the reference's bits (add.c:280-281)
__global__ void __launch_bounds__(256)
k_count_non_integer (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
                     const double* __restrict__ sig, double limit, unsigned long long* __restrict__ res)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * MMR_TILE;
	uint64_t t1 = t0 + MMR_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	unsigned int bad = 0;
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
		{
		const double v = sig[i];
		if (!(fabs (v) <= limit) || v != rint (v)) bad++;
		}
	bad = __reduce_add_sync (0xffffffffu, bad);
	if ((threadIdx.x & 31) == 0 && bad != 0) atomicAdd (res, (unsigned long long) bad);
	}

// ---------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------

extern "C" int gdsp_ivl_table_create (gdsp_ctx* c, const gdsp_layout* L_, const uint32_t* h_seg,
                                      const uint32_t* h_start, const uint32_t* h_end, const double* h_val,
                                      uint64_t n, gdsp_ivl_table** out)
	{
	const gdsp_layout* L = L_;
	GDSP_REQUIRE (c && L && out, "gdsp_ivl_table_create: NULL argument");
	GDSP_REQUIRE (n == 0 || (h_seg && h_start && h_end), "gdsp_ivl_table_create: NULL interval arrays");
	std::vector<uint64_t> s, e;
	std::vector<double> v;
	s.reserve (n);  e.reserve (n);  v.reserve (n);
	uint64_t prevEnd = 0;
	for (uint64_t k = 0; k < n; k++)
		{
		GDSP_REQUIRE (h_seg[k] < (uint32_t) L->nseg, "gdsp_ivl_table_create: interval %llu names segment %u of %d",
		              (unsigned long long) k, h_seg[k], L->nseg);
		const gdsp_seg& g = L->h[h_seg[k]];
		uint64_t p0 = g.pos0, p1 = p0 + (g.hi - g.lo);
		if (h_start[k] >= h_end[k]) continue;
		if ((uint64_t) h_end[k] <= p0 || (uint64_t) h_start[k] >= p1) continue;       // not on this piece
		uint64_t a = g.lo + (((uint64_t) h_start[k] > p0 ? (uint64_t) h_start[k] : p0) - p0);
		uint64_t b = g.lo + (((uint64_t) h_end[k]   < p1 ? (uint64_t) h_end[k]   : p1) - p0);
		GDSP_REQUIRE (a >= prevEnd, "gdsp_ivl_table_create: intervals must be sorted in layout order and disjoint (entry %llu)",
		              (unsigned long long) k);
		s.push_back (a);  e.push_back (b);  v.push_back (h_val ? h_val[k] : 1.0);
		prevEnd = b;
		}
	gdsp_ivl_table* t = new gdsp_ivl_table ();
	t->ctx = c;  t->n = s.size ();  t->d_start = t->d_end = NULL;  t->d_val = NULL;
	size_t m = t->n ? t->n : 1;
	cudaSetDevice (c->device);
	cudaError_t er = cudaMalloc (&t->d_start, m * sizeof (uint64_t));
	if (er == cudaSuccess) er = cudaMalloc (&t->d_end, m * sizeof (uint64_t));
	if (er == cudaSuccess) er = cudaMalloc (&t->d_val, m * sizeof (double));
	if (er == cudaSuccess && t->n)
		{
		er = cudaMemcpyAsync (t->d_start, s.data (), t->n * sizeof (uint64_t), cudaMemcpyHostToDevice, c->stream);
		if (er == cudaSuccess) er = cudaMemcpyAsync (t->d_end, e.data (), t->n * sizeof (uint64_t), cudaMemcpyHostToDevice, c->stream);
		if (er == cudaSuccess) er = cudaMemcpyAsync (t->d_val, v.data (), t->n * sizeof (double), cudaMemcpyHostToDevice, c->stream);
		if (er == cudaSuccess) er = cudaStreamSynchronize (c->stream);
		}
	if (er != cudaSuccess)
		{
		gdsp_set_error ("gdsp_ivl_table_create: %s", cudaGetErrorString (er));
		gdsp_ivl_table_destroy (t);
		return GDSP_ERR_CUDA;
		}
	*out = t;
	return GDSP_OK;
	}

extern "C" void gdsp_ivl_table_destroy (gdsp_ivl_table* t)
	{
	if (t == NULL) return;
	cudaStreamSynchronize (t->ctx->stream);
	if (t->d_start) cudaFree (t->d_start);
	if (t->d_end)   cudaFree (t->d_end);
	if (t->d_val)   cudaFree (t->d_val);
	delete t;
	}

extern "C" int gdsp_pointwise (gdsp_ctx* c, const gdsp_layout* L_, const double* in, double* out,
                               const gdsp_pw_op* ops, int nops)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && in && out && ops, "gdsp_pointwise: NULL argument");
	GDSP_REQUIRE (nops >= 1 && nops <= GDSP_MAX_POINTWISE, "gdsp_pointwise: %d operators (1..%d allowed)", nops, GDSP_MAX_POINTWISE);
	PwProgram P;
	memset (&P, 0, sizeof (P));
	P.nops = nops;
	for (int i = 0; i < nops; i++)
		{
		GDSP_REQUIRE (ops[i].code >= GDSP_PW_BINARIZE_GT && ops[i].code <= GDSP_PW_IVL_ACCUM_CLEAR,
		              "gdsp_pointwise: operator %d has unknown code %d", i, ops[i].code);
		P.ops[i].code = ops[i].code;  P.ops[i].flags = ops[i].flags;
		P.ops[i].a = ops[i].a;  P.ops[i].b = ops[i].b;  P.ops[i].c = ops[i].c;
		if (ops[i].code >= GDSP_PW_IVL_ADD)
			{
			GDSP_REQUIRE (ops[i].table != NULL, "gdsp_pointwise: operator %d needs an interval table", i);
			P.ops[i].start = ops[i].table->d_start;  P.ops[i].end = ops[i].table->d_end;
			P.ops[i].val = ops[i].table->d_val;      P.ops[i].n = ops[i].table->n;
			P.hasIvl = 1;
			}
		}
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, PW_TILE, &tm));
	if (tm.ntiles == 0) return GDSP_OK;
	if (!P.hasIvl)
		{
		k_pointwise<<<(unsigned) tm.ntiles, PW_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, P);
		GDSP_KERNEL_CHECK ();
		return GDSP_OK;
		}
	// one tile index per distinct table of the program
	const uint64_t* distinct[GDSP_MAX_POINTWISE];
	int nd = 0;
	for (int i = 0; i < nops; i++)
		if (P.ops[i].code >= GDSP_PW_IVL_ADD)
			{
			int d = 0;
			while (d < nd && distinct[d] != P.ops[i].start) d++;
			if (d == nd) distinct[nd++] = P.ops[i].start;
			}
	void* ws;
	GDSP_TRY (gdsp_ws (c, 3, (size_t) nd * tm.ntiles * sizeof (uint2), &ws));
	for (int d = 0; d < nd; d++)
		{
		uint2* tix = (uint2*) ws + (size_t) d * tm.ntiles;
		uint64_t n = 0;
		for (int i = 0; i < nops; i++)
			if (P.ops[i].code >= GDSP_PW_IVL_ADD && P.ops[i].start == distinct[d]) { P.ops[i].tix = tix;  n = P.ops[i].n; }
		GDSP_REQUIRE (n < 0xffffffffull, "gdsp_pointwise: interval table with 2^32 or more entries");
		k_ivl_tile_index<<<(unsigned) ((tm.ntiles + 255) / 256), 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, tm.ntiles, distinct[d], n, tix);
		GDSP_KERNEL_CHECK ();
		}
	k_pointwise_ivl<<<(unsigned) tm.ntiles, PW_THREADS, 0, c->stream>>> (L->d, tm.d_base, L->nseg, in, out, P);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

extern "C" int gdsp_minmax (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, uint32_t stride,
                            double mnAllowed, double mxAllowed, double* h_min, double* h_max, uint64_t* h_count)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig, "gdsp_minmax: NULL argument");
	if (stride == 0) stride = 1;
	void* ws;
	GDSP_TRY (gdsp_ws (c, 2, 64, &ws));
	unsigned long long init[3] = { ~0ull, 0ull, 0ull };
	GDSP_CUDA (cudaMemcpyAsync (ws, init, sizeof (init), cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, MMR_TILE, &tm));
	k_minmax<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, stride, mnAllowed, mxAllowed,
	                                                      (unsigned long long*) ws);
	GDSP_KERNEL_CHECK ();
	unsigned long long res[3];
	GDSP_CUDA (cudaMemcpyAsync (res, ws, sizeof (res), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (h_count) *h_count = res[2];
	// keys back to doubles on the host (same transform as f64_key/key_f64)
	auto unkey = [] (unsigned long long k) -> double
		{
		unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
		double d;  memcpy (&d, &b, 8);  return d;
		};
	if (h_min) *h_min = (res[0] == ~0ull && res[1] == 0ull) ? DBL_MAX  : unkey (res[0]);
	if (h_max) *h_max = (res[0] == ~0ull && res[1] == 0ull) ? -DBL_MAX : unkey (res[1]);
	return GDSP_OK;
	}

extern "C" int gdsp_count_non_integer (gdsp_ctx* c, const gdsp_layout* L_, const double* sig, double limit, uint64_t* h_count)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_count, "gdsp_count_non_integer: NULL argument");
	void* ws;
	GDSP_TRY (gdsp_ws (c, 2, 64, &ws));
	GDSP_CUDA (cudaMemsetAsync (ws, 0, sizeof (unsigned long long), c->stream));
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, MMR_TILE, &tm));
	if (tm.ntiles != 0)
		{
		k_count_non_integer<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, limit, (unsigned long long*) ws);
		GDSP_KERNEL_CHECK ();
		}
	unsigned long long res = 0;
	GDSP_CUDA (cudaMemcpyAsync (&res, ws, sizeof (res), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	*h_count = res;
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// percentile --preserve: the reference saves the signal as text with 10 decimals
// (write_all_chromosomes, genodsp.c:1754-1775: report_intervals, precision 10, zero
// runs hidden) and reads it back after the percentile (read_all_chromosomes,
// genodsp.c:1717-1742: read_intervals with clear), so every value comes back as
//     strtod (printf ("%.10f", v))            -- and every zero as +0.0.
// k_text_roundtrip computes exactly that without any text: with |v| = m * 2^-k,
// q = round-half-even (m * 10^10 / 2^k) is what glibc prints (exact arithmetic on the
// binary value), and the nearest double to q / 10^10 (ties to even) is what strtod
// returns.  All of it in 128-bit integers.  inf comes back as DBL_MAX ("inf" goes
// through string_to_double, utilities.c:350-352), NaN as the default quiet NaN.
// ---------------------------------------------------------------------------

__device__ __forceinline__ int bitlen128 (unsigned __int128 x)
	{
	const unsigned long long hi = (unsigned long long) (x >> 64), lo = (unsigned long long) x;
	return hi ? 128 - __clzll ((long long) hi) : 64 - __clzll ((long long) lo);
	}

__device__ double text_roundtrip10 (double v)
	{
	if (v == 0.0) return 0.0;
	const unsigned long long SIGN = 0x8000000000000000ull;
	const unsigned long long b = (unsigned long long) __double_as_longlong (v);
	const unsigned long long sign = b & SIGN, ab = b & ~SIGN;
	const int ex = (int) (ab >> 52);
	const unsigned long long frac = ab & 0x000fffffffffffffull;
	if (ex == 0x7ff)
		return __longlong_as_double ((long long) (sign | (frac ? 0x7ff8000000000000ull : 0x7fefffffffffffffull)));
	unsigned long long m;  int e;
	if (ex == 0) { m = frac;  e = -1074; } else { m = frac | 0x0010000000000000ull;  e = ex - 1075; }
	if (e >= 0) return v;                                       // an integer: printed exactly
	const int k = -e;                                           // |v| = m / 2^k
	if (k <= 52 && (m & ((1ull << k) - 1ull)) == 0ull) return v;   // still an integer
	const double zero = __longlong_as_double ((long long) sign);   // "-0.0000000000" reads back as -0.0
	if (k >= 120) return zero;
	const unsigned long long D = 10000000000ull;
	const unsigned __int128 P = (unsigned __int128) m * D;      // < 2^87
	unsigned __int128 q = P >> k;
	const unsigned __int128 rem = P & ((((unsigned __int128) 1) << k) - 1), half = ((unsigned __int128) 1) << (k - 1);
	if (rem > half || (rem == half && (q & 1))) q += 1;
	if (q == 0) return zero;
	// nearest double to q / D: scale q to 89 bits so the quotient has 55 or 56 bits
	const int s = 89 - bitlen128 (q);
	const unsigned __int128 N = q << s;
	// long division by D (34 bits) in 16-bit limbs: the running remainder stays below 2^50
	unsigned long long r = 0;
	unsigned __int128 Q = 0;
	#pragma unroll
	for (int limb = 5; limb >= 0; limb--)
		{
		const unsigned long long t = (r << 16) | (unsigned long long) ((N >> (16 * limb)) & 0xffffu);
		const unsigned long long qd = t / D;
		r = t - qd * D;
		Q = (Q << 16) | qd;
		}
	const int drop = bitlen128 (Q) - 53;                        // 2 or 3
	unsigned long long Qm = (unsigned long long) (Q >> drop);
	const unsigned long long low = (unsigned long long) Q & ((1ull << drop) - 1ull), hq = 1ull << (drop - 1);
	if (low > hq || (low == hq && (r != 0 || (Qm & 1ull)))) Qm += 1;
	const double res = ldexp ((double) Qm, drop - s);           // Qm <= 2^53: exact; the scaling is exact too
	return sign ? -res : res;
	}

__global__ void __launch_bounds__(256)
k_text_roundtrip (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, double* __restrict__ sig)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * 2048;
	uint64_t t1 = t0 + 2048;  if (t1 > sd.hi) t1 = sd.hi;
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256) sig[i] = text_roundtrip10 (sig[i]);
	}

extern "C" int gdsp_text_roundtrip (gdsp_ctx* c, const gdsp_layout* L_, double* sig, int decimals)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig, "gdsp_text_roundtrip: NULL argument");
	GDSP_REQUIRE (decimals == 10, "gdsp_text_roundtrip: only the 10 decimals of write_all_chromosomes are implemented");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, 2048, &tm));
	k_text_roundtrip<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// minover / maxover (minmax.c:322-343, :725-746): for every interval of a sorted, disjoint table the
// cell holding the extremum; among equal values the one farthest from both interval ends
// (inset = min (ix-start, end-ix)), the earliest of those.  One warp per interval; the winning
// buffer cell index is stored (as a double) in the table's value column, which
// GDSP_PW_IVL_KEEP_AT then compares with every cell's own index.
// ---------------------------------------------------------------------------

struct ArgBest { double v;  uint64_t inset, idx;  bool has; };

template <bool WANT_MAX>
__device__ __forceinline__ bool arg_better (const ArgBest& a, const ArgBest& b)      // is a better than b?
	{
	if (!a.has) return false;
	if (!b.has) return true;
	if (WANT_MAX ? (a.v > b.v) : (a.v < b.v)) return true;
	if (WANT_MAX ? (a.v < b.v) : (a.v > b.v)) return false;
	if (a.inset != b.inset) return a.inset > b.inset;
	return a.idx < b.idx;
	}

template <bool WANT_MAX>
__global__ void __launch_bounds__(256)
k_ivl_arg_extrema (const double* __restrict__ sig, const uint64_t* __restrict__ start, const uint64_t* __restrict__ end,
                   double* __restrict__ val, uint64_t n)
	{
	const int lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
	for (uint64_t k = warp; k < n; k += nwarps)
		{
		const uint64_t s = start[k], e = end[k];
		ArgBest best;  best.has = false;  best.v = 0;  best.inset = 0;  best.idx = 0;
		for (uint64_t i = s + lane; i < e; i += 32)
			{
			ArgBest c;  c.has = true;  c.v = sig[i];  c.idx = i;
			c.inset = (i - s < e - i) ? (i - s) : (e - i);
			if (arg_better<WANT_MAX> (c, best)) best = c;
			}
		#pragma unroll
		for (int d = 16; d > 0; d >>= 1)
			{
			ArgBest o;
			o.v = shfl_xor_f64 (best.v, d);
			o.inset = __shfl_xor_sync (0xffffffffu, best.inset, d);
			o.idx   = __shfl_xor_sync (0xffffffffu, best.idx, d);
			o.has   = __shfl_xor_sync (0xffffffffu, (int) best.has, d) != 0;
			if (arg_better<WANT_MAX> (o, best)) best = o;
			}
		if (lane == 0) val[k] = best.has ? (double) best.idx : -1.0;
		}
	}

extern "C" int gdsp_ivl_arg_extrema (gdsp_ctx* c, const gdsp_layout* L, const double* sig, gdsp_ivl_table* t, int wantMax)
	{
	GDSP_REQUIRE (c && L && sig && t, "gdsp_ivl_arg_extrema: NULL argument");
	if (t->n == 0) return GDSP_OK;
	uint64_t blocks = (t->n + 7) / 8;
	if (blocks > (uint64_t) c->sm_count * 32) blocks = (uint64_t) c->sm_count * 32;
	if (wantMax) k_ivl_arg_extrema<true><<<(unsigned) blocks, 256, 0, c->stream>>> (sig, t->d_start, t->d_end, t->d_val, t->n);
	else         k_ivl_arg_extrema<false><<<(unsigned) blocks, 256, 0, c->stream>>> (sig, t->d_start, t->d_end, t->d_val, t->n);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// map (map.c:263-357): piecewise-linear function through breakpoints (in[k], out[k]), in[] strictly
// ascending.  v <= in[0] -> out[0]; v >= in[n-1] -> out[n-1]; a breakpoint maps to its output; anything
// else on piece k to  out[k] + (v - in[k]) * (out[k+1] - out[k]) / (in[k+1] - in[k])  evaluated in
// the reference's order (product first, then the quotient, then the sum; no FMA).
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
k_map_values (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg, double* __restrict__ sig,
              const double* __restrict__ bin, const double* __restrict__ bout, int n)
	{
	extern __shared__ double s_map[];              // [in | out] when they fit
	const double* xin = bin;  const double* xout = bout;
	if (n <= 2048)
		{
		for (int i = threadIdx.x; i < n; i += 256) { s_map[i] = bin[i];  s_map[n + i] = bout[i]; }
		__syncthreads ();
		xin = s_map;  xout = s_map + n;
		}
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	const uint64_t t0 = sd.lo + tis * 4096;
	uint64_t t1 = t0 + 4096;  if (t1 > sd.hi) t1 = sd.hi;
	const double minIn = xin[0], maxIn = xin[n - 1], outMin = xout[0], outMax = xout[n - 1];
	for (uint64_t i = t0 + threadIdx.x; i < t1; i += 256)
		{
		const double x = sig[i];
		double y;
		if (x <= minIn) y = outMin;
		else if (x >= maxIn) y = outMax;
		else if (!(x == x)) y = x;                 // NaN: the reference's search is undefined here
		else
			{
			int lo = 0, hi = n - 1;                // in[lo] <= x < in[hi]
			while (lo + 1 < hi)
				{
				const int mid = (lo + hi) >> 1;
				if (x < xin[mid]) hi = mid; else lo = mid;
				}
			const double pLo = xin[lo], pHi = xin[lo + 1], oLo = xout[lo], oHi = xout[lo + 1];
			if (x == pLo) y = oLo;
			else if (x == pHi) y = oHi;
			else y = __dadd_rn (oLo, __ddiv_rn (__dmul_rn (__dsub_rn (x, pLo), __dsub_rn (oHi, oLo)), __dsub_rn (pHi, pLo)));
			}
		sig[i] = y;
		}
	}

extern "C" int gdsp_map_values (gdsp_ctx* c, const gdsp_layout* L_, double* sig, const double* h_in, const double* h_out, int n)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c && L && sig && h_in && h_out, "gdsp_map_values: NULL argument");
	GDSP_REQUIRE (n >= 1 && n <= (1 << 24), "gdsp_map_values: %d breakpoints (1..16777216 allowed)", n);
	for (int k = 1; k < n; k++)
		GDSP_REQUIRE (h_in[k] > h_in[k-1], "gdsp_map_values: input values must be strictly ascending (entry %d)", k);
	void* ws;
	GDSP_TRY (gdsp_ws (c, 6, 2 * sizeof (double) * (size_t) n, &ws));
	double* d_in = (double*) ws;  double* d_out = d_in + n;
	GDSP_CUDA (cudaMemcpyAsync (d_in, h_in, sizeof (double) * n, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (d_out, h_out, sizeof (double) * n, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));          // h_in/h_out may be host temporaries
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, 4096, &tm));
	const size_t smem = (n <= 2048) ? 2 * sizeof (double) * (size_t) n : 0;
	k_map_values<<<(unsigned) tm.ntiles, 256, smem, c->stream>>> (L->d, tm.d_base, L->nseg, sig, d_in, d_out, n);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
