// gdsp_ctx.cu -- context, memory, layout and fill entry points of the C-ABI.
#include <stdarg.h>
#include <stdlib.h>
#include "gdsp_common.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------

static thread_local char g_err[1024] = "";

void gdsp_set_error (const char* fmt, ...)
	{
	va_list ap;
	va_start (ap, fmt);
	vsnprintf (g_err, sizeof (g_err), fmt, ap);
	va_end (ap);
	}

unsigned long long g_gdsp_launches = 0;

extern "C" uint64_t gdsp_launch_count (void)
	{ return __atomic_load_n (&g_gdsp_launches, __ATOMIC_RELAXED); }

extern "C" const char* gdsp_last_error (void) { return g_err; }
extern "C" const char* gdsp_version (void) { return "genodsp-b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------

extern "C" int gdsp_ctx_create (int device, void* stream, gdsp_ctx** out)
	{
	GDSP_REQUIRE (out != NULL, "gdsp_ctx_create: out is NULL");
	int ndev = 0;
	if (cudaGetDeviceCount (&ndev) != cudaSuccess || ndev == 0)
		{
		gdsp_set_error ("gdsp_ctx_create: no CUDA device is visible; this library has no CPU path");
		return GDSP_ERR_NODEVICE;
		}
	GDSP_REQUIRE (device >= 0 && device < ndev, "gdsp_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
	GDSP_CUDA (cudaSetDevice (device));
	cudaDeviceProp prop;
	GDSP_CUDA (cudaGetDeviceProperties (&prop, device));
	if (prop.major != 10)
		{
		gdsp_set_error ("gdsp_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
		                device, prop.major, prop.minor);
		return GDSP_ERR_NODEVICE;
		}
	gdsp_ctx* c = (gdsp_ctx*) calloc (1, sizeof (gdsp_ctx));
	if (c == NULL) { gdsp_set_error ("gdsp_ctx_create: out of host memory"); return GDSP_ERR_NOMEM; }
	c->device     = device;
	c->sm_count   = prop.multiProcessorCount;
	c->cc_major   = prop.major;
	c->cc_minor   = prop.minor;
	c->smem_optin = prop.sharedMemPerBlockOptin;
	if (stream != GDSP_STREAM_PRIVATE) { c->stream = (cudaStream_t) stream;  c->owns_stream = false; }
	else
		{
		GDSP_CUDA (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
		c->owns_stream = true;
		}
	GDSP_CUDA (cudaEventCreate (&c->t0));
	GDSP_CUDA (cudaEventCreate (&c->t1));
	GDSP_CUDA (cudaEventCreateWithFlags (&c->pinned_ev[0], cudaEventDisableTiming));
	GDSP_CUDA (cudaEventCreateWithFlags (&c->pinned_ev[1], cudaEventDisableTiming));
	*out = c;
	return GDSP_OK;
	}

extern "C" void gdsp_ctx_destroy (gdsp_ctx* c)
	{
	if (c == NULL) return;
	cudaSetDevice (c->device);
	cudaStreamSynchronize (c->stream);
	for (int i = 0; i < GDSP_NUM_WS; i++) if (c->ws[i] != NULL) cudaFree (c->ws[i]);
	for (int i = 0; i < 2; i++) if (c->pinned[i] != NULL) cudaFreeHost (c->pinned[i]);
	if (c->taps_dev != NULL) cudaFree (c->taps_dev);
	if (c->taps_host != NULL) free (c->taps_host);
	if (c->host_small != NULL) cudaFreeHost (c->host_small);
	cudaEventDestroy (c->t0);  cudaEventDestroy (c->t1);
	cudaEventDestroy (c->pinned_ev[0]);  cudaEventDestroy (c->pinned_ev[1]);
	if (c->owns_stream) cudaStreamDestroy (c->stream);
	free (c);
	}

extern "C" int gdsp_ctx_set_stream (gdsp_ctx* c, void* stream)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_ctx_set_stream: ctx is NULL");
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (c->owns_stream) { cudaStreamDestroy (c->stream);  c->owns_stream = false; }
	if (stream != GDSP_STREAM_PRIVATE) c->stream = (cudaStream_t) stream;
	else
		{
		GDSP_CUDA (cudaStreamCreateWithFlags (&c->stream, cudaStreamNonBlocking));
		c->owns_stream = true;
		}
	return GDSP_OK;
	}

extern "C" int gdsp_sync (gdsp_ctx* c)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_sync: ctx is NULL");
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

extern "C" int gdsp_device_info (gdsp_ctx* c, int* sm, int* maj, int* min, size_t* fre, size_t* tot)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_device_info: ctx is NULL");
	if (sm)  *sm  = c->sm_count;
	if (maj) *maj = c->cc_major;
	if (min) *min = c->cc_minor;
	if (fre || tot)
		{
		size_t f, t;
		GDSP_CUDA (cudaSetDevice (c->device));
		GDSP_CUDA (cudaMemGetInfo (&f, &t));
		if (fre) *fre = f;
		if (tot) *tot = t;
		}
	return GDSP_OK;
	}

// page-locked host memory for callers that stream results out (text output of the CLI)
extern "C" int gdsp_malloc_host (size_t bytes, void** out)
	{
	GDSP_REQUIRE (out != NULL, "gdsp_malloc_host: out is NULL");
	GDSP_CUDA (cudaMallocHost (out, bytes ? bytes : 1));
	return GDSP_OK;
	}

extern "C" int gdsp_free_host (void* p)
	{
	if (p != NULL) GDSP_CUDA (cudaFreeHost (p));
	return GDSP_OK;
	}

int gdsp_host_scratch (gdsp_ctx* c, size_t bytes, void** out)
	{
	if (c->host_small_bytes < bytes)
		{
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		if (c->host_small) { cudaFreeHost (c->host_small);  c->host_small = NULL;  c->host_small_bytes = 0; }
		const size_t want = (bytes + ((size_t) 1 << 20) - 1) & ~(((size_t) 1 << 20) - 1);
		GDSP_CUDA (cudaMallocHost (&c->host_small, want));
		c->host_small_bytes = want;
		}
	*out = c->host_small;
	return GDSP_OK;
	}

int gdsp_ws (gdsp_ctx* c, int slot, size_t bytes, void** out)
	{
	GDSP_REQUIRE (slot >= 0 && slot < GDSP_NUM_WS, "gdsp_ws: bad slot %d", slot);
	if (bytes > c->ws_bytes[slot])
		{
		GDSP_CUDA (cudaStreamSynchronize (c->stream));
		if (c->ws[slot] != NULL) { GDSP_CUDA (cudaFree (c->ws[slot]));  c->ws[slot] = NULL;  c->ws_bytes[slot] = 0; }
		size_t want = bytes + bytes / 8 + 4096;
		cudaError_t e = cudaMalloc (&c->ws[slot], want);
		if (e != cudaSuccess)
			{
			gdsp_set_error ("workspace allocation of %zu bytes failed: %s", want, cudaGetErrorString (e));
			cudaGetLastError ();
			return GDSP_ERR_NOMEM;
			}
		c->ws_bytes[slot] = want;
		}
	*out = c->ws[slot];
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------

extern "C" int gdsp_malloc (gdsp_ctx* c, size_t bytes, void** dptr)
	{
	GDSP_REQUIRE (c != NULL && dptr != NULL, "gdsp_malloc: NULL argument");
	GDSP_CUDA (cudaSetDevice (c->device));
	cudaError_t e = cudaMalloc (dptr, bytes ? bytes : 16);
	if (e != cudaSuccess)
		{
		gdsp_set_error ("gdsp_malloc: %zu bytes: %s", bytes, cudaGetErrorString (e));
		cudaGetLastError ();
		return GDSP_ERR_NOMEM;
		}
	return GDSP_OK;
	}

extern "C" int gdsp_free (gdsp_ctx* c, void* dptr)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_free: ctx is NULL");
	if (dptr == NULL) return GDSP_OK;
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	GDSP_CUDA (cudaFree (dptr));
	return GDSP_OK;
	}

extern "C" int gdsp_host_alloc (gdsp_ctx* c, size_t bytes, void** hptr)
	{
	GDSP_REQUIRE (c != NULL && hptr != NULL, "gdsp_host_alloc: NULL argument");
	cudaError_t e = cudaMallocHost (hptr, bytes ? bytes : 16);
	if (e != cudaSuccess)
		{
		gdsp_set_error ("gdsp_host_alloc: %zu bytes: %s", bytes, cudaGetErrorString (e));
		cudaGetLastError ();
		return GDSP_ERR_NOMEM;
		}
	return GDSP_OK;
	}

extern "C" int gdsp_host_free (gdsp_ctx* c, void* hptr)
	{
	(void) c;
	if (hptr != NULL) GDSP_CUDA (cudaFreeHost (hptr));
	return GDSP_OK;
	}

extern "C" int gdsp_h2d (gdsp_ctx* c, void* dst, const void* src, size_t bytes)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_h2d: ctx is NULL");
	GDSP_CUDA (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

extern "C" int gdsp_d2h (gdsp_ctx* c, void* dst, const void* src, size_t bytes)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_d2h: ctx is NULL");
	GDSP_CUDA (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	return GDSP_OK;
	}

extern "C" int gdsp_d2d (gdsp_ctx* c, void* dst, const void* src, size_t bytes)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_d2d: ctx is NULL");
	GDSP_CUDA (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
	return GDSP_OK;
	}

extern "C" int gdsp_timer_start (gdsp_ctx* c)
	{
	GDSP_REQUIRE (c != NULL, "gdsp_timer_start: ctx is NULL");
	GDSP_CUDA (cudaEventRecord (c->t0, c->stream));
	return GDSP_OK;
	}

extern "C" int gdsp_timer_stop (gdsp_ctx* c, float* ms)
	{
	GDSP_REQUIRE (c != NULL && ms != NULL, "gdsp_timer_stop: NULL argument");
	GDSP_CUDA (cudaEventRecord (c->t1, c->stream));
	GDSP_CUDA (cudaEventSynchronize (c->t1));
	GDSP_CUDA (cudaEventElapsedTime (ms, c->t0, c->t1));
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// layout
// ---------------------------------------------------------------------------

extern "C" int gdsp_layout_pack (const uint32_t* chrom_len, int nseg, gdsp_seg* segs_out, uint64_t* total_cells)
	{
	GDSP_REQUIRE (chrom_len != NULL && nseg > 0, "gdsp_layout_pack: need at least one chromosome");
	uint64_t pos = 0;
	for (int s = 0; s < nseg; s++)
		{
		GDSP_REQUIRE (chrom_len[s] > 0, "gdsp_layout_pack: chromosome %d has length 0", s);
		if (segs_out != NULL)
			{
			segs_out[s].lo  = segs_out[s].dlo = pos;
			segs_out[s].hi  = segs_out[s].dhi = pos + chrom_len[s];
			segs_out[s].pos0 = 0;
			segs_out[s].chrom_len = chrom_len[s];
			}
		pos += chrom_len[s];
		pos = (pos + GDSP_ALIGN - 1) / GDSP_ALIGN * GDSP_ALIGN;
		}
	if (total_cells != NULL) *total_cells = pos + GDSP_ALIGN;   // tail pad for vector over-reads
	return GDSP_OK;
	}

extern "C" int gdsp_layout_create (gdsp_ctx* c, const gdsp_seg* segs, int nseg, gdsp_layout** out)
	{
	GDSP_REQUIRE (c != NULL && segs != NULL && out != NULL && nseg > 0, "gdsp_layout_create: bad argument");
	gdsp_layout* L = new gdsp_layout ();
	L->ctx = c;  L->nseg = nseg;  L->d = NULL;  L->cells = 0;  L->max_len = 0;
	L->span_lo = ~0ull;  L->span_hi = 0;
	std::vector<SegDev> tmp (nseg);
	for (int s = 0; s < nseg; s++)
		{
		const gdsp_seg& g = segs[s];
		if (!(g.lo < g.hi && g.dlo <= g.lo && g.hi <= g.dhi) || (g.lo % GDSP_ALIGN) != 0
		 || (g.hi - g.lo) > 0xffffffffull || (s > 0 && g.dlo < segs[s-1].dhi))
			{
			delete L;
			gdsp_set_error ("gdsp_layout_create: segment %d is malformed (lo=%llu hi=%llu dlo=%llu dhi=%llu)",
			                s, (unsigned long long) g.lo, (unsigned long long) g.hi,
			                (unsigned long long) g.dlo, (unsigned long long) g.dhi);
			return GDSP_ERR_ARG;
			}
		L->h.push_back (g);
		tmp[s].lo = g.lo;  tmp[s].hi = g.hi;  tmp[s].dlo = g.dlo;  tmp[s].dhi = g.dhi;
		tmp[s].pos0 = g.pos0;  tmp[s].chromLen = g.chrom_len;
		L->cells += g.hi - g.lo;
		if (g.hi - g.lo > L->max_len) L->max_len = (uint32_t) (g.hi - g.lo);
		if (g.lo < L->span_lo) L->span_lo = g.lo;
		if (g.hi > L->span_hi) L->span_hi = g.hi;
		}
	cudaSetDevice (c->device);
	cudaError_t e = cudaMalloc (&L->d, sizeof (SegDev) * nseg);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync (L->d, tmp.data (), sizeof (SegDev) * nseg, cudaMemcpyHostToDevice, c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize (c->stream);
	if (e != cudaSuccess)
		{
		gdsp_set_error ("gdsp_layout_create: %s", cudaGetErrorString (e));
		if (L->d) cudaFree (L->d);
		delete L;
		return GDSP_ERR_CUDA;
		}
	*out = L;
	return GDSP_OK;
	}

extern "C" void gdsp_layout_destroy (gdsp_layout* L)
	{
	if (L == NULL) return;
	cudaStreamSynchronize (L->ctx->stream);
	for (auto& kv : L->tiles) cudaFree (kv.second.d_base);
	if (L->d) cudaFree (L->d);
	delete L;
	}

extern "C" int gdsp_layout_nseg (const gdsp_layout* L) { return L ? L->nseg : 0; }
extern "C" const gdsp_seg* gdsp_layout_segs (const gdsp_layout* L) { return L ? L->h.data () : NULL; }
extern "C" uint64_t gdsp_layout_cells (const gdsp_layout* L) { return L ? L->cells : 0; }

int gdsp_layout_tilemap (gdsp_layout* L, uint32_t tile, TileMap* out)
	{
	auto it = L->tiles.find (tile);
	if (it != L->tiles.end ()) { *out = it->second;  return GDSP_OK; }
	std::vector<uint64_t> base (L->nseg + 1);
	uint64_t n = 0;
	for (int s = 0; s < L->nseg; s++)
		{
		base[s] = n;
		n += (L->h[s].hi - L->h[s].lo + tile - 1) / tile;
		}
	base[L->nseg] = n;
	TileMap tm;  tm.tile = tile;  tm.ntiles = n;  tm.d_base = NULL;
	GDSP_CUDA (cudaMalloc (&tm.d_base, sizeof (uint64_t) * (L->nseg + 1)));
	GDSP_CUDA (cudaMemcpyAsync (tm.d_base, base.data (), sizeof (uint64_t) * (L->nseg + 1),
	                            cudaMemcpyHostToDevice, L->ctx->stream));
	GDSP_CUDA (cudaStreamSynchronize (L->ctx->stream));
	L->tiles[tile] = tm;
	*out = tm;
	return GDSP_OK;
	}

// ---------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------

#define FILL_TILE 8192

__global__ void __launch_bounds__(256)
k_fill (const SegDev* __restrict__ segs, const uint64_t* __restrict__ base, int nseg,
        double* __restrict__ sig, double value)
	{
	int seg;  uint64_t tis;
	tile_to_seg (base, nseg, blockIdx.x, seg, tis);
	const SegDev sd = segs[seg];
	uint64_t t0 = sd.lo + tis * FILL_TILE;
	uint64_t t1 = t0 + FILL_TILE;  if (t1 > sd.hi) t1 = sd.hi;
	// lo is 64-aligned, so t0 is even; pairs entirely inside [t0,t1) as 128-bit stores
	double2 vv = make_double2 (value, value);
	for (uint64_t i = t0 + 2 * threadIdx.x; i < t1; i += 2 * blockDim.x)
		{
		if (i + 1 < t1) stg_stream (sig + i, vv);
		else            sig[i] = value;
		}
	}

extern "C" int gdsp_fill (gdsp_ctx* c, const gdsp_layout* L_, double* sig, double value)
	{
	gdsp_layout* L = (gdsp_layout*) L_;
	GDSP_REQUIRE (c != NULL && L != NULL && sig != NULL, "gdsp_fill: NULL argument");
	TileMap tm;
	GDSP_TRY (gdsp_layout_tilemap (L, FILL_TILE, &tm));
	k_fill<<<(unsigned) tm.ntiles, 256, 0, c->stream>>> (L->d, tm.d_base, L->nseg, sig, value);
	GDSP_KERNEL_CHECK ();
	return GDSP_OK;
	}
