// gdsp_format.cu -- text formatting of run-length output on the device.
//
// Replaces the fprintf loop of report_intervals (genodsp.c:1606-1678): every run becomes the line
//     <chrom> TAB <start> TAB <end> [TAB <value "%.*f">] NEWLINE
// Formatting 10^8 lines with printf is the slowest part of a whole run of the reference (and of a
// host formatter fed by the GPU); here one thread formats one line:
//   k_fmt_measure   length of every line (u16) and the total per block of 256 lines
//   k_fmt_offsets   exclusive prefix of the block totals (one block)
//   k_fmt_write     block-local prefix of the lengths, characters written to their final position
// "%.*f" is reproduced exactly: with |v| = m * 2^e, q = round-half-even (|v| * 10^p) computed in
// 128-bit integers is the digit string glibc prints (it rounds the exact binary value), the sign is
// the sign bit (so -0.0 and negative values that round to zero print "-0.000").  Values that do
// not fit (NaN, infinities, |v| >= 2^63, precision > 17) raise a flag and the caller formats that
// chunk on the host instead.
#include "gdsp_common.cuh"

#define FMT_THREADS 256
#define FMT_MAXVAL  64          // longest value string handled here

__constant__ unsigned long long c_pow10[18] = {
	1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
	10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull,
	1000000000000000ull, 10000000000000000ull, 100000000000000000ull };

// decimal digits of x, written right-aligned ending at buf[end-1]; returns the first index used
__device__ __forceinline__ int fmt_u64_right (unsigned long long x, char* buf, int end)
	{
	int i = end;
	do { const unsigned long long y = x / 10ull;  buf[--i] = (char) ('0' + (int) (x - y * 10ull));  x = y; } while (x);
	return i;
	}

// "%.*f" of v into out[] (at most FMT_MAXVAL chars); returns the length, or -1 if not handled here
__device__ int fmt_value (double v, int p, char* out)
	{
	const unsigned long long b = (unsigned long long) __double_as_longlong (v);
	const bool neg = (b >> 63) != 0;
	const unsigned long long ab = b & 0x7fffffffffffffffull;
	const int ex = (int) (ab >> 52);
	const unsigned long long frac = ab & 0x000fffffffffffffull;
	if (ex == 0x7ff) return -1;
	unsigned __int128 q = 0;
	if (ab != 0ull)
		{
		unsigned long long m;  int e;
		if (ex == 0) { m = frac;  e = -1074; } else { m = frac | 0x0010000000000000ull;  e = ex - 1075; }
		if (e >= 0)
			{
			if (e > 10) return -1;                                    // |v| >= 2^63
			q = ((unsigned __int128) (m << e)) * c_pow10[p];
			}
		else
			{
			const int k = -e;
			if (k < 124)
				{
				const unsigned __int128 P = (unsigned __int128) m * c_pow10[p];     // < 2^110
				q = P >> k;
				const unsigned __int128 rem = P & ((((unsigned __int128) 1) << k) - 1), half = ((unsigned __int128) 1) << (k - 1);
				if (rem > half || (rem == half && (q & 1))) q += 1;
				}
			}
		}
	// digits of q, zero-padded to at least p+1
	char dig[48];
	int first = 48;
	if ((q >> 64) == 0) first = fmt_u64_right ((unsigned long long) q, dig, 48);
	else
		{
		// split off 18 digits at a time
		const unsigned long long D18 = 1000000000000000000ull;
		unsigned __int128 hi = q / D18;
		unsigned long long lo = (unsigned long long) (q - hi * D18);
		int i = fmt_u64_right (lo, dig, 48);
		while (i > 48 - 18) dig[--i] = '0';
		if ((hi >> 64) == 0) first = fmt_u64_right ((unsigned long long) hi, dig, i);
		else
			{
			unsigned __int128 hi2 = hi / D18;
			unsigned long long mid = (unsigned long long) (hi - hi2 * D18);
			i = fmt_u64_right (mid, dig, i);
			while (i > 48 - 36) dig[--i] = '0';
			first = fmt_u64_right ((unsigned long long) hi2, dig, i);
			}
		}
	while (48 - first < p + 1) dig[--first] = '0';
	int n = 0;
	if (neg) out[n++] = '-';
	const int nd = 48 - first;
	for (int i = 0; i < nd - p; i++) out[n++] = dig[first + i];
	if (p > 0)
		{
		out[n++] = '.';
		for (int i = nd - p; i < nd; i++) out[n++] = dig[first + i];
		}
	return n;
	}

__device__ __forceinline__ int fmt_u32_len (uint32_t x)
	{
	int n = 1;
	while (x >= 10u) { x /= 10u;  n++; }
	return n;
	}

struct FmtArgs
	{
	const uint32_t* start;  const uint32_t* end;  const double* val;
	uint64_t n;
	uint32_t addStart, addEnd;
	int      withValue, precision, nameLen;
	char     name[256];
	};

__global__ void __launch_bounds__(FMT_THREADS)
k_fmt_measure (const __grid_constant__ FmtArgs A, unsigned short* __restrict__ lens,
               unsigned long long* __restrict__ blockSum, int* __restrict__ unsupported)
	{
	__shared__ unsigned int s_w[FMT_THREADS / 32];
	const uint64_t r = (uint64_t) blockIdx.x * FMT_THREADS + threadIdx.x;
	unsigned int len = 0;
	if (r < A.n)
		{
		len = (unsigned) A.nameLen + 1u + (unsigned) fmt_u32_len (A.start[r] + A.addStart) + 1u + (unsigned) fmt_u32_len (A.end[r] + A.addEnd) + 1u;
		if (A.withValue)
			{
			char tmp[FMT_MAXVAL];
			const int vl = fmt_value (A.val[r], A.precision, tmp);
			if (vl < 0) { atomicExch (unsupported, 1);  len = 0; }
			else len += 1u + (unsigned) vl;
			}
		lens[r] = (unsigned short) len;
		}
	unsigned int t = len;
	#pragma unroll
	for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync (0xffffffffu, t, d);
	if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
	__syncthreads ();
	if (threadIdx.x == 0)
		{
		unsigned long long tot = 0;
		for (int w = 0; w < FMT_THREADS / 32; w++) tot += s_w[w];
		blockSum[blockIdx.x] = tot;
		}
	}

// one block: in-place exclusive prefix of the block totals; total[0] = grand total
__global__ void __launch_bounds__(1024)
k_fmt_offsets (unsigned long long* __restrict__ blockSum, uint64_t nblocks, unsigned long long* __restrict__ total)
	{
	__shared__ unsigned long long s_w[32];
	__shared__ unsigned long long s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads ();
	for (uint64_t c0 = 0; c0 < nblocks; c0 += 1024)
		{
		const uint64_t i = c0 + threadIdx.x;
		const unsigned long long v = (i < nblocks) ? blockSum[i] : 0ull;
		unsigned long long inc = v;
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
		if (i < nblocks) blockSum[i] = carry + wex + inc - v;
		__syncthreads ();
		if (threadIdx.x == 0) s_carry = carry + tot;
		__syncthreads ();
		}
	if (threadIdx.x == 0) total[0] = s_carry;
	}

__global__ void __launch_bounds__(FMT_THREADS)
k_fmt_write (const __grid_constant__ FmtArgs A, const unsigned short* __restrict__ lens,
             const unsigned long long* __restrict__ blockOff, char* __restrict__ text, unsigned long long cap)
	{
	__shared__ unsigned int s_w[FMT_THREADS / 32];
	const uint64_t r = (uint64_t) blockIdx.x * FMT_THREADS + threadIdx.x;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned int len = (r < A.n) ? lens[r] : 0u;
	unsigned int inc = len;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		{
		unsigned int up = __shfl_up_sync (0xffffffffu, inc, d);
		if (lane >= d) inc += up;
		}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads ();
	unsigned int wex = 0;
	for (int w = 0; w < warp; w++) wex += s_w[w];
	const unsigned long long off = blockOff[blockIdx.x] + wex + inc - len;
	if (r >= A.n || len == 0 || off + len > cap) return;

	char* o = text + off;
	int n = 0;
	for (int i = 0; i < A.nameLen; i++) o[n++] = A.name[i];
	o[n++] = '\t';
	{
	char d[12];
	int f = fmt_u64_right ((unsigned long long) (A.start[r] + A.addStart), d, 12);
	for (int i = f; i < 12; i++) o[n++] = d[i];
	o[n++] = '\t';
	f = fmt_u64_right ((unsigned long long) (A.end[r] + A.addEnd), d, 12);
	for (int i = f; i < 12; i++) o[n++] = d[i];
	}
	if (A.withValue)
		{
		o[n++] = '\t';
		char tmp[FMT_MAXVAL];
		const int vl = fmt_value (A.val[r], A.precision, tmp);
		for (int i = 0; i < vl; i++) o[n++] = tmp[i];
		}
	o[n++] = '\n';
	}

extern "C" size_t gdsp_format_runs_max_bytes (uint64_t n, const char* chrom)
	{
	return (size_t) n * (strlen (chrom) + 1 + 10 + 1 + 10 + 1 + FMT_MAXVAL + 1);
	}

extern "C" int gdsp_format_runs (gdsp_ctx* c, const uint32_t* d_start, const uint32_t* d_end, const double* d_val,
                                 uint64_t n, const char* chrom, uint32_t add_start, uint32_t add_end,
                                 int with_value, int precision, char* d_text, uint64_t cap,
                                 uint64_t* h_bytes, int* h_unsupported)
	{
	GDSP_REQUIRE (c && chrom && h_bytes && h_unsupported, "gdsp_format_runs: NULL argument");
	*h_bytes = 0;  *h_unsupported = 0;
	if (n == 0) return GDSP_OK;
	GDSP_REQUIRE (d_start && d_end && d_text && (d_val || !with_value), "gdsp_format_runs: NULL array");
	const size_t nameLen = strlen (chrom);
	if (nameLen > 255 || precision < 0 || precision > 17) { *h_unsupported = 1;  return GDSP_OK; }
	FmtArgs A;
	memset (&A, 0, sizeof (A));
	A.start = d_start;  A.end = d_end;  A.val = d_val;  A.n = n;
	A.addStart = add_start;  A.addEnd = add_end;  A.withValue = with_value ? 1 : 0;  A.precision = precision;
	A.nameLen = (int) nameLen;
	memcpy (A.name, chrom, nameLen);
	const uint64_t nblocks = (n + FMT_THREADS - 1) / FMT_THREADS;
	void* ws;
	const size_t lensBytes = ((n * sizeof (unsigned short) + 255) / 256) * 256;
	GDSP_TRY (gdsp_ws (c, 7, 256 + lensBytes + (nblocks + 1) * sizeof (unsigned long long), &ws));
	int* d_flag = (int*) ws;
	unsigned long long* d_total = (unsigned long long*) ((char*) ws + 64);
	unsigned short* d_lens = (unsigned short*) ((char*) ws + 256);
	unsigned long long* d_blk = (unsigned long long*) ((char*) ws + 256 + lensBytes);
	GDSP_CUDA (cudaMemsetAsync (ws, 0, 256, c->stream));
	k_fmt_measure<<<(unsigned) nblocks, FMT_THREADS, 0, c->stream>>> (A, d_lens, d_blk, d_flag);
	GDSP_KERNEL_CHECK ();
	k_fmt_offsets<<<1, 1024, 0, c->stream>>> (d_blk, nblocks, d_total);
	GDSP_KERNEL_CHECK ();
	int flag = 0;  unsigned long long total = 0;
	GDSP_CUDA (cudaMemcpyAsync (&flag, d_flag, sizeof (int), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaMemcpyAsync (&total, d_total, sizeof (total), cudaMemcpyDeviceToHost, c->stream));
	GDSP_CUDA (cudaStreamSynchronize (c->stream));
	if (flag) { *h_unsupported = 1;  return GDSP_OK; }
	if (total > cap) { *h_bytes = total;  gdsp_set_error ("gdsp_format_runs: %llu bytes of text, capacity %llu", total, (unsigned long long) cap);  return GDSP_ERR_CAPACITY; }
	k_fmt_write<<<(unsigned) nblocks, FMT_THREADS, 0, c->stream>>> (A, d_lens, d_blk, d_text, cap);
	GDSP_KERNEL_CHECK ();
	*h_bytes = total;
	return GDSP_OK;
	}
