// Microbenchmark: pure-write, pure-read and copy bandwidth of one B200 with the store flavours the
// library uses.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_write_v2 (double* p, size_t n2, double v)
	{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x)
		reinterpret_cast<double2*> (p)[i] = make_double2 (v, v);
	}
__global__ void k_write_v2_na (double* p, size_t n2, double v)
	{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x)
		asm volatile ("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(p + 2 * i), "d"(v), "d"(v) : "memory");
	}
__global__ void k_write_v2_cs (double* p, size_t n2, double v)
	{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x)
		asm volatile ("st.global.cs.v2.f64 [%0], {%1,%2};" :: "l"(p + 2 * i), "d"(v), "d"(v) : "memory");
	}
__global__ void k_write_v4 (double* p, size_t n4, double v)
	{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x)
		asm volatile ("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p + 4 * i), "d"(v), "d"(v), "d"(v), "d"(v) : "memory");
	}
// one block writes one contiguous 64 KB chunk (the k_bin_final / k_scan_tiles pattern)
__global__ void k_write_tile (double* p, size_t ntiles, double v)
	{
	for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x)
		{
		double* q = p + t * 8192;
		for (int i = threadIdx.x; i < 4096; i += blockDim.x)
			asm volatile ("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" :: "l"(q + 2 * i), "d"(v), "d"(v) : "memory");
		}
	}
__global__ void k_read (const double* p, size_t n2, double* out)
	{
	double s = 0;
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x)
		{ double2 v = reinterpret_cast<const double2*> (p)[i];  s += v.x + v.y; }
	if (s == 12345.678) out[0] = s;
	}
__global__ void k_copy (const double* a, double* b, size_t n2)
	{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t) gridDim.x * blockDim.x)
		reinterpret_cast<double2*> (b)[i] = reinterpret_cast<const double2*> (a)[i];
	}

template <typename F> static void timeit (const char* name, double gbytes, F f)
	{
	cudaEvent_t a, b;  cudaEventCreate (&a);  cudaEventCreate (&b);
	float best = 1e30f;
	for (int r = 0; r < 4; r++)
		{
		cudaEventRecord (a);  f ();  cudaEventRecord (b);  cudaEventSynchronize (b);
		float ms;  cudaEventElapsedTime (&ms, a, b);  if (r && ms < best) best = ms;
		}
	printf ("%-34s %8.3f ms  %7.1f GB/s   %s\n", name, best, gbytes / (best * 1e-3), cudaGetErrorString (cudaGetLastError ()));
	}

int main ()
	{
	const size_t n = 3ull << 30;                  // 3 Gi doubles = 25.8 GB
	double *a, *b;
	cudaMalloc (&a, n * 8);  cudaMalloc (&b, n * 8);
	const double gb = n * 8 / 1e9;
	const int sms = 148;
	timeit ("cudaMemsetAsync",              gb, [&] { cudaMemsetAsync (a, 0, n * 8); });
	timeit ("write st.v2.f64 grid-stride",  gb, [&] { k_write_v2<<<sms * 16, 256>>> (a, n / 2, 1.0); });
	timeit ("write st.v2.f64 one-shot",     gb, [&] { k_write_v2<<<(unsigned) (n / 2 / 256), 256>>> (a, n / 2, 1.0); });
	timeit ("write no_allocate v2",         gb, [&] { k_write_v2_na<<<sms * 16, 256>>> (a, n / 2, 1.0); });
	timeit ("write .cs v2",                 gb, [&] { k_write_v2_cs<<<sms * 16, 256>>> (a, n / 2, 1.0); });
	timeit ("write st.v4.f64 (256-bit)",    gb, [&] { k_write_v4<<<sms * 16, 256>>> (a, n / 4, 1.0); });
	timeit ("write 64 KB tile per block",   gb, [&] { k_write_tile<<<(unsigned) (n / 8192), 256>>> (a, n / 8192, 1.0); });
	timeit ("write 64 KB tile, persistent", gb, [&] { k_write_tile<<<sms * 8, 256>>> (a, n / 8192, 1.0); });
	timeit ("read ld.v2.f64",               gb, [&] { k_read<<<sms * 16, 256>>> (a, n / 2, b); });
	timeit ("copy (read+write bytes)",  2 * gb, [&] { k_copy<<<sms * 16, 256>>> (a, b, n / 2); });
	timeit ("cudaMemcpyAsync d2d",      2 * gb, [&] { cudaMemcpyAsync (b, a, n * 8, cudaMemcpyDeviceToDevice); });
	return 0;
	}
