// cuda_init_time.cu -- where does the CLI's "device open" time go?  Times, in a fresh process:
// cuInit, primary context creation, a first allocation, a first kernel launch (module load), and the
// allocation + zero fill of a cfg1-sized and an hg38/16-sized genome.  Run it with and without
// CUDA_VISIBLE_DEVICES to see what device enumeration costs on a multi-GPU box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -cudart static -o cuda_init_time cuda_init_time.cu -lcuda
#include <cstdio>
#include <chrono>
#include <cuda.h>
#include <cuda_runtime.h>

static double now () { return std::chrono::duration<double> (std::chrono::steady_clock::now ().time_since_epoch ()).count (); }
__global__ void k_touch (double* p) { p[threadIdx.x] = 1.0; }

int main ()
	{
	double t0 = now ();
	cuInit (0);
	double t1 = now ();
	int n = 0;  cudaGetDeviceCount (&n);
	cudaSetDevice (0);
	cudaFree (0);
	double t2 = now ();
	double* p = NULL;
	cudaMalloc (&p, 160u << 20);
	double t3 = now ();
	k_touch<<<1, 32>>> (p);
	cudaDeviceSynchronize ();
	double t4 = now ();
	cudaMemset (p, 0, 160u << 20);
	cudaDeviceSynchronize ();
	double t5 = now ();
	double* q = NULL;
	cudaMalloc (&q, (size_t) 3200 << 20);
	cudaMemset (q, 0, (size_t) 3200 << 20);
	cudaDeviceSynchronize ();
	double t6 = now ();
	cudaStream_t s;  cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking);
	cudaEvent_t e;  cudaEventCreate (&e);
	double t7 = now ();
	void* h = NULL;  cudaMallocHost (&h, 64 << 20);
	double t8 = now ();
	printf ("devices=%d cuInit=%.3f ctx=%.3f malloc160M=%.3f first_kernel=%.3f memset160M=%.3f malloc+memset3.2G=%.3f stream+event=%.3f mallocHost64M=%.3f total=%.3f\n",
	        n, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7, t8 - t0);
	return 0;
	}
