// Microbenchmark: sustained FP64 instruction rate of one B200 for the instruction mix
// smooth() needs (separate DMUL + DADD, no FMA, for bit-exactness) under different ways
// of feeding the tap value.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_peak fp64_peak.cu
// MODE 0: DFMA, tap from shared memory     (1 LDS per R FMAs)
// MODE 1: DMUL+DADD, tap from shared memory (1 LDS per 2R instructions)
// MODE 2: DMUL+DADD, tap from a __grid_constant__ kernel parameter (constant bank)
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { double w[1024]; };

template <int MODE, int R>
__global__ void __launch_bounds__(128) k (double* out, const double* __restrict__ taps, const __grid_constant__ Taps tp, int iters)
	{
	__shared__ double s_w[1024];
	for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_w[i] = taps[i];
	__syncthreads ();
	double acc[R], x[R];
	for (int r = 0; r < R; r++) { acc[r] = threadIdx.x * 1e-9 + r;  x[r] = 1.0 + r * 1e-3 + threadIdx.x * 1e-6; }
	for (int it = 0; it < iters; it++)
		{
		#pragma unroll
		for (int u = 0; u < 8; u++)
			{
			const double w = (MODE == 2) ? tp.w[(it * 8 + u) & 1023] : s_w[(it * 8 + u) & 1023];
			#pragma unroll
			for (int r = 0; r < R; r++)
				{
				if (MODE == 0) acc[r] = __fma_rn (w, x[r], acc[r]);
				else           acc[r] = __dadd_rn (acc[r], __dmul_rn (w, x[r]));
				}
			}
		}
	double s = 0;
	for (int r = 0; r < R; r++) s += acc[r];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	}

template <int MODE, int R>
static void run (const char* name, int sms, double* out, double* taps, const Taps& tp)
	{
	cudaEvent_t a, b;  cudaEventCreate (&a);  cudaEventCreate (&b);
	int blocks = sms * 16, iters = 4000;
	float best = 1e30f;
	for (int rep = 0; rep < 3; rep++)
		{
		cudaEventRecord (a);
		k<MODE, R><<<blocks, 128>>> (out, taps, tp, iters);
		cudaEventRecord (b);  cudaEventSynchronize (b);
		float ms;  cudaEventElapsedTime (&ms, a, b);  if (ms < best) best = ms;
		}
	int perSM = 0;  cudaOccupancyMaxActiveBlocksPerMultiprocessor (&perSM, k<MODE, R>, 128, 0);
	double instr = (double) blocks * 128 * iters * 8 * R * (MODE == 0 ? 1 : 2);
	printf ("%-28s R=%2d  %7.3f ms  %.3e FP64 thread-instr/s  (%d CTAs/SM)  err=%s\n", name, R, best, instr / (best * 1e-3), perSM,
	        cudaGetErrorString (cudaGetLastError ()));
	}

int main ()
	{
	cudaDeviceProp p;  cudaGetDeviceProperties (&p, 0);
	double* out;  cudaMalloc (&out, sizeof (double) * p.multiProcessorCount * 16 * 128);
	double* taps;  cudaMalloc (&taps, 8192);
	static Taps tp;  for (int i = 0; i < 1024; i++) tp.w[i] = 1.0 / (i + 3);
	cudaMemcpy (taps, tp.w, 8192, cudaMemcpyHostToDevice);
	printf ("%d SMs, nominal %.3e (64 lanes/SM at %.0f MHz)\n", p.multiProcessorCount, p.multiProcessorCount * 64.0 * p.clockRate * 1e3, p.clockRate / 1e3);
	run<0, 8>  ("DFMA      smem tap",  p.multiProcessorCount, out, taps, tp);
	run<1, 8>  ("DMUL+DADD smem tap",  p.multiProcessorCount, out, taps, tp);
	run<1, 12> ("DMUL+DADD smem tap",  p.multiProcessorCount, out, taps, tp);
	run<1, 16> ("DMUL+DADD smem tap",  p.multiProcessorCount, out, taps, tp);
	run<2, 8>  ("DMUL+DADD param tap", p.multiProcessorCount, out, taps, tp);
	run<2, 16> ("DMUL+DADD param tap", p.multiProcessorCount, out, taps, tp);
	return 0;
	}
