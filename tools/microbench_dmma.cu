// Development micro-benchmark (not part of the product): issue cost of the FP64 tensor-core instruction
// mma.sync.aligned.m8n8k4.row.col.f64 on sm_100a against plain DFMA, per SM sub-partition.  One DMMA = 8x8x4 = 256 FMAs
// per warp = 8 DFMA warp-instructions' worth of work; decides whether the dense coupling mat-vec should use it.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void bench_dmma(double* out, long long* cyc, int iters)
{
	double c0[ILP], c1[ILP];
	double a = threadIdx.x * 1e-3, b = 1.0 - threadIdx.x * 1e-4;
	#pragma unroll
	for (int i = 0; i < ILP; i++) { c0[i] = i; c1[i] = -i; }
	long long t0 = clock64();
	for (int it = 0; it < iters; it++)
	{
		#pragma unroll
		for (int r = 0; r < 8; r++)
		#pragma unroll
		for (int i = 0; i < ILP; i++)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
	}
	long long t1 = clock64();
	double s = 0;
	#pragma unroll
	for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
	if (s == 123.456) out[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run()
{
	double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
	printf("DMMA m8n8k4 ILP=%d :", ILP);
	for (int w = 1; w <= 8; w *= 2)
	{
		const int threads = 128 * w;
		const int iters = 2000;
		cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
		bench_dmma<ILP><<<148, threads>>>(d, c, iters);   // warm
		cudaEventRecord(e0);
		bench_dmma<ILP><<<148, threads>>>(d, c, iters);
		cudaEventRecord(e1);
		cudaDeviceSynchronize();
		float ms; cudaEventElapsedTime(&ms, e0, e1);
		long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		const double per_smsp = (double)iters * 8 * ILP * w;
		const double tflops = 148.0 * 4 * per_smsp * 512 / (ms * 1e-3) / 1e12;
		printf("  w%d: %.2f cyc/DMMA (%.1f TFLOP/s)", w, h / per_smsp, tflops);
	}
	printf("\n");
	cudaFree(d); cudaFree(c);
}

int main()
{
	run<1>(); run<2>(); run<4>(); run<8>();
	return 0;
}
