// Development micro-benchmark (not part of the product): issue cost of FP64-pipe instructions on
// sm_100a per SM sub-partition, for 1..8 resident warps per sub-partition and dependent chains of
// varying ILP.  Prints cycles per warp-instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int ILP>
__global__ void bench(double* out, long long* cyc, int iters, double a, double b)
{
	double x[ILP];
	#pragma unroll
	for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
	int p = 0;
	long long t0 = clock64();
	for (int it = 0; it < iters; it++)
	{
		#pragma unroll
		for (int r = 0; r < 8; r++)
		#pragma unroll
		for (int i = 0; i < ILP; i++)
		{
			if (OP == 0) x[i] = fma(x[i], a, b);                 // DFMA, uniform operands
			if (OP == 1) x[i] = fma(x[i], x[i], x[i]);           // DFMA, registers only
			if (OP == 2) x[i] = x[i] * a;                        // DMUL
			if (OP == 3) x[i] = x[i] + b;                        // DADD
			if (OP == 4) { p += (x[i] > a + p) ? 1 : 0; }        // DSETP (+ int)
			if (OP == 5) { float f = (float)x[i]; x[i] = (double)f + 0.0; x[i] = __longlong_as_double(__double_as_longlong(x[i]) + 1); }  // F2F both ways
			if (OP == 6) x[i] = fma(x[i], a, b) , p += __double2hiint(x[i]) >> 20;   // DFMA + 2 int ops
			if (OP == 7) x[i] = fma(x[i], x[(i + 1) % ILP], x[(i + 2) % ILP]);      // DFMA, three distinct register pairs
			if (OP == 8) x[i] = fma(x[i], x[(i + 1) % ILP], b);                      // DFMA, two distinct register pairs + uniform
			if (OP == 9) { bool q = x[i] > x[(i + 1) % ILP]; x[i] = q ? x[i] + b : x[(i + 2) % ILP]; }   // DSETP + DADD + 2 SEL
			if (OP == 10) { x[i] = fma(x[i], x[(i + 1) % ILP], x[(i + 2) % ILP]); p = (p ^ i) + it; }     // DFMA 3 regs + 2 int (independent)
		}
	}
	long long t1 = clock64();
	double s = p;
	#pragma unroll
	for (int i = 0; i < ILP; i++) s += x[i];
	if (s == 123.456) out[0] = s;
	if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP, int ILP>
void run(const char* name, int dp_per_iter_unit)
{
	double* d; long long* c; cudaMalloc(&d, 8); cudaMalloc(&c, 8);
	printf("%-28s ILP=%d :", name, ILP);
	for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2)
	{
		const int threads = 128 * warps_per_smsp > 1024 ? 1024 : 128 * warps_per_smsp;   // one block per SM, 4 SMSPs
		const int iters = 2000;
		bench<OP, ILP><<<148, threads>>>(d, c, iters, 0.999999, 1e-7);
		cudaDeviceSynchronize();
		long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
		const double insts_per_smsp = (double)iters * 8 * ILP * dp_per_iter_unit * (threads / 128);
		printf("  w%d: %.2f cyc/inst", threads / 128, h / insts_per_smsp);
	}
	printf("\n");
	cudaFree(d); cudaFree(c);
}

int main()
{
	run<0, 1>("DFMA uniform ops", 1); run<0, 2>("DFMA uniform ops", 1); run<0, 4>("DFMA uniform ops", 1); run<0, 8>("DFMA uniform ops", 1);
	run<1, 1>("DFMA regs", 1); run<1, 4>("DFMA regs", 1); run<1, 8>("DFMA regs", 1);
	run<2, 4>("DMUL", 1); run<2, 8>("DMUL", 1);
	run<3, 4>("DADD", 1); run<3, 8>("DADD", 1);
	run<4, 4>("DSETP+int", 1); run<4, 8>("DSETP+int", 1);
	run<5, 4>("F2F x2 (+DADD)", 3); run<5, 8>("F2F x2 (+DADD)", 3);
	run<6, 4>("DFMA + 2 int", 1); run<6, 8>("DFMA + 2 int", 1);
	run<7, 4>("DFMA 3 distinct regs", 1); run<7, 8>("DFMA 3 distinct regs", 1);
	run<8, 4>("DFMA 2 regs + uniform", 1); run<8, 8>("DFMA 2 regs + uniform", 1);
	run<9, 4>("DSETP+DADD+2SEL (2 DP)", 2); run<9, 8>("DSETP+DADD+2SEL (2 DP)", 2);
	run<10, 4>("DFMA 3 regs + 2 int indep", 1); run<10, 8>("DFMA 3 regs + 2 int indep", 1);
	return 0;
}
