// TEST INFRASTRUCTURE ONLY -- host stand-ins that let g++ compile the reference's coupled (block-per-system)
// kernel source and run it with ONE std::thread PER CUDA THREAD, blocks one after the other:
//   __shared__            -> static            (function-local statics are shared by the emulated threads of a block)
//   __syncthreads()       -> a barrier over the block's threads
//   threadIdx / blockIdx  -> thread_local
//   atomicAdd/And/Max/CAS -> GCC __atomic builtins
// The kernel source is reached through -I and a sed-patched copy written to oracle/_ref/ at build time (the only
// change: the `extern __shared__ int DynamicSharedMemory[];` declarations, which have no host equivalent).
#ifndef MPGOS_B200_ORACLE_REF_COUPLED_SHIM_H
#define MPGOS_B200_ORACLE_REF_COUPLED_SHIM_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <condition_variable>
#include <mutex>

#define __forceinline__ inline
#define __device__
#define __global__
#define __host__
#define __shared__ static

struct RefShimDim3 { int x, y, z; };
static thread_local RefShimDim3 threadIdx = {0, 0, 0};
static thread_local RefShimDim3 blockIdx  = {0, 0, 0};
static thread_local RefShimDim3 blockDim  = {1, 1, 1};
static thread_local RefShimDim3 gridDim   = {1, 1, 1};

struct RefShimBarrier
{
	std::mutex m; std::condition_variable cv; int count = 0, waiting = 0; unsigned generation = 0;
	void reset(int n) { count = n; waiting = 0; }
	void wait()
	{
		std::unique_lock<std::mutex> lk(m);
		unsigned g = generation;
		if (++waiting == count) { waiting = 0; generation++; cv.notify_all(); }
		else cv.wait(lk, [&] { return g != generation; });
	}
};
static RefShimBarrier g_block_barrier;
static inline void __syncthreads() { g_block_barrier.wait(); }

using std::isfinite;
using std::abs;
using std::pow;
using std::sqrt;
using std::sin;
using std::cos;
using std::max;
using std::min;

static inline void sincospi(double x, double* s, double* c)
{
	double r = std::remainder(x, 2.0);
	*s = std::sin(3.14159265358979323846 * r);
	*c = std::cos(3.14159265358979323846 * r);
}

static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned long long __double_as_longlong(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(unsigned long long u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline int atomicCAS(int* a, int cmp, int val) { __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val)
{ __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicAnd(int* a, int v) { return __atomic_fetch_and(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* a, int v) { int o = __atomic_load_n(a, __ATOMIC_SEQ_CST); while (o < v && !__atomic_compare_exchange_n(a, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }

#endif
