// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Host-side stand-ins for the handful of CUDA keywords, built-ins and math functions the
// reference's per-thread solver sources use, so that g++ can compile the reference's own
// kernel body (SourceCodes/SingleSystem_PerThread_Solver.cuh and the four numerics headers it
// includes) unmodified, where it lies under /root/reference, and run it once per instance on
// the CPU.  Nothing from the reference is copied: the sources are reached through -I at build
// time and the kernel-argument structs are extracted from the reference header by the
// Makefile into oracle/_ref/ (git-ignored).
#ifndef MPGOS_B200_ORACLE_REF_SHIM_H
#define MPGOS_B200_ORACLE_REF_SHIM_H

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <algorithm>

#define __forceinline__ inline
#define __device__
#define __global__
#define __shared__
#define __host__

struct RefShimDim3 { int x, y, z; };
static thread_local RefShimDim3 threadIdx = {0, 0, 0};
static thread_local RefShimDim3 blockIdx  = {0, 0, 0};
static thread_local RefShimDim3 blockDim  = {1, 1, 1};
static thread_local RefShimDim3 gridDim   = {1, 1, 1};

// `extern __shared__ int DynamicSharedMemory[];` inside the kernel resolves to this array.
// The harness always runs with PreferSharedMemory = 0, so it is never written.
int DynamicSharedMemory[4096];

using std::isfinite;
using std::abs;
using std::pow;
using std::sqrt;
using std::sin;
using std::cos;
using std::max;
using std::min;

// sin(pi x), cos(pi x) with an exact argument reduction (CUDA's sincospi is <= 1-2 ulp).
static inline void sincospi(double x, double* s, double* c)
{
	double r = std::remainder(x, 2.0);          // exact, r in [-1, 1]
	*s = std::sin(3.14159265358979323846 * r);
	*c = std::cos(3.14159265358979323846 * r);
}
static inline void sincospi(float x, float* s, float* c)
{
	float r = std::remainder(x, 2.0f);
	*s = (float)std::sin(3.14159265358979323846 * (double)r);
	*c = (float)std::cos(3.14159265358979323846 * (double)r);
}

// bit casts + CAS used by the reference's atomicFMIN/atomicFMAX helpers (coupled module only;
// they must merely compile for the per-thread harness).
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned long long __double_as_longlong(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return u; }
static inline double __longlong_as_double(unsigned long long u) { double d; std::memcpy(&d, &u, 8); return d; }
static inline int atomicCAS(int* a, int cmp, int val)
{ __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val)
{ __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }

#endif
