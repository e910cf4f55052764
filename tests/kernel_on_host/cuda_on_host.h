// TEST INFRASTRUCTURE ONLY -- lets g++ compile the BODY of the product's thread-per-system kernel
// (include/mpgos_b200/perthread_kernel.cuh, everything above its `#ifdef __CUDACC__`) as an ordinary C++ function, so that the
// kernel's own source -- not a restatement of it -- can be walked through its branches on the CPU and compared with the checker.
// One "thread" per "block": blockDim = 1, threadIdx = 0, blockIdx = the instance.  The product itself has no CPU path; nothing
// under include/ or the package refers to this directory.
#ifndef MPGOS_B200_TESTS_CUDA_ON_HOST_H
#define MPGOS_B200_TESTS_CUDA_ON_HOST_H

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>      // host side of the toolkit headers: __device__ / __global__ / __shared__ expand to nothing under g++

#ifdef __launch_bounds__
#undef __launch_bounds__
#endif
#define __launch_bounds__(...)
#ifdef __grid_constant__
#undef __grid_constant__
#endif
#define __grid_constant__
#ifndef __noinline__
#define __noinline__
#endif

// built-in variables
struct host_dim3 { unsigned x, y, z; };
static thread_local host_dim3 host_threadIdx = {0, 0, 0}, host_blockIdx = {0, 0, 0}, host_blockDim = {1, 1, 1}, host_gridDim = {1, 1, 1};
#define threadIdx host_threadIdx
#define blockIdx  host_blockIdx
#define blockDim  host_blockDim
#define gridDim   host_gridDim

// intrinsics the kernel body uses (single-thread semantics for the warp-wide ones)
static inline void __syncthreads() {}
static inline int  __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int  __float_as_int(float x) { int b; std::memcpy(&b, &x, 4); return b; }
static inline float __double2float_rn(double x) { return (float)x; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int  __any_sync(unsigned, int p) { return p != 0; }
static unsigned char host_dynamic_shared[1 << 16];
#define MPGOS_B200_DYNAMIC_SHARED_MEMORY(name) unsigned char* const name = host_dynamic_shared
static inline int  __popc(unsigned v) { return __builtin_popcount(v); }
static inline int  __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline void sincospi(double x, double* s, double* c)
{
	const double r = std::remainder(x, 2.0);       // as the checker's stand-in (oracle/oracle_systems.h host_sincospi)
	*s = std::sin(3.14159265358979323846 * r);  *c = std::cos(3.14159265358979323846 * r);
}
static inline void sincospi(float x, float* s, float* c)
{
	const float r = std::remainder(x, 2.0f);
	*s = std::sin(3.14159265358979323846f * r);  *c = std::cos(3.14159265358979323846f * r);
}
using std::fma;  using std::fmin;  using std::fmax;  using std::fabs;  using std::pow;  using std::sin;  using std::cos;  using std::sqrt;
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int      max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline int      min(int a, int b) { return a < b ? a : b; }

// stand-ins for include/mpgos_b200/ptx_primitives.cuh (its include guard is taken here): the SFU approximations become libm
// calls (more accurate than the hardware's 2^-22; the controller's Newton steps absorb either), the packed FP32 pair two floats.
#define MPGOS_B200_PTX_PRIMITIVES_CUH
// ... and for include/mpgos_b200/fast_math.cuh: on the host both forms of the built-in systems use libm
#define MPGOS_B200_FAST_MATH_CUH
namespace mpgos_b200 { namespace fm {
template <bool FAST> struct Math
{
	template <class P> static inline P    pow_(P x, P y)             { return std::pow(x, y); }
	template <class P> static inline P    sin_(P x)                  { return std::sin(x); }
	template <class P> static inline P    cos_(P x)                  { return std::cos(x); }
	template <class P> static inline void sincospi_(P x, P* s, P* c) { sincospi(x, s, c); }
	template <class P> static inline void sincos_(P x, P* s, P* c)   { *s = std::sin(x);  *c = std::cos(x); }
	template <class P> static inline P    rcp_(P x)                  { return P(1.0) / x; }
	template <class P> static inline void pow_sincospi_sincos_(P x, P y, P a, P b, P* p, P* s1, P* c1, P* s2, P* c2)
	{ *p = pow_(x, y);  sincospi_(a, s1, c1);  sincos_(b, s2, c2); }
	template <class P> static inline void rcp2_(P a, P b, P* ra, P* rb) { *ra = P(1.0) / a;  *rb = P(1.0) / b; }
};
} }
namespace mpgos_b200 {
static inline bool warp_any(bool p) { return p; }      // a lane's own condition is a valid answer (see ptx_primitives.cuh)
static inline float sfu_lg2(float x) { return std::log2(x); }
static inline float sfu_ex2(float x) { return std::exp2(x); }
static inline float sfu_rcp(float x) { return 1.0f / x; }
static inline unsigned long long f32x2_pack(float lo, float hi) { unsigned long long r; unsigned a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4); r = (unsigned long long)a | ((unsigned long long)b << 32); return r; }
static inline void f32x2_unpack(unsigned long long v, float& lo, float& hi) { unsigned a = (unsigned)v, b = (unsigned)(v >> 32); std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4); }
static inline unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{ float a0, a1, b0, b1, c0, c1; f32x2_unpack(a, a0, a1); f32x2_unpack(b, b0, b1); f32x2_unpack(c, c0, c1); return f32x2_pack(std::fma(a0, b0, c0), std::fma(a1, b1, c1)); }
static inline unsigned long long f32x2_mul(unsigned long long a, unsigned long long b)
{ float a0, a1, b0, b1; f32x2_unpack(a, a0, a1); f32x2_unpack(b, b0, b1); return f32x2_pack(a0 * b0, a1 * b1); }
}

// the volatile constant-memory row loads are PTX: the host build reads the same table through the generic path
#define MPGOS_B200_TABLEAU_ROW_LOADS 0

#endif
