// TEST INFRASTRUCTURE ONLY -- lets g++ compile the BODY of the product's block-per-system kernel
// (include/mpgos_b200/perblock_kernel.cuh above its `#ifdef __CUDACC__`) and run it with ONE std::thread PER CUDA THREAD of a block,
// blocks one after the other:
//   __shared__ / dynamic shared memory -> statics shared by the emulated threads
//   __syncthreads / __syncwarp / named barriers -> spinning barriers over the block / warp / the given thread count
//   __shfl_xor_sync, __all_sync, __any_sync -> exchange through a per-warp buffer between two warp barriers
//   the FP64 tensor-core instruction mma.m8n8k4 -> the same product formed from the 32 lanes' fragments
// The product itself has no CPU path; nothing under include/ or the package refers to this directory.
#ifndef MPGOS_B200_TESTS_CUDA_BLOCK_ON_HOST_H
#define MPGOS_B200_TESTS_CUDA_BLOCK_ON_HOST_H

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <cuda_runtime.h>      // host side of the toolkit headers (float4, the qualifiers expand to nothing under g++)

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
#ifdef __shared__
#undef __shared__
#endif
#define __shared__ static

struct host_dim3 { unsigned x, y, z; };
static thread_local host_dim3 host_threadIdx = {0, 0, 0}, host_blockIdx = {0, 0, 0}, host_blockDim = {1, 1, 1}, host_gridDim = {1, 1, 1};
#define threadIdx host_threadIdx
#define blockIdx  host_blockIdx
#define blockDim  host_blockDim
#define gridDim   host_gridDim

// sense-reversing barrier; the emulated threads outnumber the cores, so waiting yields
struct HostBarrier
{
	std::atomic<int> arrived{0};
	std::atomic<unsigned> generation{0};
	void wait(int count)
	{
		const unsigned g = generation.load(std::memory_order_acquire);
		if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == count) { arrived.store(0, std::memory_order_relaxed); generation.fetch_add(1, std::memory_order_acq_rel); }
		else while (generation.load(std::memory_order_acquire) == g) std::this_thread::yield();
	}
};
enum { HOST_MAX_WARPS = 32 };
struct HostWarp { HostBarrier bar; unsigned long long x[32]; int vote[32]; double db[32]; float fa[128], fb[64]; };
static HostBarrier g_block_bar, g_named_bar[16];
static HostWarp    g_warp[HOST_MAX_WARPS];
static std::atomic<int> g_block_or{0};
alignas(16) static unsigned char g_dynamic_shared[227 * 1024];      // the most a block can have on sm_100a

static inline HostWarp& host_warp() { return g_warp[host_threadIdx.x >> 5]; }
static inline void __syncthreads() { g_block_bar.wait((int)host_blockDim.x); }
#ifdef KOH_BREAK_SYNCWARP      // negative control of the ThreadSanitizer run: a slot barrier that does not synchronise must be reported
static inline void __syncwarp() {}
#else
static inline void __syncwarp() { host_warp().bar.wait(32); }
#endif
static inline int  __syncthreads_or(int p)
{
	if (p) g_block_or.store(1);
	__syncthreads();
	const int r = g_block_or.load();
	__syncthreads();
	if (host_threadIdx.x == 0) g_block_or.store(0);
	__syncthreads();
	return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o)
{
	static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
	HostWarp& w = host_warp();  const int lane = host_threadIdx.x & 31;
	std::memcpy(&w.x[lane], &v, sizeof(T));
	w.bar.wait(32);
	T r;  std::memcpy(&r, &w.x[lane ^ o], sizeof(T));
	w.bar.wait(32);
	return r;
}
static inline int __all_sync(unsigned, int p)
{
	HostWarp& w = host_warp();  w.vote[host_threadIdx.x & 31] = p ? 1 : 0;
	w.bar.wait(32);
	int r = 1;  for (int i = 0; i < 32; i++) r &= w.vote[i];
	w.bar.wait(32);
	return r;
}
static inline int __any_sync(unsigned, int p)
{
	HostWarp& w = host_warp();  w.vote[host_threadIdx.x & 31] = p ? 1 : 0;
	w.bar.wait(32);
	int r = 0;  for (int i = 0; i < 32; i++) r |= w.vote[i];
	w.bar.wait(32);
	return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int  __double2hiint(double x) { long long b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int  __float_as_int(float x) { int b; std::memcpy(&b, &x, 4); return b; }
static inline unsigned __float_as_uint(float x) { unsigned b; std::memcpy(&b, &x, 4); return b; }
static inline float __uint_as_float(unsigned b) { float x; std::memcpy(&x, &b, 4); return x; }
static inline float __double2float_rn(double x) { return (float)x; }
static inline unsigned __ballot_sync(unsigned, int p)
{
	HostWarp& w = host_warp();  w.vote[host_threadIdx.x & 31] = p ? 1 : 0;
	w.bar.wait(32);
	unsigned r = 0;  for (int i = 0; i < 32; i++) r |= (unsigned)w.vote[i] << i;
	w.bar.wait(32);
	return r;
}
static inline int  __popc(unsigned v) { return __builtin_popcount(v); }
static inline int  __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int  __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src)
{
	HostWarp& w = host_warp();  const int lane = host_threadIdx.x & 31;
	std::memcpy(&w.x[lane], &v, sizeof(T));
	w.bar.wait(32);
	T r;  std::memcpy(&r, &w.x[src & 31], sizeof(T));
	w.bar.wait(32);
	return r;
}
template <class T> static inline T __shfl_down_sync(unsigned, T, int) { std::abort(); }
static inline void sincospi(double x, double* s, double* c)
{
	const double r = std::remainder(x, 2.0);       // as the checker's stand-in (oracle/ref_coupled_shim.h)
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

// stand-ins for include/mpgos_b200/ptx_primitives.cuh (its include guard is taken here)
#define MPGOS_B200_PTX_PRIMITIVES_CUH
#define MPGOS_B200_DYNAMIC_SHARED_MEMORY(name) unsigned char* const name = g_dynamic_shared
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
static inline unsigned long long f32x2_pack(float lo, float hi) { unsigned a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4); return (unsigned long long)a | ((unsigned long long)b << 32); }
static inline void f32x2_unpack(unsigned long long v, float& lo, float& hi) { unsigned a = (unsigned)v, b = (unsigned)(v >> 32); std::memcpy(&lo, &a, 4); std::memcpy(&hi, &b, 4); }
static inline unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{ float a0, a1, b0, b1, c0, c1; f32x2_unpack(a, a0, a1); f32x2_unpack(b, b0, b1); f32x2_unpack(c, c0, c1); return f32x2_pack(std::fma(a0, b0, c0), std::fma(a1, b1, c1)); }
static inline unsigned long long f32x2_mul(unsigned long long a, unsigned long long b)
{ float a0, a1, b0, b1; f32x2_unpack(a, a0, a1); f32x2_unpack(b, b0, b1); return f32x2_pack(a0 * b0, a1 * b1); }
static inline void named_barrier_sync(int id, int count) { g_named_bar[id].wait(count); }
// D[8x8] += A[8x4] . B[4x8]: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2 (l%4) + {0,1}]  (PTX ISA, mma.m8n8k4 .f64);
// k ascending, one fused multiply-add per term
static inline void dmma_8x8x4(double& d0, double& d1, double a, double b)
{
	HostWarp& w = host_warp();  const int lane = host_threadIdx.x & 31;
	double* A = reinterpret_cast<double*>(w.x);                 // A[m][k] sits in lane m*4 + k
	double* B = w.db;                                           // B[k][n] sits in lane n*4 + k
	A[lane] = a;  B[lane] = b;
	w.bar.wait(32);
	const int row = lane >> 2, c0 = 2 * (lane & 3);
	for (int k = 0; k < 4; k++)
	{
		d0 = std::fma(A[row * 4 + k], B[c0 * 4 + k], d0);
		d1 = std::fma(A[row * 4 + k], B[(c0 + 1) * 4 + k], d1);
	}
	w.bar.wait(32);
}
// cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 mantissa bits kept
static inline unsigned to_tf32(float x) { unsigned u; std::memcpy(&u, &x, 4); if ((u & 0x7f800000u) != 0x7f800000u) u += 0x1000u; return u & ~0x1fffu; }
// D[16x8] += A[16x8] . B[8x8], g = lane/4, t = lane%4: a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4); b0 (k = t, n = g), b1 (k = t+4, n = g);
// d0 (g, 2t), d1 (g, 2t+1), d2 (g+8, 2t), d3 (g+8, 2t+1)  (PTX ISA, mma.m16n8k8 .tf32); operands are TF32 values already
static inline void mma_16x8x8_tf32(float (&d)[4], const float4& a, float b0, float b1)
{
	HostWarp& w = host_warp();  const int lane = host_threadIdx.x & 31, g = lane >> 2, t = lane & 3;
	float* A = w.fa;  float* B = w.fb;                          // A[m][k] at m*8 + k, B[k][n] at k*8 + n
	A[g * 8 + t] = a.x;  A[(g + 8) * 8 + t] = a.y;  A[g * 8 + t + 4] = a.z;  A[(g + 8) * 8 + t + 4] = a.w;
	B[t * 8 + g] = b0;   B[(t + 4) * 8 + g] = b1;
	w.bar.wait(32);
	for (int k = 0; k < 8; k++)
	{
		d[0] = std::fma(A[g * 8 + k], B[k * 8 + 2 * t], d[0]);
		d[1] = std::fma(A[g * 8 + k], B[k * 8 + 2 * t + 1], d[1]);
		d[2] = std::fma(A[(g + 8) * 8 + k], B[k * 8 + 2 * t], d[2]);
		d[3] = std::fma(A[(g + 8) * 8 + k], B[k * 8 + 2 * t + 1], d[3]);
	}
	w.bar.wait(32);
}
}

#define MPGOS_B200_TABLEAU_ROW_LOADS 0

#endif
