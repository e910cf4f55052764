// ptx_primitives.cuh -- the handful of sm_100a instructions the integration kernels reach through inline PTX: the SFU
// approximations behind the pow-free step-size controller and the packed single-precision multiply-add of the FP32 stages.
// Kept apart from the kernels so that everything else in perthread_kernel.cuh is plain C++: the test suite compiles the
// kernel body for the host against stand-ins for these few functions (tests/kernel_on_host/, test infrastructure only --
// the product has no CPU path) and walks its branches on the CPU.
#ifndef MPGOS_B200_PTX_PRIMITIVES_CUH
#define MPGOS_B200_PTX_PRIMITIVES_CUH

#include <cuda_runtime.h>

namespace mpgos_b200 {

// SFU approximations without the denormal / special-case wrappers of the CUDA math library
// (operands are range-checked by the caller).
__device__ __forceinline__ float sfu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// two single-precision lanes in one 64-bit register pair (fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2)
__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c)
{ unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long f32x2_mul(unsigned long long a, unsigned long long b)
{ unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// true if `p` holds on any lane of the warp that executes this call together with the caller.  Used only to decide whether the
// warp enters a cold region at all (inside it every lane tests its own condition again), so the result may depend on which lanes
// happen to be converged here without any lane's arithmetic depending on it (the host stand-in simply returns `p`).
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(__activemask(), p) != 0; }

// ---- block-per-system kernel (perblock_kernel.cuh) ----------------------------------------------------------------------
// named barrier over `count` threads (a slot of the block synchronises on its own, never with __syncthreads())
__device__ __forceinline__ void named_barrier_sync(int id, int count) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory"); }
// the block's dynamic shared memory as a byte array
#define MPGOS_B200_DYNAMIC_SHARED_MEMORY(name) extern __shared__ __align__(16) unsigned char name[]
// FP64 tensor-core product D[8x8] += A[8x4] . B[4x8], fragments as in the PTX ISA (mma.m8n8k4 .f64)
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b)
{
	asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// TF32 rounding and the m16n8k8 TF32 product with FP32 accumulation (3xTF32 coupling of the FP32 kernel)
__device__ __forceinline__ unsigned to_tf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_16x8x8_tf32(float (&d)[4], const float4& a, float b0, float b1)
{
	asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
	    : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
	    : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

} // namespace mpgos_b200
#endif
