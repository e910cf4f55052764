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

} // namespace mpgos_b200
#endif
