// perthread_kernel.cuh -- thread-per-system ensemble integrator for sm_100a (B200).
//
// One lane integrates one ODE instance through one integration phase (the reference's
// SingleSystem_PerThread kernel, SourceCodes/SingleSystem_PerThread_Solver.cuh:11-437, with its
// steppers ..._ExplicitRungeKutta_Steppers.cuh:7-290, error controllers
// ..._ErrorControllers.cuh:6-86, event step control ..._EventHandling.cuh:6-45 and dense output
// ..._DenseOutput.cuh:6-58).  Same algorithm and decision semantics, different machine mapping:
//
//  * state, stage vectors, control parameters, accessories and event values live in registers
//    (all loops over SD/NE/... are fully unrolled; callbacks are inlined through `Sys`);
//  * solver options and tolerances travel in the by-value kernel argument (constant bank), so
//    they cost neither registers nor loads; shared parameters are staged in shared memory;
//  * the kernel is bound by instruction issue, an FP64 instruction taking two issue slots: everything
//    below is about executing fewer instructions per step;
//  * the RKCK45 controller needs no FP64 division and no pow(): the limiting component is found
//    by cross-multiplication and (tol/err)^(1/5 | 1/4) by division-free Newton steps from a
//    single-precision seed (3e-14; exotic operand ranges take the exact slow path); whether a step-size limit can
//    be active at all is screened in single precision;
//  * the Cash-Karp tableau comes from constant memory row by row, dyadic entries and exact rational relations between
//    entries as instruction immediates, so that few constants compete for the 63 uniform registers;
//  * DispatchOrder: from the second phase on, thread blocks integrate the tiles of the sweep longest first (previous
//    phase's step counts), because a launch ends with whatever its last blocks cost;
//  * LaneRefill = n: persistent warps.  Lanes that finish their instance store it and go idle;
//    once a warp ballot counts at least n idle lanes, the warp takes that many instance ids from
//    a device-side queue with ONE atomic and the idle lanes load them together (coalesced when
//    n is large), so a warp never idles long behind its slowest lane (the "tailing effect" the
//    reference's README describes) and never pays the divergent load/store path once per lane.
//    Per-instance arithmetic does not depend on the lane an instance runs on, so results are
//    identical in all mappings.
//
// The kernel is instantiated in the translation unit that defines the callbacks; the host side
// reaches it through a plain function pointer (mpgos_perthread_launcher, mpgos_b200.h).
#ifndef MPGOS_B200_PERTHREAD_KERNEL_CUH
#define MPGOS_B200_PERTHREAD_KERNEL_CUH

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "../mpgos_b200.h"
#include "ptx_primitives.cuh"

namespace mpgos_b200 {

constexpr int kMaxThreadsPerBlock = 256;
// MINB (resident blocks per SM the kernel is compiled for) sets the register cap: 1..4 with 256-thread blocks
// (255 / 128 / 80 / 64 registers), 5..8 with blocks of at most 128 threads (96 / 80 / 72 / 64 registers).
__host__ __device__ constexpr int launch_bound_threads(int minb) { return minb >= 5 ? 128 : kMaxThreadsPerBlock; }
#ifndef MPGOS_B200_MIN_BLOCKS_PER_SM
#define MPGOS_B200_MIN_BLOCKS_PER_SM 2
#endif
constexpr int kAlgRK4 = 0, kAlgRKCK45 = 1;

// Compile-time choices of the RKCK45 arithmetic (see the comments at their use).  The parity build
// (libmpgos_b200_strict.so) sets MPGOS_B200_STAGE_FORM=0: the reference's expression shapes.
#ifndef MPGOS_B200_STAGE_FORM
#define MPGOS_B200_STAGE_FORM 1
#endif
#ifndef MPGOS_B200_NEWTON_F64
#define MPGOS_B200_NEWTON_F64 1
#endif

#ifndef MPGOS_B200_ACCEPT_FROM_LIMITING
#define MPGOS_B200_ACCEPT_FROM_LIMITING 1
#endif

template <int N> struct AtLeastOne { static constexpr int value = (N > 0 ? N : 1); };

// By-value kernel argument (constant bank).  Role of the reference's Struct_GlobalVariables +
// Struct_SolverOptions + Struct_ThreadConfiguration (Interface.cuh:72-115).
// shared parameters travel by value (constant bank -> uniform registers) when there are at most this many of each kind
constexpr int kSharedByValueLimit = 8;
__host__ __device__ constexpr int shared_by_value_count(int n, int other) { return (n <= kSharedByValueLimit && other <= kSharedByValueLimit && n > 0) ? n : 1; }

template <class P, int SD, int NE, int NSP = 0, int NISP = 0>
struct PerThreadArgs
{
	P*   TimeDomain;  P* ActualState;  P* ActualTime;  P* ControlParameters;
	P*   SharedParameters;  int* IntegerSharedParameters;
	P*   Accessories;  int* IntegerAccessories;
	int* DenseOutputIndex;  P* DenseOutputTimeInstances;  P* DenseOutputStates;
	unsigned long long* StepsAttempted;
	unsigned long long* StepsAccepted;
	unsigned int*       WorkQueue;
	unsigned int*       LastPhaseSteps;     // [Stride] attempted steps of this phase (cost history for the dispatch order)
	const int*          DispatchOrder;      // tile ids in dispatch order, or nullptr: block b integrates tile b
	int  Stride;             // instances held by the buffers (SoA stride)
	int  Active;             // instances [0, Active) are integrated
	int  InstanceOffset;     // global id of local instance 0
	int  GlobalInstances;    // NT as the callbacks see it
	int  ShardIndex, ShardCount, ShardTile;      // tile-cyclic shard of a multi-device solver (mpgos_b200.h); ShardCount <= 1: none
	int  DenseCapacity;      // NDO
	int  DenseOutputSaveFrequency;
	int  DeviceMessages;
	int  RefillThreshold;    // LaneRefill: idle lanes per warp that trigger a refill (1..32)
	int  EventScreen;        // 1: every EventTolerance >= 0, so an event whose value keeps its sign and stays outside the band cannot fire
	int  SharedInGlobal;     // more shared parameters than fit by value: 0 = staged in dynamic shared memory, 1 = read from global memory
	P    InitialTimeStep, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit;
	P    DenseOutputMinimumTimeStep;
	P    MinimumTimeStepGuard;    // MinimumTimeStep * 1.01 (ErrorControllers.cuh:63,72), formed once on the host in P
	// single-precision copies of the limits, each moved inwards by 1e-4: a step whose single-precision step-size factor
	// and new step lie strictly inside them cannot touch a limit in working precision
	float GuardAbove, ShrinkAbove, GrowBelow, MinimumAbove, MaximumBelow;
	P    RelativeTolerance[SD];
	P    AbsoluteTolerance[SD];
	P    EventTolerance[AtLeastOne<NE>::value];
	int  EventDirection[AtLeastOne<NE>::value];
	// Shared parameters by value (at most kSharedByValueLimit of each kind; values as of the last host-to-device copy): in the
	// constant bank they cost neither vector registers nor shared-memory loads -- FP64 instructions take them as uniform-register
	// operands.  The impact oscillator (4 shared parameters) lost its 68-byte register spill to this.
	P    SharedValues[shared_by_value_count(NSP, NISP)];
	int  IntegerSharedValues[shared_by_value_count(NISP, NSP)];
};

// ---------------------------------------------------------------------------------------------
// Cash-Karp tableau (Steppers.cuh:167-285) in constant memory.  On sm_100a ptxas feeds FP64
// instructions with constants from uniform registers; literals are rebuilt with two UMOVs per
// use once the 63 uniform registers run out (66 UMOVs per step for Lorenz), whereas constant
// memory is fetched two doubles at a time (LDCU.128).  Same values as the reference's literals.
// Rows start on even indices so that a row is fetched with 16-byte loads (LDCU.128).
enum CashKarp
{
	CK_C2, CK_C3, CK_C4, CK_C6,
	CK_A21, CK_PAD0,
	CK_A31, CK_A32,
	CK_A41, CK_A42, CK_A43, CK_PAD1,
	CK_A51, CK_A52, CK_A53, CK_A54,
	CK_A61, CK_A62, CK_A63, CK_A64, CK_A65, CK_PAD2,
	CK_B1, CK_B3, CK_B4, CK_B6,
	CK_E1, CK_E3, CK_E4, CK_E5, CK_E6, CK_ERRFLOOR,
	CK_SAFETY, CK_PAD3,
	CK_COUNT
};

#define MPGOS_B200_CK_TABLE(T)                                                                                     \
	{ (T)(1.0 / 5.0), (T)(3.0 / 10.0), (T)(3.0 / 5.0), (T)(7.0 / 8.0),                                              \
	  (T)(1.0 / 5.0), (T)0,                                                                                         \
	  (T)(3.0 / 40.0), (T)(9.0 / 40.0),                                                                             \
	  (T)(3.0 / 10.0), (T)(-9.0 / 10.0), (T)(6.0 / 5.0), (T)0,                                                      \
	  (T)(-11.0 / 54.0), (T)(5.0 / 2.0), (T)(-70.0 / 27.0), (T)(35.0 / 27.0),                                       \
	  (T)(1631.0 / 55296.0), (T)(175.0 / 512.0), (T)(575.0 / 13824.0), (T)(44275.0 / 110592.0), (T)(253.0 / 4096.0), (T)0, \
	  (T)(37.0 / 378.0), (T)(250.0 / 621.0), (T)(125.0 / 594.0), (T)(512.0 / 1771.0),                               \
	  (T)(37.0 / 378.0 - 2825.0 / 27648.0), (T)(250.0 / 621.0 - 18575.0 / 48384.0),                                 \
	  (T)(125.0 / 594.0 - 13525.0 / 55296.0), (T)(0.0 - 277.0 / 14336.0), (T)(512.0 / 1771.0 - 1.0 / 4.0), (T)(1e-18), \
	  (T)(0.9), (T)0 }

static __constant__ __align__(16) double kCashKarpF64[CK_COUNT] = MPGOS_B200_CK_TABLE(double);
static __constant__ __align__(16) float  kCashKarpF32[CK_COUNT] = MPGOS_B200_CK_TABLE(float);

template <class P> __device__ __forceinline__ P ck(int i);
template <> __device__ __forceinline__ double ck<double>(int i) { return kCashKarpF64[i]; }
#ifdef MPGOS_B200_F32_TABLE_IN_CONSTANT_MEMORY
template <> __device__ __forceinline__ float  ck<float>(int i)  { return kCashKarpF32[i]; }
#else
// FP32 instructions take 32-bit immediates: with a compile-time index the literal folds into the FFMA/FMUL itself and
// needs neither a uniform register nor a load (the FP32 kernel is issue-bound: every instruction saved counts).
template <> __device__ __forceinline__ float  ck<float>(int i)  { constexpr float t[CK_COUNT] = MPGOS_B200_CK_TABLE(float); return t[i]; }
#endif

// One tableau row, fetched where it is used.  Left to itself the compiler hoists all 28 constant loads out of the
// step loop, runs out of uniform registers, parks the rest in vector registers (some even on the stack) and moves
// them back with two R2UR per constant and step.  A volatile load cannot be hoisted: a row costs one LDCU.128 per two
// constants and lives in uniform registers only while its stage is formed.
#ifndef MPGOS_B200_TABLEAU_ROW_LOADS
#define MPGOS_B200_TABLEAU_ROW_LOADS 1
#endif
#ifndef MPGOS_B200_TABLEAU_IMMEDIATES
#define MPGOS_B200_TABLEAU_IMMEDIATES 1
#endif
#ifndef MPGOS_B200_TABLEAU_COMMON_FACTORS
#define MPGOS_B200_TABLEAU_COMMON_FACTORS 1
#endif
template <class P, int FIRST, int N> struct CkRow
{
	static __device__ __forceinline__ void load(P* c)
	{
		#pragma unroll
		for (int j = 0; j < N; j++) c[j] = ck<P>(FIRST + j);
	}
};
#if MPGOS_B200_TABLEAU_ROW_LOADS
template <int FIRST, int N> struct CkRow<double, FIRST, N>
{
	static_assert(FIRST % 2 == 0, "rows start on 16-byte boundaries");
	static __device__ __forceinline__ void load(double* c)
	{
		#pragma unroll
		for (int j = 0; j + 1 < N; j += 2)
			asm volatile("ld.const.v2.f64 {%0, %1}, [%2];" : "=d"(c[j]), "=d"(c[j + 1]) : "l"(__cvta_generic_to_constant(&kCashKarpF64[FIRST + j])));
		if (N % 2 == 1)
			asm volatile("ld.const.f64 %0, [%1];" : "=d"(c[N - 1]) : "l"(__cvta_generic_to_constant(&kCashKarpF64[FIRST + N - 1])));
#if MPGOS_B200_TABLEAU_IMMEDIATES
		// dyadic entries whose low mantissa word is zero fold into the instruction as 32-bit immediates: no uniform register
		#pragma unroll
		for (int j = 0; j < N; j++)
		{
			if (FIRST + j == CK_A52) c[j] = 5.0 / 2.0;
			if (FIRST + j == CK_A62) c[j] = 175.0 / 512.0;
			if (FIRST + j == CK_A65) c[j] = 253.0 / 4096.0;
		}
#endif
	}
};
#endif

// ---------------------------------------------------------------------------------------------
// numeric helpers

template <class P> struct Num;

// `top(x)` is the word that holds sign and exponent; for positive normal numbers it orders like
// the value, and one unsigned comparison of (top - kLow) against kSpan tests
// "positive, finite, and 2^-100 <= x < 2^30" (2^-60 <= x < 2^30 in single precision).
template <> struct Num<double>
{
	static __device__ __forceinline__ bool finite(double x) { return (__double2hiint(x) & 0x7ff00000) != 0x7ff00000; }
	static __device__ __forceinline__ unsigned top(double x) { return (unsigned)__double2hiint(x); }
	static constexpr unsigned kLow = (1023u - 100u) << 20, kSpan = 130u << 20;
	static constexpr int kRatioLimit = 98 << 20;    // top(tol) - top(err) below this  =>  tol/err < 2^99 < 1e30
	static constexpr int kRatioFloor = 60 << 20;    // and above minus this  =>  tol/err > 2^-61: (tol/err)^(4/4) stays a normal float
	static constexpr int kNewton = MPGOS_B200_NEWTON_F64;
	static constexpr bool kNewtonSeed = (MPGOS_B200_NEWTON_F64 == 1);
	static __device__ __forceinline__ float  narrow(double x) { return __double2float_rn(x); }
};

template <> struct Num<float>
{
	static __device__ __forceinline__ bool finite(float x) { return (__float_as_int(x) & 0x7f800000) != 0x7f800000; }
	static __device__ __forceinline__ unsigned top(float x) { return (unsigned)__float_as_int(x); }
	static constexpr unsigned kLow = (127u - 60u) << 23, kSpan = 90u << 23;
	static constexpr int kRatioLimit = 98 << 23;
	static constexpr int kRatioFloor = 60 << 23;
	static constexpr int kNewton = 1;
	static constexpr bool kNewtonSeed = false;
	static __device__ __forceinline__ float  narrow(float x) { return x; }
};

// y = (tol/err)^(1/5) if accept, (tol/err)^(1/4) otherwise, without FP64 division or pow.
// With a = err/tol, y = a^(-1/n) is the fixed point of the division-free Newton map
//     y <- y + y (1 - a y^n)/n  =  y + y (tol - err y^n) / (n tol).
// The residual tol - err y^n is one FMA in working precision; the factor 1/(n tol) only scales a
// correction of relative size <= 1e-5, so a single-precision reciprocal is exact enough.  The
// seed comes from the SFU (lg2/ex2) with relative error <~ 4e-6.
//   MPGOS_B200_NEWTON_F64 == 1 (default): one Newton step in FP32 (FP32 pipe, idle in this kernel) takes the seed to
//     ~1e-7, one step in FP64 to ~3e-14.  The step-size factor then differs from the exact root by <~ 100 ulp, which
//     moves the next step's error estimate by 5 x 3e-14 relative -- five orders below the rounding noise of the embedded
//     error estimate itself (a difference of nearly cancelling stage terms: 1e-16 |k| dt against err ~ 1e-8 |x|).
//   MPGOS_B200_NEWTON_F64 == 2: two steps in FP64 straight from the seed (~1 ulp), 6 more FP64-pipe instructions per step.
template <class P>
__device__ __forceinline__ P ratio_root_fast(P tol, P err, bool accept, float& y_single)
{
	const float tf  = Num<P>::narrow(tol);
	const float ef  = Num<P>::narrow(err);
	const float inv = accept ? 0.2f : 0.25f;
	float yf = sfu_ex2((sfu_lg2(tf) - sfu_lg2(ef)) * inv);
	const float rnf = sfu_rcp(tf) * inv;
	if (Num<P>::kNewtonSeed)
	{
		const float y2 = yf * yf;
		const float y4 = y2 * y2;
		const float yn = accept ? y4 * yf : y4;
		yf = fmaf(yf * fmaf(-ef, yn, tf), rnf, yf);
	}
	y_single = yf;                      // the root to ~1e-6: enough to tell whether a step-size limit can be active
	P y = (P)yf;
	const P rn = (P)rnf;
	#pragma unroll
	for (int it = 0; it < Num<P>::kNewton; it++)
	{
		const P y2 = y * y;
		const P y4 = y2 * y2;
		const P yn = accept ? y4 * y : y4;
		const P res = fma(-err, yn, tol);
		y = fma(y * res, rn, y);
	}
	return y;
}

// ---------------------------------------------------------------------------------------------
// Component-wise helpers of the stages.  FP32: two components per instruction with the packed single-precision
// instructions of sm_100 (fma.rn.f32x2 / mul.rn.f32x2 -> FFMA2 / FMUL2; a scalar factor is broadcast by the instruction
// itself) -- the FP32 kernel is issue-bound, a packed instruction takes one issue slot for two FMAs.  Same roundings as the
// scalar instructions (results are bit-identical).  Measured: Lorenz (3 components) 0.901 -> 0.838 ms, impact 0.563 -> 0.545;
// with 2 components the pairing constrains register allocation more than it saves (Keller-Miksis 1.305 -> 1.346 ms), so
// the packed form is used from 3 components on.
#ifndef MPGOS_B200_PACKED_F32
#define MPGOS_B200_PACKED_F32 1
#endif
template <class P, int N, bool PACKED = (MPGOS_B200_PACKED_F32 != 0) && (sizeof(P) == 4) && (N >= 3)>
struct Vec
{
	// out = c * a + b, out = c * a   (out may alias a or b)
	static __device__ __forceinline__ void fma_c(P* out, P c, const P* a, const P* b)
	{
		#pragma unroll
		for (int i = 0; i < N; i++) out[i] = fma(c, a[i], b[i]);
	}
	static __device__ __forceinline__ void mul_c(P* out, P c, const P* a)
	{
		#pragma unroll
		for (int i = 0; i < N; i++) out[i] = c * a[i];
	}
};
template <int N>
struct Vec<float, N, true>
{
	static __device__ __forceinline__ void fma_c(float* out, float c, const float* a, const float* b)
	{
		const unsigned long long cc = f32x2_pack(c, c);
		#pragma unroll
		for (int i = 0; i + 1 < N; i += 2)
			f32x2_unpack(f32x2_fma(cc, f32x2_pack(a[i], a[i + 1]), f32x2_pack(b[i], b[i + 1])), out[i], out[i + 1]);
		if (N & 1) out[N - 1] = fmaf(c, a[N - 1], b[N - 1]);
	}
	static __device__ __forceinline__ void mul_c(float* out, float c, const float* a)
	{
		const unsigned long long cc = f32x2_pack(c, c);
		#pragma unroll
		for (int i = 0; i + 1 < N; i += 2)
			f32x2_unpack(f32x2_mul(cc, f32x2_pack(a[i], a[i + 1])), out[i], out[i + 1]);
		if (N & 1) out[N - 1] = c * a[N - 1];
	}
};

// device-side warnings of the step loop (the reference's texts), out of line: a printf call inside the loop body costs the
// ordinary step register moves and flag bookkeeping around it even when it never runs
static __device__ __noinline__ void device_message(int which, int gid)
{   // [cold>
	if (which == 0) printf("Error: State is not a finite number. Try to use smaller step size. (thread id: %d)\n", gid);
	else            printf("Warning: Event cannot be detected without reducing the step size below the minimum! Event detection omitted!, (thread id: %d)\n", gid);
}   // <cold]

// ---------------------------------------------------------------------------------------------
// The cold half of the RKCK45 controller (ErrorControllers.cuh:38-85), out of line: every step that is near a step-size limit, at
// the minimum step, or whose tolerances / error estimates leave the fast path's operand range (including every NaN / Inf) comes
// here.  `fast`: the step-size factor of the fast path is valid and only the limits remain; otherwise the exact reference
// arithmetic (division, pow, isfinite flags) decides everything anew.  Returns bit 0: accept, bit 1: terminate.
template <class P, int SD>
struct ControllerCold
{
	P x[SD], xn[SD], err[SD];
	P dt, mult, dt_new;
	int gid;
	bool fast, update;
};

template <class P, int SD, int NE, int NSP, int NISP>
__device__ __noinline__ int controller_exact(const PerThreadArgs<P, SD, NE, NSP, NISP>* ap, ControllerCold<P, SD>* c)
{   // [cold>
	const PerThreadArgs<P, SD, NE, NSP, NISP>& a = *ap;
	const P dt = c->dt;
	const int gid = c->gid;
	P mult = c->mult;
	bool update = c->update, terminate = false, finite = true;
	const bool at_minimum = dt < a.MinimumTimeStepGuard;
	if (c->fast)
	{
		if (at_minimum)
		{
			if (a.DeviceMessages) printf("Warning: Minimum step size reached! Continue with fixed minimum step size! Tolerance cannot be guaranteed!, thread id: %d, time step: %+6.5e, min step size: %+6.5e \n", gid, (double)dt, (double)a.MinimumTimeStep);
			update = true;
		}
	}
	else
	{
		// exact reference arithmetic (division, pow) for exotic operand ranges and non-finite states
		P rel = P(1e30);
		update = true;
		for (int i = 0; i < SD; i++)
		{
			finite = finite && Num<P>::finite(c->xn[i]) && Num<P>::finite(c->err[i]);
			const P tol_i = fmax(a.RelativeTolerance[i] * fmax(fabs(c->xn[i]), fabs(c->x[i])), a.AbsoluteTolerance[i]);
			update = update && (c->err[i] < tol_i);
			rel = fmin(rel, tol_i / c->err[i]);
		}
		mult = P(0.9) * (pow)(rel, update ? P(1.0 / 5.0) : P(1.0 / 4.0));
		if (!Num<P>::finite(mult)) finite = false;
		if (!finite)
		{
			mult = a.TimeStepShrinkLimit;
			update = false;
			if (at_minimum)
			{
				if (a.DeviceMessages) printf("Error: State is not a finite number even with the minimal step size. Try to use less stringent tolerances. (thread id: %d)\n", gid);
				terminate = true;
			}
		}
		else if (at_minimum)
		{
			if (a.DeviceMessages) printf("Warning: Minimum step size reached! Continue with fixed minimum step size! Tolerance cannot be guaranteed!, thread id: %d, time step: %+6.5e, min step size: %+6.5e \n", gid, (double)dt, (double)a.MinimumTimeStep);
			update = true;
		}
	}
	// all operands positive: plain compare-selects instead of fmin/fmax
	mult = (mult > a.TimeStepGrowLimit) ? a.TimeStepGrowLimit : mult;
	mult = (mult < a.TimeStepShrinkLimit) ? a.TimeStepShrinkLimit : mult;
	P dt_new = dt * mult;
	dt_new = (dt_new > a.MaximumTimeStep) ? a.MaximumTimeStep : dt_new;
	dt_new = (dt_new < a.MinimumTimeStep) ? a.MinimumTimeStep : dt_new;
	c->dt_new = dt_new;
	return (update ? 1 : 0) | (terminate ? 2 : 0);
}   // <cold]

// ---------------------------------------------------------------------------------------------
// One instance's registers.

template <class P, int SD, int NCP, int NE, int NA, int NIA>
struct LaneState
{
	P   td[2];
	P   x[SD];
	P   t;
	P   cpar[AtLeastOne<NCP>::value];
	P   acc[AtLeastOne<NA>::value];
	int acci[AtLeastOne<NIA>::value];
	P   ev[AtLeastOne<NE>::value];       // event values at the actual state
	P   dt_new;
	P   dense_next_time;
	int dense_index;
	int dense_skipped;
	int gid;                             // the instance's id in the whole sweep: what the callbacks receive as `tid`
	unsigned attempted, accepted;
};

// ---------------------------------------------------------------------------------------------
// The kernel.  Sys supplies the six callbacks as static device functions with the reference's
// argument lists (see include/SingleSystem_PerThread_Interface.cuh for the adapter that binds the
// user's PerThread_* functions).

template <class Sys, class P, int SD, int NCP, int NSP, int NISP, int NE, int NA, int NIA, bool DENSE, int ALG, bool REFILL, int MINB>
__global__ void __launch_bounds__(launch_bound_threads(MINB), MINB)
perthread_kernel(const __grid_constant__ PerThreadArgs<P, SD, NE, NSP, NISP> a)
{
	// Shared parameters (Solver.cuh:66-88).  Up to kSharedByValueLimit of each kind arrive by value in the kernel argument and
	// are copied into a thread-local array: as long as the callbacks only read them the compiler forwards the constant-bank
	// values (uniform registers), a callback that writes one gets a private copy (in the reference such a write goes to the
	// block's shared memory, unsynchronised).  Larger sets are staged in dynamic shared memory once per block -- with the
	// __syncthreads() the reference omits -- or, when they do not fit or PreferSharedMemory = 0, read from global memory.
	constexpr bool kSharedByValue = (NSP <= kSharedByValueLimit && NISP <= kSharedByValueLimit);
	P   r_par[kSharedByValue ? AtLeastOne<NSP>::value : 1];
	int r_pari[kSharedByValue ? AtLeastOne<NISP>::value : 1];
	P*   p_par  = r_par;
	int* p_pari = r_pari;
	if (kSharedByValue)
	{
		#pragma unroll
		for (int i = 0; i < NSP; i++)  r_par[i]  = a.SharedValues[i];
		#pragma unroll
		for (int i = 0; i < NISP; i++) r_pari[i] = a.IntegerSharedValues[i];
	}
	else if (a.SharedInGlobal)
	{
		p_par  = a.SharedParameters;
		p_pari = a.IntegerSharedParameters;
	}
	else
	{
		MPGOS_B200_DYNAMIC_SHARED_MEMORY(dyn_smem);
		P*   s_par  = reinterpret_cast<P*>(dyn_smem);
		int* s_pari = reinterpret_cast<int*>(s_par + NSP);
		for (int i = threadIdx.x; i < NSP; i += blockDim.x)  s_par[i]  = a.SharedParameters[i];
		for (int i = threadIdx.x; i < NISP; i += blockDim.x) s_pari[i] = a.IntegerSharedParameters[i];
		__syncthreads();
		p_par  = s_par;
		p_pari = s_pari;
	}

	const int NTG    = a.GlobalInstances;
	const int stride = a.Stride;

	LaneState<P, SD, NCP, NE, NA, NIA> s;

	// ---- load one instance into the lane's registers (Solver.cuh:117-193) --------------------------
	auto load = [&](const int inst)
	{
		// ---- load (coalesced on the first fill), Solver.cuh:117-142 ----------------------
		#pragma unroll
		for (int i = 0; i < 2; i++)   s.td[i]   = a.TimeDomain[inst + i * stride];
		#pragma unroll
		for (int i = 0; i < SD; i++)  s.x[i]    = a.ActualState[inst + i * stride];
		#pragma unroll
		for (int i = 0; i < NCP; i++) s.cpar[i] = a.ControlParameters[inst + i * stride];
		#pragma unroll
		for (int i = 0; i < NA; i++)  s.acc[i]  = a.Accessories[inst + i * stride];
		#pragma unroll
		for (int i = 0; i < NIA; i++) s.acci[i] = a.IntegerAccessories[inst + i * stride];
		s.t = a.ActualTime[inst];
		s.dense_index = DENSE ? a.DenseOutputIndex[inst] : 0;
		s.dense_skipped = 0;
		s.dense_next_time = P(0);
		s.dt_new = a.InitialTimeStep;
		s.attempted = 0;  s.accepted = 0;

		P dt0 = a.InitialTimeStep;
		// global id: the instance's own index, or its place in the sweep a multi-device solver deals out in tiles
		s.gid = (a.ShardCount > 1 ? ((inst / a.ShardTile) * a.ShardCount + a.ShardIndex) * a.ShardTile + inst % a.ShardTile : inst) + a.InstanceOffset;
		const int gid = s.gid;
		Sys::Initialization(gid, NTG, s.dense_index, s.t, dt0, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);
		if (NE > 0)
			Sys::EventFunction(gid, NTG, s.ev, s.t, dt0, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);
		if (DENSE)   // first point of the phase, stored without looking at the index (Solver.cuh:179-193)
		{
			if (s.dense_index < a.DenseCapacity)
			{
				a.DenseOutputTimeInstances[inst + (size_t)s.dense_index * stride] = s.t;
				#pragma unroll
				for (int i = 0; i < SD; i++)
					a.DenseOutputStates[inst + (size_t)i * stride + (size_t)s.dense_index * stride * SD] = s.x[i];
			}
			s.dense_index++;
			s.dense_skipped = 0;
			s.dense_next_time = fmin(s.t + a.DenseOutputMinimumTimeStep, s.td[1]);
		}
	};

	// ---- one attempted step (Solver.cuh:197-394); returns true when the phase of this instance ends ----
	auto step = [&](const int inst, P& dt) -> bool
	{
		// ======================== one attempted step (Solver.cuh:197-394) =====================
		const int gid = s.gid;
		bool update = true, finite = true, terminate = false, end_reached = false;
		dt = s.dt_new;
		{
			const P remaining = s.td[1] - s.t;
			if (dt > remaining) { dt = remaining; end_reached = true; }
		}
		s.attempted++;

		P xn[SD];
		if (ALG == kAlgRK4)
		{
			// Steppers.cuh:7-117 -- xn accumulates k1 + 2 k2 + 2 k3 (+ k4)
			P X[SD], k[SD], T;
			const P half = P(0.5) * dt;
			const P sixth = P(1.0 / 6.0) * dt;
			Sys::OdeFunction(gid, NTG, xn, s.x, s.t, s.cpar, p_par, p_pari, s.acc, s.acci);
			T = s.t + half;
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + xn[i] * half;
			Sys::OdeFunction(gid, NTG, k, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			#pragma unroll
			for (int i = 0; i < SD; i++) { xn[i] = xn[i] + P(2.0) * k[i]; X[i] = s.x[i] + k[i] * half; }
			Sys::OdeFunction(gid, NTG, k, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			T = s.t + dt;
			#pragma unroll
			for (int i = 0; i < SD; i++) { xn[i] = xn[i] + P(2.0) * k[i]; X[i] = s.x[i] + k[i] * dt; }
			Sys::OdeFunction(gid, NTG, k, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			#pragma unroll
			for (int i = 0; i < SD; i++)
			{
				xn[i] = s.x[i] + sixth * (xn[i] + k[i]);
				finite = finite && Num<P>::finite(xn[i]);
			}
			// ErrorControllers.cuh:6-20: never rejects; a non-finite state ends the phase, committed
			if (!finite)
			{
				if (a.DeviceMessages) device_message(0, gid);
				terminate = true;
			}
			s.dt_new = a.InitialTimeStep;
		}
		else
		{
			// Steppers.cuh:122-290 -- Cash-Karp 4(5)
			P X[SD], T, k1[SD], k2[SD], k3[SD], k4[SD], k5[SD], k6[SD], err[SD];
#if MPGOS_B200_STAGE_FORM == 1
			// Increment form: each slope is scaled by dt once (h_j = dt k_j), every stage is then a chain of FMAs with one
			// tableau constant as operand.  Same instruction count as the reference's shape x + dt (a1 k1 + a2 k2 + ...), but
			// no FMA reads three register pairs: on sm_100a such a DFMA occupies the FP64 pipe for 3 cycles instead of 2
			// (register-file read bandwidth; tools/microbench_fp64.cu).  Rounding differs from the reference's shape by
			// O(1 ulp) per stage, as FMA contraction already does; the parity build keeps the reference's shape (#else).
			using V = Vec<P, SD>;
			P c[6], u[SD];
			Sys::OdeFunction(gid, NTG, k1, s.x, s.t, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k1, dt, k1);

			T = s.t + dt * ck<P>(CK_C2);
			CkRow<P, CK_A21, 1>::load(c);
			V::fma_c(X, c[0], k1, s.x);
			Sys::OdeFunction(gid, NTG, k2, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k2, dt, k2);

			T = s.t + dt * ck<P>(CK_C3);
#if MPGOS_B200_TABLEAU_COMMON_FACTORS
			// Exact rational relations inside the tableau -- row 3 = (3/40)(1, 3), row 4 = (3/10)(1, -3, 4), a53 = -2 a54,
			// a64 = (77/8) a63, e5 = 4.5 e1 -- let a dyadic immediate stand in for a second tableau constant: 6 fewer constants
			// compete for the 63 uniform registers (fewer R2UR moves of parked constants per step), same instruction count.
			CkRow<P, CK_A31, 1>::load(c);
			V::fma_c(u, P(3.0), k2, k1);
			V::fma_c(X, c[0], u, s.x);
#else
			CkRow<P, CK_A31, 2>::load(c);
			V::fma_c(u, c[0], k1, s.x);
			V::fma_c(X, c[1], k2, u);
#endif
			Sys::OdeFunction(gid, NTG, k3, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k3, dt, k3);

			T = s.t + dt * ck<P>(CK_C4);
#if MPGOS_B200_TABLEAU_COMMON_FACTORS
			CkRow<P, CK_A41, 1>::load(c);
			V::fma_c(u, P(-3.0), k2, k1);
			V::fma_c(u, P(4.0), k3, u);
			V::fma_c(X, c[0], u, s.x);
#else
			CkRow<P, CK_A41, 3>::load(c);
			V::fma_c(u, c[0], k1, s.x);
			V::fma_c(u, c[1], k2, u);
			V::fma_c(X, c[2], k3, u);
#endif
			Sys::OdeFunction(gid, NTG, k4, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k4, dt, k4);

			T = s.t + dt;
			CkRow<P, CK_A51, 4>::load(c);
			V::fma_c(X, c[0], k1, s.x);
			V::fma_c(X, c[1], k2, X);
#if MPGOS_B200_TABLEAU_COMMON_FACTORS
			V::fma_c(u, P(-2.0), k3, k4);                 // a53 = -2 a54
			V::fma_c(X, c[3], u, X);
#else
			V::fma_c(X, c[2], k3, X);
			V::fma_c(X, c[3], k4, X);
#endif
			Sys::OdeFunction(gid, NTG, k5, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k5, dt, k5);

			T = s.t + dt * ck<P>(CK_C6);
			CkRow<P, CK_A61, 5>::load(c);
			V::fma_c(X, c[0], k1, s.x);
			V::fma_c(X, c[1], k2, X);
#if MPGOS_B200_TABLEAU_COMMON_FACTORS
			V::fma_c(u, P(77.0 / 8.0), k4, k3);           // a64 = (77/8) a63
			V::fma_c(X, c[2], u, X);
#else
			V::fma_c(X, c[2], k3, X);
			V::fma_c(X, c[3], k4, X);
#endif
			V::fma_c(X, c[4], k5, X);
			Sys::OdeFunction(gid, NTG, k6, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);
			V::mul_c(k6, dt, k6);

			CkRow<P, CK_B1, 4>::load(c);
			V::fma_c(xn, c[0], k1, s.x);
			V::fma_c(xn, c[1], k3, xn);
			V::fma_c(xn, c[2], k4, xn);
			V::fma_c(xn, c[3], k6, xn);
			CkRow<P, CK_E1, 6>::load(c);
#if MPGOS_B200_TABLEAU_COMMON_FACTORS
			V::fma_c(u, P(4.5), k5, k1);                  // e5 = 4.5 e1
			V::mul_c(u, c[0], u);
#else
			V::mul_c(u, c[0], k1);
			V::fma_c(u, c[3], k5, u);
#endif
			V::fma_c(u, c[1], k3, u);
			V::fma_c(u, c[2], k4, u);
			V::fma_c(u, c[4], k6, u);
			#pragma unroll
			for (int i = 0; i < SD; i++) err[i] = fabs(u[i]) + c[5];
#else
			// expression shapes as in the reference
			Sys::OdeFunction(gid, NTG, k1, s.x, s.t, s.cpar, p_par, p_pari, s.acc, s.acci);

			T = s.t + dt * ck<P>(CK_C2);
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (ck<P>(CK_A21) * k1[i]);
			Sys::OdeFunction(gid, NTG, k2, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);

			T = s.t + dt * ck<P>(CK_C3);
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (ck<P>(CK_A31) * k1[i] + ck<P>(CK_A32) * k2[i]);
			Sys::OdeFunction(gid, NTG, k3, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);

			T = s.t + dt * ck<P>(CK_C4);
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (ck<P>(CK_A41) * k1[i] + ck<P>(CK_A42) * k2[i] + ck<P>(CK_A43) * k3[i]);
			Sys::OdeFunction(gid, NTG, k4, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);

			T = s.t + dt;
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (ck<P>(CK_A51) * k1[i] + ck<P>(CK_A52) * k2[i] + ck<P>(CK_A53) * k3[i] + ck<P>(CK_A54) * k4[i]);
			Sys::OdeFunction(gid, NTG, k5, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);

			T = s.t + dt * ck<P>(CK_C6);
			#pragma unroll
			for (int i = 0; i < SD; i++) X[i] = s.x[i] + dt * (ck<P>(CK_A61) * k1[i] + ck<P>(CK_A62) * k2[i] + ck<P>(CK_A63) * k3[i] + ck<P>(CK_A64) * k4[i] + ck<P>(CK_A65) * k5[i]);
			Sys::OdeFunction(gid, NTG, k6, X, T, s.cpar, p_par, p_pari, s.acc, s.acci);

			#pragma unroll
			for (int i = 0; i < SD; i++)
			{
				xn[i] = s.x[i] + dt * (ck<P>(CK_B1) * k1[i] + ck<P>(CK_B3) * k3[i] + ck<P>(CK_B4) * k4[i] + ck<P>(CK_B6) * k6[i]);
				const P e = ck<P>(CK_E1) * k1[i] + ck<P>(CK_E3) * k3[i] + ck<P>(CK_E4) * k4[i] + ck<P>(CK_E5) * k5[i] + ck<P>(CK_E6) * k6[i];
				err[i] = dt * fabs(e) + ck<P>(CK_ERRFLOOR);
			}
#endif

			// ErrorControllers.cuh:24-86 -- accept test, limiting component, new step size.
			// The selects are ordered so that a NaN/Inf in xn or err reaches tol/err unchanged; the
			// range test below then routes the step to the exact path, which also sets `finite`.
			P tol[SD];
			#pragma unroll
			for (int i = 0; i < SD; i++)
			{
				// max(|x|, |xn|): select the signed value, take |.| as an operand modifier of the product (a materialised
				// fabs() is a DADD on the FP64 pipe)
				const P m  = (fabs(s.x[i]) > fabs(xn[i])) ? s.x[i] : xn[i];
				const P t  = a.RelativeTolerance[i] * fabs(m);
				tol[i] = (a.AbsoluteTolerance[i] > t) ? a.AbsoluteTolerance[i] : t;
			}
			// component with the smallest tol/err: tol_a/err_a < tol_b/err_b  <=>  tol_a err_b < tol_b err_a
			P tsel = tol[0], esel = err[0];
			#pragma unroll
			for (int i = 1; i < SD; i++)
			{
				const bool less = tol[i] * esel < tsel * err[i];
				tsel = less ? tol[i] : tsel;
				esel = less ? err[i] : esel;
			}
#if MPGOS_B200_ACCEPT_FROM_LIMITING
			// "every err_i < tol_i" is "the largest err_i/tol_i is below 1": one comparison on the limiting pair instead of one per
			// component.  The pair is found by cross-multiplication, whose two roundings can misorder components only when their
			// ratios agree to an ulp; the decision could then differ from the component-wise one only if, in addition, those ratios
			// straddle 1 within that ulp (err_i == tol_i to the last bit for one of them).  The parity build tests every component.
			update = esel < tsel;
#else
			#pragma unroll
			for (int i = 0; i < SD; i++) update = update && (err[i] < tol[i]);
#endif
			// Operand range of the fast path: every tol_i and err_i positive, finite and in [2^-100, 2^30) (in single precision
			// [2^-60, 2^30)), tested on the words that hold sign and exponent: the largest of them (as unsigned integers: a negative
			// number, an infinity or a NaN is larger than any admissible value) and the smallest tolerance (err_i carries the
			// 1e-18 floor and cannot be small).  3-input integer min/max: 4 instructions for 6 operands.
			unsigned top_max = 0u, top_min = 0xffffffffu;
			#pragma unroll
			for (int i = 0; i < SD; i++)
			{
				top_max = max(top_max, max(Num<P>::top(tol[i]), Num<P>::top(err[i])));
				top_min = min(top_min, Num<P>::top(tol[i]));
			}
			const bool fast = (top_max < Num<P>::kLow + Num<P>::kSpan) && (top_min >= Num<P>::kLow) &&
			                  ((unsigned)((int)(Num<P>::top(tsel) - Num<P>::top(esel)) + Num<P>::kRatioFloor) < (unsigned)(Num<P>::kRatioLimit + Num<P>::kRatioFloor));
			// ErrorControllers.cuh:60-85 -- forced acceptance at the minimum step, limits on the factor and on the new step.
			// None of them is active in an ordinary step, and whether one CAN be active is decided in single precision with a
			// margin (the args carry each limit moved inwards by 1e-4, the single-precision root is good to ~1e-6): five FP32
			// compares instead of six FP64 compares and eight 32-bit selects.  Any step near a limit, at the minimum step or
			// outside the fast path's operand range takes the exact code below, so the decisions themselves are unchanged.
			P mult;
			bool plain = false;
			if (fast)
			{
				float yf;
				mult = ck<P>(CK_SAFETY) * ratio_root_fast<P>(tsel, esel, update, yf);
				if (sizeof(P) == 8)
				{
					const float df = Num<P>::narrow(dt);
					const float mf = 0.9f * yf;
					const float nf = df * mf;
					plain = (df > a.GuardAbove) && (mf > a.ShrinkAbove) && (mf < a.GrowBelow) && (nf > a.MinimumAbove) && (nf < a.MaximumBelow);
					if (plain) s.dt_new = dt * mult;
				}
				else if (!(dt < a.MinimumTimeStepGuard))
				{
					// in FP32 the exact limits cost the same as the screening (measured: 0.896 vs 0.917 ms on C2 FP32): applied in place;
					// only the forced acceptance at the minimum step goes out of line
					mult = (mult > a.TimeStepGrowLimit) ? a.TimeStepGrowLimit : mult;
					mult = (mult < a.TimeStepShrinkLimit) ? a.TimeStepShrinkLimit : mult;
					P dt_new = dt * mult;
					dt_new = (dt_new > a.MaximumTimeStep) ? a.MaximumTimeStep : dt_new;
					dt_new = (dt_new < a.MinimumTimeStep) ? a.MinimumTimeStep : dt_new;
					s.dt_new = dt_new;
					plain = true;
				}
			}
			if (!plain)
			{   // [cold>
				// the rest of the controller is cold and lives out of line (controller_exact): the ordinary step neither holds its
				// temporaries in registers nor keeps the accept / terminate flags out of predicates because of it
				ControllerCold<P, SD> c;
				#pragma unroll
				for (int i = 0; i < SD; i++) { c.x[i] = s.x[i];  c.xn[i] = xn[i];  c.err[i] = err[i]; }
				c.dt = dt;  c.mult = mult;  c.fast = fast;  c.update = update;  c.gid = gid;
				const int r = controller_exact<P, SD, NE, NSP, NISP>(&a, &c);
				s.dt_new = c.dt_new;
				update = (r & 1) != 0;
				terminate = (r & 2) != 0;
			}   // <cold]
		}

		// ---- events: step-size control (EventHandling.cuh:6-45) ------------------------------
		// Both scans -- the secant step-size control here and the detection after the commit -- can only fire for an event whose
		// value changes sign over the step or ends inside the tolerance band (tolerances >= 0, checked on the host:
		// EventScreen).  That is one integer test and one comparison per event; a lane without such an event skips both scans
		// (a warp without such a lane jumps over them; the conditions inside are the reference's, unchanged).
		P evn[AtLeastOne<NE>::value];
		bool scan = true;
		if (NE > 0)
		{
			Sys::EventFunction(gid, NTG, evn, s.t + dt, dt, s.td, xn, s.cpar, p_par, p_pari, s.acc, s.acci);
			bool near = false;
			#pragma unroll
			for (int i = 0; i < NE; i++)
				near = near || ((int)(Num<P>::top(s.ev[i]) ^ Num<P>::top(evn[i])) < 0) || !(fabs(evn[i]) >= a.EventTolerance[i]);
			scan = (a.EventScreen == 0) || near;
			if (scan && update && !terminate)
			{   // [rare>
				P dt_event = dt;
				bool corrected = false;
				#pragma unroll
				for (int i = 0; i < NE; i++)
				{
					const P et = a.EventTolerance[i];
					if (((s.ev[i] > et) && (evn[i] < -et) && (a.EventDirection[i] <= 0)) ||
					    ((s.ev[i] < -et) && (evn[i] > et) && (a.EventDirection[i] >= 0)))
					{
						dt_event = fmin(dt_event, -s.ev[i] / (evn[i] - s.ev[i]) * dt);
						corrected = true;
					}
				}
				if (corrected)
				{
					if (dt_event < a.MinimumTimeStep)
					{
						if (a.DeviceMessages) device_message(1, gid);
					}
					else { s.dt_new = dt_event; update = false; }
				}
			}   // <rare]
		}

		// ---- accepted step (Solver.cuh:302-392) ------------------------------------------------
		if (update)
		{
			int user_terminate = 0;
			s.accepted++;
			s.t += dt;
			#pragma unroll
			for (int i = 0; i < SD; i++) s.x[i] = xn[i];

			Sys::ActionAfterSuccessfulTimeStep(gid, NTG, user_terminate, s.t, dt, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);

			if (NE > 0)
			{
				if (scan)
				{   // [rare>
					#pragma unroll
					for (int i = 0; i < NE; i++)
					{
						const P et = a.EventTolerance[i];
						if (((s.ev[i] > et) && (fabs(evn[i]) < et) && (a.EventDirection[i] <= 0)) ||
						    ((s.ev[i] < -et) && (fabs(evn[i]) < et) && (a.EventDirection[i] >= 0)))
							Sys::ActionAfterEventDetection(gid, NTG, i, user_terminate, s.t, dt, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);
					}
				}   // <rare]
				Sys::EventFunction(gid, NTG, s.ev, s.t, dt, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);
			}

			if (DENSE)   // DenseOutput.cuh:41-58 then :6-37
			{
				const bool room = s.dense_index < a.DenseCapacity;
				bool store = room && (s.dense_next_time < s.t) && (s.dense_skipped >= (a.DenseOutputSaveFrequency - 1));
				if (room && (end_reached || user_terminate == 1)) store = true;
				if (store)
				{
					a.DenseOutputTimeInstances[inst + (size_t)s.dense_index * stride] = s.t;
					#pragma unroll
					for (int i = 0; i < SD; i++)
						a.DenseOutputStates[inst + (size_t)i * stride + (size_t)s.dense_index * stride * SD] = s.x[i];
					s.dense_index++;
					s.dense_skipped = 0;
					s.dense_next_time = fmin(s.t + a.DenseOutputMinimumTimeStep, s.td[1]);
				}
				else s.dense_skipped++;
			}

			if (end_reached || user_terminate == 1) terminate = true;
		}
		return terminate;
	};

	// ---- end of the phase for this instance (Solver.cuh:398-435) ------------------------------------
	auto store = [&](const int inst, P& dt)
	{
		const int gid = s.gid;
		Sys::Finalization(gid, NTG, s.dense_index, s.t, dt, s.td, s.x, s.cpar, p_par, p_pari, s.acc, s.acci);
		#pragma unroll
		for (int i = 0; i < 2; i++)   a.TimeDomain[inst + i * stride] = s.td[i];
		#pragma unroll
		for (int i = 0; i < SD; i++)  a.ActualState[inst + i * stride] = s.x[i];
		#pragma unroll
		for (int i = 0; i < NCP; i++) a.ControlParameters[inst + i * stride] = s.cpar[i];
		#pragma unroll
		for (int i = 0; i < NA; i++)  a.Accessories[inst + i * stride] = s.acc[i];
		#pragma unroll
		for (int i = 0; i < NIA; i++) a.IntegerAccessories[inst + i * stride] = s.acci[i];
		a.ActualTime[inst] = s.t;
		if (DENSE) a.DenseOutputIndex[inst] = s.dense_index;
		a.StepsAttempted[inst] += s.attempted;
		a.StepsAccepted[inst]  += s.accepted;
		a.LastPhaseSteps[inst] = s.attempted;
	};

	if (!REFILL)
	{
		// direct mapping: thread i integrates instance i; the step loop is the whole kernel
		// blocks are dispatched in blockIdx order; with a cost history the launcher lists the tiles longest first
		const int tile = a.DispatchOrder ? a.DispatchOrder[blockIdx.x] : (int)blockIdx.x;
		const int inst = tile * blockDim.x + threadIdx.x;
		if (inst >= a.Active) return;
		load(inst);
		P dt;
		while (!step(inst, dt)) { }
		store(inst, dt);
		return;
	}

	// persistent warps
	int  inst  = 0;
	bool alive = false;
	bool queue_empty = false;       // warp-uniform
	for (;;)
	{
		bool take = false;
		// ---- warp ballot: refill the idle lanes together once there are enough of them -------
		const unsigned idle = __ballot_sync(0xffffffffu, !alive);
		if (idle != 0u)
		{
			const int n = __popc(idle);
			if (!queue_empty && (n >= a.RefillThreshold || n == 32))
			{
				const int lane = threadIdx.x & 31;
				const int leader = __ffs(idle) - 1;
				unsigned base = 0;
				if (lane == leader) base = atomicAdd(a.WorkQueue, (unsigned)n);
				base = __shfl_sync(0xffffffffu, base, leader);
				if (!alive)
				{
					inst = (int)(base + (unsigned)__popc(idle & ((1u << lane) - 1u)));
					take = inst < a.Active;
				}
				queue_empty = (base + (unsigned)n) >= (unsigned)a.Active;
			}
			if (idle == 0xffffffffu && !__any_sync(0xffffffffu, take)) break;
		}
		if (take) { load(inst); alive = true; }
		if (!alive) continue;
		P dt;
		if (step(inst, dt)) { store(inst, dt); alive = false; }
	}
}

#ifdef __CUDACC__      // from here on: kernel launches (the host-compiled test harness of tests/kernel_on_host stops here)
// ---------------------------------------------------------------------------------------------
// Dispatch order of the direct mapping.  The hardware hands out thread blocks in blockIdx order, so a launch ends with
// whatever its last blocks cost: when the expensive instances sit at the end of a sweep the SMs idle behind them (the
// "tailing effect"; Keller-Miksis frequency sweep, cost ratio 28 along the sweep: 37.5 ms with the expensive end first,
// 48.4 ms with it last; Lorenz: 4.66 vs 4.77 ms).  From the second phase on the previous phase's step counts say what
// each tile costs, and three small kernels list the tiles in order of decreasing cost (counting sort over 256
// logarithmic buckets, 8 per octave).  Results do not depend on the order: instances are independent.
static __global__ void dispatch_cost_histogram(const unsigned int* __restrict__ last, int active, int block, int tiles,
                                               unsigned int* __restrict__ bucket_of, unsigned int* __restrict__ hist)
{
	const int tile = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (tile >= tiles) return;
	unsigned cost = 0;
	for (int i = lane; i < block; i += 32)
	{
		const int inst = tile * block + i;
		if (inst < active) cost += last[inst];
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, o);
	if (lane == 0)
	{
		unsigned b = 0;
		if (cost)
		{
			const int e = 31 - __clz(cost);
			b = (unsigned)e * 8u + ((e >= 3 ? cost >> (e - 3) : cost << (3 - e)) & 7u);
		}
		bucket_of[tile] = b;
		atomicAdd(&hist[b], 1u);
	}
}

static __global__ void dispatch_bucket_starts(unsigned int* __restrict__ hist)      // hist[0..255] -> starts [256..511] (descending cost), fill [512..767] = 0
{
	__shared__ unsigned int h[256];
	h[threadIdx.x] = hist[threadIdx.x];
	__syncthreads();
	unsigned start = 0;
	for (int b = 255; b > (int)threadIdx.x; b--) start += h[b];
	hist[256 + threadIdx.x] = start;
	hist[512 + threadIdx.x] = 0;
}

static __global__ void dispatch_scatter(const unsigned int* __restrict__ bucket_of, unsigned int* __restrict__ hist, int tiles, int* __restrict__ order)
{
	const int tile = blockIdx.x * blockDim.x + threadIdx.x;
	if (tile >= tiles) return;
	const unsigned b = bucket_of[tile];
	order[hist[256 + b] + atomicAdd(&hist[512 + b], 1u)] = tile;
}

// ---------------------------------------------------------------------------------------------
// Host-side launcher generated next to the kernel.

// MINB = resident 256-thread blocks per SM the kernel is compiled for, i.e. its register cap: 1 -> 255, 2 -> 128,
// 3 -> 80, 4 -> 64 registers per thread.  Measured on B200 (profiles/): small systems (Lorenz, the impact
// oscillator) run fastest at 80 registers, a transcendental-heavy right-hand side (Keller-Miksis) at 128; the
// compiler's own choice without a cap (120 / 144) is slower in both cases.  Default 2; the built-in systems state
// theirs; a user translation unit may #define MPGOS_B200_MIN_BLOCKS_PER_SM before including the interface header.
template <class Sys, class P, int SD, int NCP, int NSP, int NISP, int NE, int NA, int NIA, bool DENSE, int ALG, int MINB = MPGOS_B200_MIN_BLOCKS_PER_SM>
struct PerThreadLauncher
{
	template <bool REFILL>
	static cudaError_t enqueue(const mpgos_perthread_launch* L, PerThreadArgs<P, SD, NE, NSP, NISP>& args, cudaStream_t stream)
	{
		auto kernel = perthread_kernel<Sys, P, SD, NCP, NSP, NISP, NE, NA, NIA, DENSE, ALG, REFILL, MINB>;
		// shared parameters that do not travel by value: dynamic shared memory (opt-in above 48 KB), or global memory when they
		// do not fit next to the launch's other needs or the user asked for it (PreferSharedMemory = 0, as in the reference)
		constexpr bool kSharedByValue = (NSP <= kSharedByValueLimit && NISP <= kSharedByValueLimit);
		size_t smem = kSharedByValue ? 0 : sizeof(P) * (size_t)NSP + sizeof(int) * (size_t)NISP;
		if (smem > 200 * 1024 || (!kSharedByValue && L->PreferSharedMemory == 0)) { smem = 0;  args.SharedInGlobal = 1; }
		if (smem > 48 * 1024)
		{
			cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			if (e != cudaSuccess) return e;
		}
		int block = L->ThreadsPerBlock;
		if (block < 32) block = 32;
		if (block > launch_bound_threads(MINB)) block = launch_bound_threads(MINB);
		block = (block / 32) * 32;
		const int active = args.Active;
		if (active <= 0) return cudaSuccess;
		// ThreadsPerBlock is a hint: a launch of many waves runs with blocks of at least 128 threads (the tutorials ask for 64, which
		// suits their 10^4 instances -- more blocks than SMs; measured on 2^18 Keller-Miksis instances: 38.2 ms with 64, 36.8 ms with
		// 128, Lorenz / impact unchanged: profiles/r2_variant_ab_runs.txt).  Per-instance arithmetic does not depend on the block shape.
		{
			const int sms = L->MultiProcessorCount > 0 ? L->MultiProcessorCount : 148;
			if (block < 128 && active >= 128 * 8 * sms) block = 128;      // at least two waves of 128-thread blocks at 4 resident blocks per SM
		}
		int grid = (active + block - 1) / block;
		if (REFILL)
		{
			// persistent grid: as many blocks as are resident at once, a multiple of the SM count
			int per_sm = 0;
			cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
			if (e != cudaSuccess) return e;
			if (per_sm < 1) per_sm = 1;
			const int resident = per_sm * (L->MultiProcessorCount > 0 ? L->MultiProcessorCount : 148);
			if (grid > resident) grid = resident;
			e = cudaMemsetAsync(L->d_WorkQueue, 0, sizeof(unsigned int), stream);
			if (e != cudaSuccess) return e;
		}
		else
		{
			// cost-ordered dispatch (see dispatch_cost_histogram): only when the launch spans several waves of blocks
			args.DispatchOrder = nullptr;
			const int sms = L->MultiProcessorCount > 0 ? L->MultiProcessorCount : 148;
			if (L->DispatchHistory && L->d_DispatchOrder && L->d_DispatchScratch && grid > 16 * sms)
			{
				unsigned int* bucket_of = L->d_DispatchScratch;
				unsigned int* hist = L->d_DispatchScratch + (L->desc.NT / 32 + 1);
				cudaError_t e = cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned int), stream);
				if (e != cudaSuccess) return e;
				dispatch_cost_histogram<<<(grid + 7) / 8, 256, 0, stream>>>(L->d_LastPhaseSteps, active, block, grid, bucket_of, hist);
				dispatch_bucket_starts<<<1, 256, 0, stream>>>(hist);
				dispatch_scatter<<<(grid + 255) / 256, 256, 0, stream>>>(bucket_of, hist, grid, L->d_DispatchOrder);
				args.DispatchOrder = L->d_DispatchOrder;
			}
		}
		kernel<<<grid, block, smem, stream>>>(args);
		return cudaGetLastError();
	}

	static int launch(const mpgos_perthread_launch* L, void* stream_v)
	{
		cudaStream_t stream = (cudaStream_t)stream_v;
		if (L->AbiVersion != MPGOS_B200_ABI_VERSION) return (int)cudaErrorInvalidValue;      // translation unit built against another mpgos_b200.h
		const mpgos_perthread_desc& d = L->desc;
		if (d.SD != SD || d.NCP != NCP || d.NSP != NSP || d.NISP != NISP || d.NE != NE || d.NA != NA || d.NIA != NIA ||
		    d.Algorithm != ALG || d.PrecisionBytes != (int)sizeof(P) || (d.NDO > 0) != DENSE)
			return (int)cudaErrorInvalidValue;

		PerThreadArgs<P, SD, NE, NSP, NISP> a;
		a.TimeDomain = (P*)L->d_TimeDomain;  a.ActualState = (P*)L->d_ActualState;  a.ActualTime = (P*)L->d_ActualTime;
		a.ControlParameters = (P*)L->d_ControlParameters;  a.SharedParameters = (P*)L->d_SharedParameters;
		a.IntegerSharedParameters = L->d_IntegerSharedParameters;  a.Accessories = (P*)L->d_Accessories;
		a.IntegerAccessories = L->d_IntegerAccessories;  a.DenseOutputIndex = L->d_DenseOutputIndex;
		a.DenseOutputTimeInstances = (P*)L->d_DenseOutputTimeInstances;  a.DenseOutputStates = (P*)L->d_DenseOutputStates;
		a.StepsAttempted = L->d_StepsAttempted;  a.StepsAccepted = L->d_StepsAccepted;  a.WorkQueue = L->d_WorkQueue;
		a.LastPhaseSteps = L->d_LastPhaseSteps;  a.DispatchOrder = nullptr;
		a.Stride = d.NT;
		a.Active = L->ActiveNumberOfThreads < d.NT ? L->ActiveNumberOfThreads : d.NT;
		a.InstanceOffset = d.InstanceOffset;
		a.ShardIndex = d.ShardIndex;  a.ShardCount = d.ShardCount;  a.ShardTile = d.ShardTile > 0 ? d.ShardTile : 1;
		a.GlobalInstances = d.GlobalInstances > 0 ? d.GlobalInstances : d.NT;
		a.DenseCapacity = d.NDO;
		a.DenseOutputSaveFrequency = L->DenseOutputSaveFrequency;
		a.DeviceMessages = L->DeviceMessages;
		a.RefillThreshold = L->LaneRefill < 1 ? 1 : (L->LaneRefill > 32 ? 32 : L->LaneRefill);
		a.InitialTimeStep = (P)L->InitialTimeStep;  a.MaximumTimeStep = (P)L->MaximumTimeStep;
		a.MinimumTimeStep = (P)L->MinimumTimeStep;  a.TimeStepGrowLimit = (P)L->TimeStepGrowLimit;
		a.TimeStepShrinkLimit = (P)L->TimeStepShrinkLimit;  a.DenseOutputMinimumTimeStep = (P)L->DenseOutputMinimumTimeStep;
		a.MinimumTimeStepGuard = a.MinimumTimeStep * P(1.01);
		a.GuardAbove   = (float)((double)a.MinimumTimeStepGuard * (1.0 + 1e-4));
		a.ShrinkAbove  = (float)((double)a.TimeStepShrinkLimit * (1.0 + 1e-4));
		a.GrowBelow    = (float)((double)a.TimeStepGrowLimit * (1.0 - 1e-4));
		a.MinimumAbove = (float)((double)a.MinimumTimeStep * (1.0 + 1e-4));
		a.MaximumBelow = (float)((double)a.MaximumTimeStep * (1.0 - 1e-4));
		for (int i = 0; i < SD; i++) { a.RelativeTolerance[i] = (P)L->RelativeTolerance[i];  a.AbsoluteTolerance[i] = (P)L->AbsoluteTolerance[i]; }
		for (int i = 0; i < AtLeastOne<NE>::value; i++)
		{
			a.EventTolerance[i] = (i < NE) ? (P)L->EventTolerance[i] : P(0);
			a.EventDirection[i] = (i < NE) ? L->EventDirection[i] : 0;
		}
		a.EventScreen = 1;
		for (int i = 0; i < NE; i++) if (!(a.EventTolerance[i] >= P(0))) a.EventScreen = 0;
		constexpr bool kSharedByValue = (NSP <= kSharedByValueLimit && NISP <= kSharedByValueLimit);
		a.SharedValues[0] = P(0);  a.IntegerSharedValues[0] = 0;  a.SharedInGlobal = 0;
		if (kSharedByValue)
		{
			for (int i = 0; i < NSP; i++)  a.SharedValues[i] = (P)L->SharedParameterValues[i];
			for (int i = 0; i < NISP; i++) a.IntegerSharedValues[i] = L->IntegerSharedParameterValues[i];
		}
		return (int)(L->LaneRefill ? enqueue<true>(L, a, stream) : enqueue<false>(L, a, stream));
	}
};
#endif // __CUDACC__

} // namespace mpgos_b200
#endif
