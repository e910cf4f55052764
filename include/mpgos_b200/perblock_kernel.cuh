// perblock_kernel.cuh -- block-per-system (coupled systems) ensemble integrator for sm_100a (B200).
//
// One thread integrates one UNIT, the units of one system occupy whole warps (a "slot"), several slots share a thread
// block.  Same algorithm and decision semantics as the reference's coupled module -- the four kernels of
// SourceCodes/CoupledSystmes_PerBlock_Solver.cuh:11, 787, 1508, 2215 with CoupledSystems_PerBlock_ExplicitRungeKutta_Steppers.cuh,
// ..._ErrorControllers.cuh (safety 0.8), ..._EventHandling.cuh, ..._DenseOutput.cuh -- on a different machine mapping:
//
//  * a slot synchronises with its OWN named barrier (bar.sync id, count) or __syncwarp() when a system fits one warp,
//    never with __syncthreads(): systems in a block advance independently (the reference's multi-system kernels step all
//    systems of a block in lockstep behind ~25 block-wide barriers per RKCK45 step; here 9 slot-wide ones);
//  * the per-system reductions (accept flag, limiting error ratio, earliest event) are warp shuffles plus one
//    shared-memory word per warp, not shared-memory atomics (atomicAnd / CAS-loop atomicFMIN in the reference);
//  * every thread evaluates the step-size controller redundantly from the reduced values (no thread-0 serial section,
//    no broadcast barrier); its pow/division are replaced by the division-free Newton root of perthread_kernel.cuh;
//  * the coupling matrices live in shared memory once per block, coupling terms are double-buffered (one barrier per
//    stage instead of two); the column sum of a dense matrix runs over MPGOS_B200_COUPLING_ACCUMULATORS (default 4)
//    interleaved partial sums -- same terms as ComputeCouplingValue_* (Steppers.cuh:2551-2713), different summation order,
//    results track the reference to rounding; 1 accumulator reproduces the reference's ascending single chain bit for bit;
//  * slot refill: a slot that finishes its system takes the next system id from a device-side queue (persistent grid);
//  * systems of more than 1024 units (the reference's ..._MultipleBlockLaunches kernels): UPT units per thread, thread t of the slot owns
//    units t, t + TPS, t + 2 TPS, ... in [UPT][...] register arrays; their coupling terms live in dynamic shared memory (up to 4096 units);
//  * TENSOR-CORE COUPLING (MMA variant; dense matrices shared by all systems, FP64, UPS = 32, 64 or 128): the column sums of the
//    8 systems of a block (4 for UPS = 128) are ONE matrix product  V[UPS x 8] = M[UPS x UPS] . CPT[UPS x 8]  per stage, issued as
//    mma.sync.m8n8k4.f64 (DMMA) by all warps of the block; the matrix sits in shared memory in fragment order.  The scalar
//    form needs one 8-byte shared-memory operand per FMA and is bound by shared-memory bandwidth (ncu: 45 % of the LSU
//    wavefront peak at 44 % FP64-pipe activity, profiles/); one DMMA performs 256 FMAs on two 8-byte operands per lane.
//    The 8 systems then advance in lockstep through the stages (two block barriers per stage); everything else stays
//    per slot.  The FP32 instantiation does the same product with mma.m16n8k8 on TF32 operands split into big + small parts
//    (3xTF32: FP32 accuracy from 10-bit operands), UPS = 32 or 64.
#ifndef MPGOS_B200_PERBLOCK_KERNEL_CUH
#define MPGOS_B200_PERBLOCK_KERNEL_CUH

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include "../mpgos_b200_coupled.h"
#include "perthread_kernel.cuh"     // AtLeastOne, Num<P>, ck<P>(), sfu_*, Cash-Karp table

#ifndef MPGOS_B200_COUPLING_ACCUMULATORS
#define MPGOS_B200_COUPLING_ACCUMULATORS 4
#endif
#ifndef MPGOS_B200_COUPLED_THREADS
#define MPGOS_B200_COUPLED_THREADS 512
#endif

namespace mpgos_b200 {

template <class P, int UD, int NE>
struct PerBlockArgs
{
	P* TimeDomain;  P* ActualState;  P* ActualTime;  P* UnitParameters;  P* SystemParameters;  P* GlobalParameters;  int* IntegerGlobalParameters;
	P* UnitAccessories;  int* IntegerUnitAccessories;  P* SystemAccessories;  int* IntegerSystemAccessories;
	P* CouplingMatrix;  P* CouplingStrength;  int* CouplingIndex;
	int* DenseOutputIndex;  P* DenseOutputTimeInstances;  P* DenseOutputStates;
	unsigned long long* StepsAttempted;  unsigned long long* StepsAccepted;  unsigned int* WorkQueue;
	unsigned int* LastPhaseSteps;  const int* DispatchOrder;      // cost history and queue order (perthread_kernel.cuh, dispatch_cost_histogram)
	// time slices inside one phase (first phase of a sweep: no cost history yet).  SliceMode 1: the pilot -- every system is paused on
	// its first accepted step at or past t0 + SliceFraction (TD1 - t0), nothing is clipped, so the step sequence is the one of the
	// unsliced phase; its controller state goes to the Slice* arrays.  SliceMode 2: the continuation, in the order the pilot's step
	// counts give; systems that ended inside the pilot are skipped.  0: the whole phase in one launch.
	int SliceMode;  P SliceFraction;
	int MatrixInGlobal;      // scalar coupling only: the packed matrices do not fit into shared memory and are read from global memory
	P* SliceDt;  P* SliceLastDt;  P* SliceDenseNext;  int* SliceSkipped;  int* SlicePaused;
	// what a callback may have written into the parameter arrays before the pause: unit parameters, system parameters and coupling
	// strengths live in registers / shared memory for the whole phase in the reference, so the continuation must see them as they
	// were at the pause, not as they are in global memory
	P* SliceUnitParameters;  P* SliceSystemParameters;  P* SliceCouplingStrength;
	int ShardIndex, ShardCount, ShardTile;      // tile-cyclic shard of a multi-device solver (mpgos_b200_coupled.h); ShardCount <= 1: none
	int NS, Active, SystemOffset, DenseCapacity, DenseOutputSaveFrequency, DeviceMessages;
	P   InitialTimeStep, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit, DenseOutputMinimumTimeStep, MinimumTimeStepGuard;
	P   RelativeTolerance[UD];  P AbsoluteTolerance[UD];
	P   EventTolerance[AtLeastOne<NE>::value];  int EventDirection[AtLeastOne<NE>::value];
};

// resident blocks per SM the kernel is compiled for: as many as keep 512 threads (128 registers each) on the SM
// (MPGOS_B200_COUPLED_THREADS_PER_SM: A/B builds with 768 threads at 80 registers or 1024 at 64)
#ifndef MPGOS_B200_COUPLED_THREADS_PER_SM
#define MPGOS_B200_COUPLED_THREADS_PER_SM 512
#endif
__host__ __device__ constexpr int coupled_min_blocks(int threads)
{
	return threads >= MPGOS_B200_COUPLED_THREADS_PER_SM ? 1 : (MPGOS_B200_COUPLED_THREADS_PER_SM / threads > 4 ? 4 : MPGOS_B200_COUPLED_THREADS_PER_SM / threads);
}

__host__ __device__ constexpr int packed_matrix_elements(int UPS, int NC, int CBW, int CCI)
{
	return (CCI == 0 && CBW == 0) ? NC * UPS * UPS : (CCI == 0) ? NC * UPS * (2 * CBW + 1) : (CBW > 0) ? NC * (2 * CBW + 1) : NC * UPS;
}

// y = a^(-1/n), n = 5 (accept) or 4, for a = max_i err_i/tol_i of a whole system: the same division-free Newton map
// as ratio_root_fast with tol = 1, err = a.
template <class P>
__device__ __forceinline__ P inverse_root_fast(P a, bool accept)
{
	const float inv = accept ? 0.2f : 0.25f;
	P y = (P)sfu_ex2(-sfu_lg2(Num<P>::narrow(a)) * inv);
	const P rn = accept ? P(0.2) : P(0.25);
	#pragma unroll
	for (int it = 0; it < (sizeof(P) == 8 ? 2 : 1); it++)
	{
		const P y2 = y * y, y4 = y2 * y2;
		const P yn = accept ? y4 * y : y4;
		y = fma(y * fma(-a, yn, P(1.0)), rn, y);
	}
	return y;
}

template <class P, int UPS, int UPSP, int NC, int NSP, int NSA, int NiSA, bool MMA>
struct SlotShared
{
	static constexpr int WPS = UPSP / 32;
	alignas(16) P cpt[2][AtLeastOne<NC>::value][MMA ? 2 : UPSP];   // coupling terms, double buffered over stages (scalar variant)
	P   t, dt, td[2];
	P   spar[AtLeastOne<NSP>::value], sacc[AtLeastOne<NSA>::value], strength[AtLeastOne<NC>::value];
	P   dense_next_time;
	P   slice_time;          // SliceMode 1: pause at the first accepted step at or past this time
	P   red_ratio[WPS];      // per-warp max of err/tol
	P   red_evdt[WPS];       // per-warp min of the secant event step
	int red_flags[WPS];      // bit 0: all accepted, bit 1: all finite
	int red_corrected[WPS];
	int isacc[AtLeastOne<NiSA>::value];
	int dense_index, dense_skipped, udt, next_system;
};

template <int UPSP>
__device__ __forceinline__ void slot_barrier(int slot)
{
	if (UPSP == 32) __syncwarp();
	else named_barrier_sync(slot + 1, UPSP);
}

// Steppers.cuh:2551-2713 on the packed matrix; `cpt` is one coupling term of all units of the system.
template <class P, int UPS, int CBW, int CCI>
__device__ __forceinline__ P coupling_value(const P* __restrict__ M, const P* cpt, int unit)
{
	P v = 0;
	if (CCI == 0 && CBW == 0)
	{
		// MPGOS_B200_COUPLING_ACCUMULATORS independent partial sums over interleaved columns break the UPS-long chain of
		// dependent FMAs (8 cycles each); 1 reproduces the reference's summation order exactly.
		constexpr int NA = (UPS % MPGOS_B200_COUPLING_ACCUMULATORS == 0) ? MPGOS_B200_COUPLING_ACCUMULATORS : 1;
		P acc[NA];
		#pragma unroll
		for (int j = 0; j < NA; j++) acc[j] = 0;
		#pragma unroll 4
		for (int col = 0; col < UPS; col += NA)
		{
			#pragma unroll
			for (int j = 0; j < NA; j++) acc[j] += M[unit + (col + j) * UPS] * cpt[col + j];
		}
		#pragma unroll
		for (int j = 0; j < NA; j++) v += acc[j];
	}
	else
	{
		constexpr int ND = (CBW > 0) ? 2 * CBW + 1 : UPS;
		#pragma unroll 8
		for (int d = 0; d < ND; d++)
		{
			int s = unit + d - CBW;
			if (s >= UPS) s -= UPS;
			if (s < 0) s += UPS;
			v += ((CCI == 0) ? M[unit + d * UPS] : M[d]) * cpt[s];
		}
	}
	return v;
}

// ---- tensor-core coupling -------------------------------------------------------------------------------------------
// D[8x8] += A[8x4] . B[4x8] in FP64.  Fragments (PTX ISA, mma.m8n8k4 .f64): lane l holds A[l/4][l%4], B[l%4][l/4] and
// D[l/4][2*(l%4) + {0,1}].

// Geometry of the block-wide product for UPS units and SLOTS <= 8 systems (warp-groups of UPS threads; the N extent of the
// instruction is 8, columns SLOTS..7 stay zero).  The RT x KG work items (one row tile x one group of k tiles) are dealt
// round-robin to the warps of the block.
template <int UPS, int SLOTS_>
struct MmaGeometry
{
	static constexpr int SLOTS = SLOTS_;
	static constexpr int NCOL = 8;                     // the N extent of the instruction
	static constexpr int RT = UPS / 8;                 // row tiles of the matrix
	static constexpr int KT = UPS / 4;                 // column (k) tiles
	static constexpr int WARPS = SLOTS * UPS / 32;
	static constexpr int KG = (WARPS >= RT) ? WARPS / RT : 1;      // a warp owns one row tile and 1/KG of the k tiles (or several row tiles)
	static constexpr int ITEMS = RT * KG;
	static constexpr int KPG = KT / KG;
	static constexpr int CS = UPS + 4;                 // stride of a system's coupling terms  (conflict-free B fragment loads)
	static constexpr int VS = UPS + 2;                 // stride of a system's partial sums    (conflict-free D fragment stores)
	static constexpr int MCOPIES = 1, CCOPIES = 1;
	static_assert(UPS % 32 == 0 && SLOTS >= 1 && SLOTS <= NCOL && ITEMS % WARPS == 0 && KT % KG == 0 && KPG % 2 == 0, "unsupported shape for the tensor-core coupling");
};

// One coupling term of all systems of the block.  frag: the matrix in fragment order [RT][KT][32]; cpt: [8][CS]; vpart: [KG][8][VS].
template <int UPS, int SLOTS>
__device__ __forceinline__ void mma_coupling_product(const double* __restrict__ frag, const double* cpt, double* vpart, int warp, int lane)
{
	using G = MmaGeometry<UPS, SLOTS>;
	#pragma unroll
	for (int item = warp; item < G::ITEMS; item += G::WARPS)
	{
		const int rt = item % G::RT, kg = item / G::RT;
		const double* A = frag + (size_t)(rt * G::KT + kg * G::KPG) * 32 + lane;
		const double* B = cpt + (lane >> 2) * G::CS + kg * G::KPG * 4 + (lane & 3);
		double c0 = 0, c1 = 0, e0 = 0, e1 = 0;             // two independent accumulator pairs hide the DMMA latency (four: measured slower, 13.86 vs 13.61 ms on C5)
		#pragma unroll
		for (int j = 0; j < G::KPG; j += 2)
		{
			const double a0 = A[j * 32], b0 = B[j * 4], a1 = A[(j + 1) * 32], b1 = B[(j + 1) * 4];
			dmma_8x8x4(c0, c1, a0, b0);
			dmma_8x8x4(e0, e1, a1, b1);
		}
		double* V = vpart + (size_t)kg * G::NCOL * G::VS + rt * 8 + (lane >> 2);
		V[(2 * (lane & 3)) * G::VS]     = c0 + e0;
		V[(2 * (lane & 3) + 1) * G::VS] = c1 + e1;
	}
}

// ---- tensor-core coupling in single precision: 3xTF32 ------------------------------------------------------------------
// D[16x8] += A[16x8] . B[8x8] with TF32 operands and FP32 accumulation (mma.m16n8k8).  A TF32 operand keeps 10 mantissa bits, so
// every FP32 value is split into big = tf32(v) and small = tf32(v - big) and the product is formed as
// A_small.B_big + A_big.B_small + A_big.B_big (the small.small term is below FP32 resolution): ~2e-7 relative, FP32 accuracy.
// Fragments (PTX ISA, mma.m16n8k8 .tf32), g = lane/4, t = lane%4: A: a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4);
// B: b0 (k = t, n = g), b1 (k = t+4, n = g); D: d0 (g, 2t), d1 (g, 2t+1), d2 (g+8, 2t), d3 (g+8, 2t+1).
__device__ __forceinline__ void tf32_split(float v, float& big, float& small)
{
	big = __uint_as_float(to_tf32(v));
	small = __uint_as_float(to_tf32(v - big));
}

template <int UPS, int SLOTS_>
struct MmaGeometryF32
{
	static constexpr int SLOTS = SLOTS_;
	static constexpr int NCOL = 8;
	static constexpr int RT = UPS / 16;                // row tiles
	static constexpr int KT = UPS / 8;                 // k tiles
	static constexpr int WARPS = SLOTS * UPS / 32;
	static constexpr int KG = (WARPS >= RT) ? WARPS / RT : 1;
	static constexpr int ITEMS = RT * KG;
	static constexpr int KPG = KT / KG;
	static constexpr int CS = UPS + 4;                 // conflict-free B fragment loads (bank = 4 g + t)
	static constexpr int VS = UPS + 4;                 // conflict-free D fragment stores (bank = 8 t + g)
	static constexpr int MCOPIES = 2, CCOPIES = 2;     // big and small parts
	static_assert(UPS % 32 == 0 && SLOTS >= 1 && SLOTS <= NCOL && ITEMS % WARPS == 0 && KT % KG == 0, "unsupported shape for the TF32 tensor-core coupling");
};

// frag: the matrix in fragment order [RT][KT][32] float4, big part then small part (`small_offset` float4s further);
// cpt: [8][CS] big, then small (`cpt_small` floats further); vpart: [KG][8][VS].
template <int UPS, int SLOTS>
__device__ __forceinline__ void mma_coupling_product_f32(const float4* __restrict__ frag, int small_offset, const float* cpt, int cpt_small, float* vpart, int warp, int lane)
{
	using G = MmaGeometryF32<UPS, SLOTS>;
	const int g = lane >> 2, t = lane & 3;
	#pragma unroll
	for (int item = warp; item < G::ITEMS; item += G::WARPS)
	{
		const int rt = item % G::RT, kg = item / G::RT;
		float d[4] = {0.f, 0.f, 0.f, 0.f};
		#pragma unroll
		for (int j = 0; j < G::KPG; j++)
		{
			const int kt = kg * G::KPG + j;
			const float4 ab = frag[(rt * G::KT + kt) * 32 + lane];
			const float4 as = frag[(rt * G::KT + kt) * 32 + lane + small_offset];
			const float* B = cpt + g * G::CS + kt * 8 + t;
			const float bb0 = B[0], bb1 = B[4], bs0 = B[cpt_small], bs1 = B[cpt_small + 4];
			mma_16x8x8_tf32(d, as, bb0, bb1);
			mma_16x8x8_tf32(d, ab, bs0, bs1);
			mma_16x8x8_tf32(d, ab, bb0, bb1);
		}
		float* V = vpart + (size_t)kg * G::NCOL * G::VS + rt * 16 + g;
		V[(2 * t) * G::VS]         = d[0];
		V[(2 * t + 1) * G::VS]     = d[1];
		V[(2 * t) * G::VS + 8]     = d[2];
		V[(2 * t + 1) * G::VS + 8] = d[3];
	}
}

template <class P, int UPS, int SLOTS> struct MmaGeometryFor { using type = MmaGeometry<UPS, SLOTS>; };
template <int UPS, int SLOTS> struct MmaGeometryFor<float, UPS, SLOTS> { using type = MmaGeometryF32<UPS, SLOTS>; };

// threads of a slot when every thread integrates UPT units (UPT > 1: systems of more than 1024 units; the reference's "multiple block
// launches", Solver.cuh:1508, 2215): thread t of the slot owns units t, t + TPS, t + 2 TPS, ...
__host__ __device__ constexpr int slot_threads(int UPS, int UPT) { return (((UPS + UPT - 1) / UPT + 31) / 32) * 32; }

template <class Sys, class P, int UPS, int UD, int NC, int CBW, int CCI, int NUP, int NSP, int NGP, int NiGP, int NUA, int NiUA, int NSA, int NiSA,
          int NE, bool DENSE, int ALG, int SPBK, bool MMA, int UPT = 1>
__global__ void __launch_bounds__(SPBK * slot_threads(UPS, UPT), coupled_min_blocks(SPBK * slot_threads(UPS, UPT)))
perblock_kernel(const __grid_constant__ PerBlockArgs<P, UD, NE> a)
{
	constexpr int UPSP = ((UPS + 31) / 32) * 32;      // units of a system, padded to whole warps (shared arrays)
	constexpr int TPS = slot_threads(UPS, UPT);       // threads of a slot (UPT == 1: UPSP)
	constexpr int WPS = TPS / 32;
	static_assert(UPT >= 1 && (UPT == 1 || !MMA), "several units per thread: scalar coupling only");
	constexpr int MSIZE = packed_matrix_elements(UPS, NC, CBW, CCI);
	constexpr int MSTRIDE = MSIZE / (NC > 0 ? NC : 1);
	// coupling terms of the slot's units: inside the slot's static shared struct, or (several units per thread: up to 4096 units) in
	// dynamic shared memory behind the global parameters
	constexpr bool BIG = UPT > 1;
	using Slot = SlotShared<P, UPS, UPSP, NC, NSP, NSA, NiSA, MMA || BIG>;
	static_assert(sizeof(Slot) * SPBK <= 47 * 1024, "the coupling terms of the block's systems do not fit into static shared memory");
	using G = typename MmaGeometryFor<P, MMA ? UPS : 32, MMA ? SPBK : 8>::type;
	constexpr int MCOPIES   = MMA ? G::MCOPIES : 1;                        // FP32: big and small TF32 parts of the matrix
	constexpr int CPT_ELEMS = MMA ? NC * G::NCOL * G::CS : 0;              // coupling terms of the systems of the block (MMA variant)
	constexpr int CPT_ALL   = MMA ? CPT_ELEMS * G::CCOPIES : 0;
	constexpr int V_ELEMS   = MMA ? NC * G::KG * G::NCOL * G::VS : 0;      // partial column sums per k group
	static_assert(!MMA || (CBW == 0 && CCI == 0 && NC > 0), "tensor-core coupling: dense matrices shared by all systems");

	MPGOS_B200_DYNAMIC_SHARED_MEMORY(dyn_smem);
	P*   s_matrix = reinterpret_cast<P*>(dyn_smem);                      // [MSIZE x MCOPIES]  packed (scalar) or in fragment order (MMA)
	P*   s_cptm   = s_matrix + ((!MMA && a.MatrixInGlobal) ? 0 : MSIZE * MCOPIES);      // [CPT_ALL]
	P*   s_vpart  = s_cptm + CPT_ALL;                                   // [V_ELEMS]
	P*   s_gpar   = s_vpart + V_ELEMS;                                  // [NGP]
	int* s_igpar  = reinterpret_cast<int*>(s_gpar + NGP);               // [NiGP]
	P*   s_cptbig = reinterpret_cast<P*>((reinterpret_cast<size_t>(s_igpar + NiGP) + 15) & ~(size_t)15);      // [SPBK][2][NC][UPSP] when BIG (16-byte aligned)
	__shared__ Slot s_slots[SPBK];
	__shared__ int  s_cidx[AtLeastOne<NC>::value];

	if (MMA)
	{
		if constexpr (sizeof(P) == 8)
		{
			// fragment order: tile (rt, kt) = rows 8 rt .. 8 rt + 7, columns 4 kt .. 4 kt + 3; lane l holds M[8 rt + l/4][4 kt + l%4]
			for (int i = threadIdx.x; i < MSIZE; i += blockDim.x)
			{
				const int k = i / (UPS * UPS), r = i % (UPS * UPS), tile = r / 32, l = r % 32;
				const int row = (tile / G::KT) * 8 + (l >> 2), col = (tile % G::KT) * 4 + (l & 3);
				s_matrix[i] = a.CouplingMatrix[(size_t)k * UPS * UPS + row + col * UPS];
			}
		}
		else
		{
			// fragment order of m16n8k8: tile (rt, kt) = rows 16 rt .., columns 8 kt ..; lane l holds the float4 (a0, a1, a2, a3); big
			// parts first, small parts MSIZE elements further
			for (int i = threadIdx.x; i < MSIZE; i += blockDim.x)
			{
				const int k = i / (UPS * UPS), r = i % (UPS * UPS), tile = r / 128, w = r % 128, l = w >> 2, reg = w & 3;
				const int row = (tile / G::KT) * 16 + (l >> 2) + ((reg & 1) ? 8 : 0), col = (tile % G::KT) * 8 + (l & 3) + ((reg & 2) ? 4 : 0);
				float big, small;
				tf32_split((float)a.CouplingMatrix[(size_t)k * UPS * UPS + row + col * UPS], big, small);
				s_matrix[i] = (P)big;  s_matrix[i + MSIZE] = (P)small;
			}
		}
		for (int i = threadIdx.x; i < CPT_ALL; i += blockDim.x) s_cptm[i] = P(0);      // columns without a slot stay zero
	}
	else if (!a.MatrixInGlobal)
		for (int i = threadIdx.x; i < MSIZE; i += blockDim.x) s_matrix[i] = a.CouplingMatrix[i];
	// matrices too large for shared memory (dense FP64 from ~170 units on, or several couplings of 128 units) stay in global memory,
	// where the reference keeps them by default (SharedCouplingMatrices = 0); they are shared by all systems, so L1 / L2 serve them
	const P* const matrix = (!MMA && a.MatrixInGlobal) ? a.CouplingMatrix : s_matrix;
	for (int i = threadIdx.x; i < NGP; i += blockDim.x)   s_gpar[i] = a.GlobalParameters[i];
	for (int i = threadIdx.x; i < NiGP; i += blockDim.x)  s_igpar[i] = a.IntegerGlobalParameters[i];
	for (int i = threadIdx.x; i < NC; i += blockDim.x)    s_cidx[i] = a.CouplingIndex[i];
	__syncthreads();

	const int  slot = threadIdx.x / TPS;
	const int  unit = threadIdx.x % TPS;     // the thread's first unit (its only one when UPT == 1); unit 0's thread does the system-scope work
	const bool is_unit = unit < UPS;         // padding lanes only take part in barriers
	// u-th unit of this thread, and whether it exists
	auto unit_of = [&](int u) { return unit + u * TPS; };
	auto is_unit_of = [&](int u) { return unit + u * TPS < UPS; };
	const int  wis  = unit / 32;             // warp within the slot
	const int  lane = unit & 31;
	const int  warp = threadIdx.x / 32;
	Slot& S = s_slots[slot];
	const size_t NU = (size_t)a.NS * UPS;    // unit-scope stride

	// unit-scope data of the slot's current system: registers
	P   x[UPT][UD], upar[UPT][AtLeastOne<NUP>::value], uacc[UPT][AtLeastOne<NUA>::value], ev[UPT][AtLeastOne<NE>::value];
	int iuacc[UPT][AtLeastOne<NiUA>::value];
	P   dt_new = a.InitialTimeStep;
	unsigned attempted = 0, accepted = 0;
	int    pos = blockIdx.x * SPBK + slot;      // position in the system queue
	int    sys = 0, gsys = 0;
	auto ug_of = [&](int u) { return (size_t)sys * UPS + (is_unit_of(u) ? unit_of(u) : 0); };      // unit-scope index of the thread's u-th unit
	bool   have = false, dry = false, first = true, resumed = false;
	int    buf = 0;

	// One trip of this loop = one attempted step of the slot's system.  Scalar variant: slots are independent.  MMA variant:
	// the 8 slots of the block pass through the stages together (a slot without a system only takes part in the barriers
	// and in the matrix product) until no slot has a system left.
	for (;;)
	{
		// ---- take a system ------------------------------------------------------------------------------
		if (!have && !dry)
		{
			if (!first)
			{
				if (unit == 0) S.next_system = (int)(gridDim.x * SPBK + atomicAdd(a.WorkQueue, 1u));
				slot_barrier<TPS>(slot);
				pos = S.next_system;
			}
			first = false;
			if (pos >= a.Active) dry = true;
			else
			{
				sys = a.DispatchOrder ? a.DispatchOrder[pos] : pos;      // with a cost history: the longest systems first
				resumed = (a.SliceMode == 2);
			}
			if (!dry && !(resumed && a.SlicePaused[sys] == 0))      // continuation: a system that ended inside the pilot is skipped
			{
				gsys = (a.ShardCount > 1 ? ((sys / a.ShardTile) * a.ShardCount + a.ShardIndex) * a.ShardTile + sys % a.ShardTile : sys) + a.SystemOffset;
				// ---- load: system scope -> shared, unit scope -> registers (Solver.cuh:2279-2435) ---------------
				if (unit == 0)
				{
					S.td[0] = a.TimeDomain[sys];  S.td[1] = a.TimeDomain[sys + a.NS];
					S.t = a.ActualTime[sys];  S.dt = resumed ? a.SliceLastDt[sys] : a.InitialTimeStep;
					S.dense_index = DENSE ? a.DenseOutputIndex[sys] : 0;  S.udt = 0;
					S.dense_skipped = resumed ? a.SliceSkipped[sys] : 0;  S.dense_next_time = resumed ? a.SliceDenseNext[sys] : P(0);
				}
				if (unit < NSP)  S.spar[unit]     = (resumed ? a.SliceSystemParameters : a.SystemParameters)[sys + unit * a.NS];
				if (unit < NSA)  S.sacc[unit]     = a.SystemAccessories[sys + unit * a.NS];
				if (unit < NiSA) S.isacc[unit]    = a.IntegerSystemAccessories[sys + unit * a.NS];
				if (unit < NC)   S.strength[unit] = (resumed ? a.SliceCouplingStrength : a.CouplingStrength)[sys + unit * a.NS];
				#pragma unroll
				for (int u = 0; u < UPT; u++)
				{
					const size_t g = ug_of(u);
					#pragma unroll
					for (int i = 0; i < UD; i++)   x[u][i]     = a.ActualState[g + i * NU];
					#pragma unroll
					for (int i = 0; i < NUP; i++)  upar[u][i]  = (resumed ? a.SliceUnitParameters : a.UnitParameters)[g + i * NU];
					#pragma unroll
					for (int i = 0; i < NUA; i++)  uacc[u][i]  = a.UnitAccessories[g + i * NU];
					#pragma unroll
					for (int i = 0; i < NiUA; i++) iuacc[u][i] = a.IntegerUnitAccessories[g + i * NU];
				}
				slot_barrier<TPS>(slot);

				if (!resumed)
				{
					#pragma unroll
					for (int u = 0; u < UPT; u++)
						if (is_unit_of(u)) Sys::Initialization(gsys, unit_of(u), S.dense_index, S.t, S.dt, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
				}
				slot_barrier<TPS>(slot);
				if (a.SliceMode == 1)
				{
					if (unit == 0) S.slice_time = S.t + a.SliceFraction * (S.td[1] - S.t);
					slot_barrier<TPS>(slot);
				}
				if (NE > 0)
				{
					#pragma unroll
					for (int u = 0; u < UPT; u++)
						if (is_unit_of(u)) Sys::EventFunction(gsys, unit_of(u), ev[u], S.t, S.dt, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
					slot_barrier<TPS>(slot);
				}
				if (DENSE && !resumed)   // first point of the phase (Solver.cuh:2480-2497), stored without looking at the index
				{
					const int di = S.dense_index;
					if (di < a.DenseCapacity)
					{
						#pragma unroll
						for (int u = 0; u < UPT; u++)
							if (is_unit_of(u))
							{
								#pragma unroll
								for (int i = 0; i < UD; i++) a.DenseOutputStates[ug_of(u) + i * NU + (size_t)di * NU * UD] = x[u][i];
							}
						if (unit == 0) a.DenseOutputTimeInstances[sys + (size_t)di * a.NS] = S.t;
					}
					slot_barrier<TPS>(slot);
					if (unit == 0) { S.dense_index = di + 1;  S.dense_skipped = 0;  S.dense_next_time = fmin(S.t + a.DenseOutputMinimumTimeStep, S.td[1]); }
					slot_barrier<TPS>(slot);
				}
				dt_new = resumed ? a.SliceDt[sys] : a.InitialTimeStep;
				attempted = 0;  accepted = 0;
				have = true;
			}
			else if (!dry) slot_barrier<TPS>(slot);      // skipped: everybody has read the queue position before unit 0 asks for the next one
		}
		// leave when no slot has a system and none can get one any more (a slot that skipped a system takes another next trip)
		if (MMA) { if (!__syncthreads_or((have || !dry) ? 1 : 0)) break; }
		else if (!have) { if (dry) break; else continue; }
		const bool live = have && is_unit;      // the thread's first unit
		auto live_of = [&](int u) { return have && is_unit_of(u); };

		// right-hand side of this unit including the coupling term
		auto rhs = [&](P (*F)[UD], P (*X)[UD], P T) {
			P cpt[AtLeastOne<NC>::value], cpf[UPT][AtLeastOne<NC>::value];
			if (MMA)
			{
				// all 8 systems of the block at once on the tensor cores; two block barriers
				if (live)
				{
					Sys::OdeFunction(gsys, unit, F[0], X[0], T, upar[0], S.spar, s_gpar, s_igpar, uacc[0], iuacc[0], S.sacc, S.isacc, cpt, cpf[0]);
					#pragma unroll
					for (int k = 0; k < NC; k++)
					{
						if constexpr (sizeof(P) == 8) s_cptm[(k * G::NCOL + slot) * G::CS + unit] = cpt[k];
						else
						{
							float big, small;
							tf32_split((float)cpt[k], big, small);
							s_cptm[(k * G::NCOL + slot) * G::CS + unit] = (P)big;
							s_cptm[CPT_ELEMS + (k * G::NCOL + slot) * G::CS + unit] = (P)small;
						}
					}
				}
				__syncthreads();
				#pragma unroll
				for (int k = 0; k < NC; k++)
				{
					if constexpr (sizeof(P) == 8)
						mma_coupling_product<MMA ? UPS : 32, MMA ? SPBK : 8>((const double*)s_matrix + (size_t)k * UPS * UPS, (const double*)s_cptm + k * G::NCOL * G::CS,
						                                                     (double*)s_vpart + k * G::KG * G::NCOL * G::VS, warp, lane);
					else
						mma_coupling_product_f32<MMA ? UPS : 32, MMA ? SPBK : 8>((const float4*)((const float*)s_matrix + (size_t)k * UPS * UPS), MSIZE / 4,
						                                                         (const float*)s_cptm + k * G::NCOL * G::CS, CPT_ELEMS,
						                                                         (float*)s_vpart + k * G::KG * G::NCOL * G::VS, warp, lane);
				}
				__syncthreads();
				if (live)
				{
					#pragma unroll
					for (int k = 0; k < NC; k++)
					{
						P v = 0;
						#pragma unroll
						for (int g = 0; g < G::KG; g++) v += s_vpart[((k * G::KG + g) * G::NCOL + slot) * G::VS + unit];
						const P add = S.strength[k] * cpf[0][k] * v;
						const int ci = s_cidx[k];
						#pragma unroll
						for (int j = 0; j < UD; j++) if (j == ci) F[0][j] += add;
					}
				}
			}
			else
			{
				// one slot barrier
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u))
					{
						Sys::OdeFunction(gsys, unit_of(u), F[u], X[u], T, upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc, cpt, cpf[u]);
						#pragma unroll
						for (int k = 0; k < NC; k++) (BIG ? s_cptbig + (size_t)((slot * 2 + buf) * NC + k) * UPSP : S.cpt[buf][k])[unit_of(u)] = cpt[k];
					}
				slot_barrier<TPS>(slot);
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u))
					{
						#pragma unroll
						for (int k = 0; k < NC; k++)
						{
							const P v = coupling_value<P, UPS, CBW, CCI>(matrix + k * MSTRIDE, BIG ? s_cptbig + (size_t)((slot * 2 + buf) * NC + k) * UPSP : S.cpt[buf][k], unit_of(u));
							const P add = S.strength[k] * cpf[u][k] * v;
							const int ci = s_cidx[k];
							#pragma unroll
							for (int j = 0; j < UD; j++) if (j == ci) F[u][j] += add;
						}
					}
				buf ^= 1;
			}
		};

		// ======================= one attempted step of the whole system =====================================
		bool terminated = false;
		const P t = S.t;
		P dt = dt_new;
		bool end_reached = false;
		{
			const P remaining = S.td[1] - t;
			if (dt > remaining) { dt = remaining; end_reached = true; }
		}
		if (unit == 0) S.dt = dt;
		attempted++;
		bool update = true, finite = true;

		P xn[UPT][UD];
		if (ALG == kAlgRK4)
		{
			// Steppers.cuh:2015-2227 (expression shapes of the coupled stepper)
			P X[UPT][UD], k[UPT][UD];
			rhs(xn, x, t);
			P T = t + P(0.5) * dt;
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + xn[u][i] * P(0.5) * dt;
			rhs(k, X, T);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) { xn[u][i] += 2 * k[u][i]; X[u][i] = x[u][i] + k[u][i] * P(0.5) * dt; }
			rhs(k, X, T);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) { xn[u][i] += P(2.0) * k[u][i]; X[u][i] = x[u][i] + k[u][i] * dt; }
			T = t + dt;
			rhs(k, X, T);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++)
				{
					xn[u][i] = x[u][i] + dt * (xn[u][i] + k[u][i]) * P(1.0 / 6.0);
					finite = finite && (Num<P>::finite(xn[u][i]) || !live_of(u));
				}
			// per-system "all finite" (ErrorControllers.cuh:421-441): not finite => message, terminate, step not committed
			int f = finite || (UPT == 1 && !live);
			f = __all_sync(0xffffffffu, f);
			if (WPS > 1)
			{
				if (lane == 0) S.red_flags[wis] = f;
				slot_barrier<TPS>(slot);
				#pragma unroll
				for (int w = 0; w < WPS; w++) f = f && S.red_flags[w];
			}
			if (!f)
			{
				if (a.DeviceMessages && unit == 0) printf("Error: State is not a finite number. Try to use smaller step size. (global system id: %d)\n", gsys);
				terminated = true;  update = false;
			}
			dt_new = a.InitialTimeStep;
		}
		else
		{
			// Steppers.cuh:2228-2544
			P X[UPT][UD], k1[UPT][UD], k2[UPT][UD], k3[UPT][UD], k4[UPT][UD], k5[UPT][UD], k6[UPT][UD];
			rhs(k1, x, t);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_A21)) * dt;
			rhs(k2, X, t + ck<P>(CK_C2) * dt);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_A31) + k2[u][i] * ck<P>(CK_A32)) * dt;
			rhs(k3, X, t + ck<P>(CK_C3) * dt);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_A41) + k2[u][i] * ck<P>(CK_A42) + k3[u][i] * ck<P>(CK_A43)) * dt;
			rhs(k4, X, t + ck<P>(CK_C4) * dt);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_A51) + k2[u][i] * ck<P>(CK_A52) + k3[u][i] * ck<P>(CK_A53) + k4[u][i] * ck<P>(CK_A54)) * dt;
			rhs(k5, X, t + dt);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) X[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_A61) + k2[u][i] * ck<P>(CK_A62) + k3[u][i] * ck<P>(CK_A63) + k4[u][i] * ck<P>(CK_A64) + k5[u][i] * ck<P>(CK_A65)) * dt;
			rhs(k6, X, t + ck<P>(CK_C6) * dt);

			// per-unit error ratio q = max_i err_i/tol_i, accept flag, finite flag (ErrorControllers.cuh:483-500)
			P q = P(0);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
			{
				P qu = P(0);
				bool update_u = true, finite_u = true;
				#pragma unroll
				for (int i = 0; i < UD; i++)
				{
					xn[u][i] = x[u][i] + (k1[u][i] * ck<P>(CK_B1) + k3[u][i] * ck<P>(CK_B3) + k4[u][i] * ck<P>(CK_B4) + k6[u][i] * ck<P>(CK_B6)) * dt;
					const P e = k1[u][i] * ck<P>(CK_E1) + k3[u][i] * ck<P>(CK_E3) + k4[u][i] * ck<P>(CK_E4) + k5[u][i] * ck<P>(CK_E5) + k6[u][i] * ck<P>(CK_E6);
					const P err = dt * fabs(e) + ck<P>(CK_ERRFLOOR);
					finite_u = finite_u && Num<P>::finite(xn[u][i]) && Num<P>::finite(err);
					const P tol = fmax(a.RelativeTolerance[i] * fmax(fabs(xn[u][i]), fabs(x[u][i])), a.AbsoluteTolerance[i]);
					update_u = update_u && (err < tol);
					qu = fmax(qu, err / tol);
				}
				if (!live_of(u)) { qu = P(0); update_u = true; finite_u = true; }
				q = fmax(q, qu);  update = update && update_u;  finite = finite && finite_u;
			}
			// slot-wide reduction: shuffles inside the warp, one word per warp across warps
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) q = fmax(q, __shfl_xor_sync(0xffffffffu, q, o));
			int flags = (__all_sync(0xffffffffu, update) ? 1 : 0) | (__all_sync(0xffffffffu, finite) ? 2 : 0);
			if (WPS > 1)
			{
				if (lane == 0) { S.red_ratio[wis] = q;  S.red_flags[wis] = flags; }
				slot_barrier<TPS>(slot);
				#pragma unroll
				for (int w = 0; w < WPS; w++) { q = fmax(q, S.red_ratio[w]);  flags &= S.red_flags[w]; }
			}
			update = (flags & 1) != 0;
			finite = (flags & 2) != 0;

			// every thread evaluates the controller (ErrorControllers.cuh:504-539) from the reduced values
			const bool at_minimum = dt < a.MinimumTimeStepGuard;
			P mult;
			const unsigned qt = Num<P>::top(q) - Num<P>::kLow;       // 2^-100 (7.9e-31) <= q < 2^30; below that the exact path applies the reference's cap rel = min(1e30, 1/q)
			if (finite && qt < Num<P>::kSpan)
				mult = P(0.8) * inverse_root_fast<P>(q, update);
			else if (MMA && !have)
				mult = P(1.0);                                      // a slot without a system: nothing to control
			else
			{
				const P rel = fmin(P(1e30), P(1.0) / q);
				mult = P(0.8) * (pow)(rel, update ? P(1.0 / 5.0) : P(1.0 / 4.0));
				if (!Num<P>::finite(mult)) finite = false;
			}
			if (!finite)
			{
				mult = a.TimeStepShrinkLimit;
				update = false;
				if (at_minimum)
				{
					if (a.DeviceMessages && unit == 0) printf("Error: State is not a finite number, minimum step size reached. Try to use less stringent tolerances. (global system id: %d)\n", gsys);
					terminated = true;
				}
			}
			else if (at_minimum)
			{
				if (a.DeviceMessages && unit == 0 && have) printf("Warning: Minimum step size reached! Continue with fixed minimum step size! Tolerance cannot be guaranteed! (global system id: %d)\n", gsys);
				update = true;
			}
			mult = fmin(mult, a.TimeStepGrowLimit);
			mult = fmax(mult, a.TimeStepShrinkLimit);
			dt_new = dt * mult;
			dt_new = fmin(dt_new, a.MaximumTimeStep);
			dt_new = fmax(dt_new, a.MinimumTimeStep);
		}
		if (MMA && !have) { update = false;  terminated = false; }     // idle slot: no commit, no store

		// ---- events: the earliest secant estimate over all units and events (EventHandling.cuh:216-268) ----
		P evn[UPT][AtLeastOne<NE>::value];
		if (NE > 0)
		{
			P dt_event = dt;
			int corrected = 0;
			#pragma unroll
			for (int u = 0; u < UPT; u++)
			{
				if (live_of(u) && !terminated)
					Sys::EventFunction(gsys, unit_of(u), evn[u], t + dt, dt_new, S.td, xn[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
				if (live_of(u) && update && !terminated)
				{
					#pragma unroll
					for (int i = 0; i < NE; i++)
					{
						const P et = a.EventTolerance[i];
						if (((ev[u][i] > et) && (evn[u][i] < -et) && (a.EventDirection[i] <= 0)) || ((ev[u][i] < -et) && (evn[u][i] > et) && (a.EventDirection[i] >= 0)))
						{
							dt_event = fmin(dt_event, -ev[u][i] / (evn[u][i] - ev[u][i]) * dt);
							corrected = 1;
						}
					}
				}
			}
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) dt_event = fmin(dt_event, __shfl_xor_sync(0xffffffffu, dt_event, o));
			corrected = __any_sync(0xffffffffu, corrected);
			if (WPS > 1)
			{
				if (lane == 0) { S.red_evdt[wis] = dt_event;  S.red_corrected[wis] = corrected; }
				slot_barrier<TPS>(slot);
				#pragma unroll
				for (int w = 0; w < WPS; w++) { dt_event = fmin(dt_event, S.red_evdt[w]);  corrected |= S.red_corrected[w]; }
			}
			if (corrected)
			{
				if (dt_event < a.MinimumTimeStep)
				{
					if (a.DeviceMessages && unit == 0) printf("Warning: Event cannot be detected without reducing the step size below the minimum! Event detection omitted!, (global system id: %d)\n", gsys);
				}
				else { dt_new = dt_event;  update = false; }
			}
		}

		// ---- accepted step (Solver.cuh:2670-2806) ----------------------------------------------------------
		if (update)
		{
			accepted++;
			if (unit == 0) S.t = t + dt;
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				#pragma unroll
				for (int i = 0; i < UD; i++) x[u][i] = xn[u][i];
			slot_barrier<TPS>(slot);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				if (live_of(u)) Sys::ActionAfterSuccessfulTimeStep(gsys, unit_of(u), S.udt, S.t, S.dt, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
			slot_barrier<TPS>(slot);
			if (NE > 0)
			{
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u))
					{
						#pragma unroll
						for (int i = 0; i < NE; i++)
						{
							const P et = a.EventTolerance[i];
							if (((ev[u][i] > et) && (fabs(evn[u][i]) < et) && (a.EventDirection[i] <= 0)) || ((ev[u][i] < -et) && (fabs(evn[u][i]) < et) && (a.EventDirection[i] >= 0)))
								Sys::ActionAfterEventDetection(gsys, unit_of(u), i, S.udt, S.t, S.dt, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
						}
					}
				slot_barrier<TPS>(slot);
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u)) Sys::EventFunction(gsys, unit_of(u), ev[u], S.t, dt_new, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
				slot_barrier<TPS>(slot);
			}
		}
		const bool user_stop = update && (S.udt == 1);

		if (DENSE && have)   // DenseOutput.cuh:362-395, then :314-360
		{
			const int di = S.dense_index;
			bool store = false;
			if (update)
			{
				const bool room = di < a.DenseCapacity;
				store = room && (S.dense_next_time < S.t) && (S.dense_skipped >= (a.DenseOutputSaveFrequency - 1));
				if (room && (end_reached || user_stop)) store = true;
			}
			if (store)
			{
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u))
					{
						#pragma unroll
						for (int i = 0; i < UD; i++) a.DenseOutputStates[ug_of(u) + i * NU + (size_t)di * NU * UD] = x[u][i];
					}
				if (unit == 0) a.DenseOutputTimeInstances[sys + (size_t)di * a.NS] = S.t;
			}
			slot_barrier<TPS>(slot);
			if (unit == 0)
			{
				if (store) { S.dense_index = di + 1;  S.dense_skipped = 0;  S.dense_next_time = fmin(S.t + a.DenseOutputMinimumTimeStep, S.td[1]); }
				else if (update) S.dense_skipped++;
			}
			slot_barrier<TPS>(slot);
		}

		if (update && (end_reached || user_stop)) terminated = true;
		const bool paused = (a.SliceMode == 1) && have && update && !terminated && (S.t >= S.slice_time);

		// ---- end of the phase for this system (Solver.cuh:2812-2887), or of its pilot slice ------------------------
		if (terminated || paused)
		{
			slot_barrier<TPS>(slot);
			if (!paused)
			{
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u)) Sys::Finalization(gsys, unit_of(u), S.dense_index, S.t, S.dt, S.td, x[u], upar[u], S.spar, s_gpar, s_igpar, uacc[u], iuacc[u], S.sacc, S.isacc);
			}
			slot_barrier<TPS>(slot);
			#pragma unroll
			for (int u = 0; u < UPT; u++)
				if (live_of(u))
				{
					const size_t g = ug_of(u);
					#pragma unroll
					for (int i = 0; i < UD; i++)   a.ActualState[g + i * NU] = x[u][i];
					#pragma unroll
					for (int i = 0; i < NUA; i++)  a.UnitAccessories[g + i * NU] = uacc[u][i];
					#pragma unroll
					for (int i = 0; i < NiUA; i++) a.IntegerUnitAccessories[g + i * NU] = iuacc[u][i];
				}
			if (paused)
			{
				#pragma unroll
				for (int u = 0; u < UPT; u++)
					if (live_of(u))
					{
						#pragma unroll
						for (int i = 0; i < NUP; i++) a.SliceUnitParameters[ug_of(u) + i * NU] = upar[u][i];
					}
				if (unit < NSP) a.SliceSystemParameters[sys + unit * a.NS] = S.spar[unit];
				if (unit < NC)  a.SliceCouplingStrength[sys + unit * a.NS] = S.strength[unit];
			}
			if (unit < NSA)  a.SystemAccessories[sys + unit * a.NS] = S.sacc[unit];
			if (unit < NiSA) a.IntegerSystemAccessories[sys + unit * a.NS] = S.isacc[unit];
			if (unit == 0)
			{
				a.ActualTime[sys] = S.t;
				a.TimeDomain[sys] = S.td[0];  a.TimeDomain[sys + a.NS] = S.td[1];
				if (DENSE) a.DenseOutputIndex[sys] = S.dense_index;
				a.StepsAttempted[sys] += attempted;
				a.StepsAccepted[sys]  += accepted;
				a.LastPhaseSteps[sys] = attempted + (resumed ? a.LastPhaseSteps[sys] : 0u);
				if (a.SliceMode == 1) a.SlicePaused[sys] = paused ? 1 : 0;
				if (paused) { a.SliceDt[sys] = dt_new;  a.SliceLastDt[sys] = S.dt;  a.SliceDenseNext[sys] = S.dense_next_time;  a.SliceSkipped[sys] = S.dense_skipped; }
			}
			slot_barrier<TPS>(slot);
			have = false;
		}
	}
}

#ifdef __CUDACC__      // from here on: kernel launches (the host-compiled test harness of tests/kernel_on_host stops here)
// ---------------------------------------------------------------------------------------------------------------
template <class Sys, class P, int UPS, int UD, int NC, int CBW, int CCI, int NUP, int NSP, int NGP, int NiGP, int NUA, int NiUA, int NSA, int NiSA,
          int NE, bool DENSE, int ALG>
struct PerBlockLauncher
{
	static constexpr int UPSP = ((UPS + 31) / 32) * 32;
	// systems of more than 1024 units: every thread integrates UPT units (the reference's multiple block launches), one system per block
	static constexpr int UPT = (UPSP + 1023) / 1024;
	static constexpr int TPS = slot_threads(UPS, UPT);      // threads of a slot
	// slots per block: fill 512 threads (one block per SM at <= 128 registers), at most 15 named barriers
	static constexpr int SPBK = (MPGOS_B200_COUPLED_THREADS / TPS) < 1 ? 1 : ((MPGOS_B200_COUPLED_THREADS / TPS) > 15 ? 15 : (MPGOS_B200_COUPLED_THREADS / TPS));
	static_assert(UPT <= 4, "UnitsPerSystem above 4096: not built");
	static_assert(NSP <= 32 && NSA <= 32 && NiSA <= 32 && NC <= 32, "system-scope arrays are staged by the first warp");

	// tensor-core coupling: dense FP64 matrices shared by all systems, whole-warp systems, 8 systems x UPS threads <= 512
	static constexpr bool kCanMMA = CBW == 0 && CCI == 0 && NC > 0 && (UPS == 32 || UPS == 64 || (UPS == 128 && sizeof(P) == 8));
#ifndef MPGOS_B200_MMA_SLOTS
#define MPGOS_B200_MMA_SLOTS 8
#endif
	static constexpr int  kMmaSlots = (UPS == 128) ? 4 : MPGOS_B200_MMA_SLOTS;      // 512 threads; with 128 units half of the instruction's N extent is used
	// dynamic shared memory a block may ask for (227 KB per SM less the static part: slots, reduction words)
	static constexpr size_t kDynamicSharedLimit = 200 * 1024;
	template <int SLOTS> static constexpr size_t mma_shared_bytes()
	{
		using G = typename MmaGeometryFor<P, kCanMMA ? UPS : 32, kCanMMA ? SLOTS : 8>::type;
		return sizeof(P) * ((size_t)packed_matrix_elements(UPS, NC, CBW, CCI) * G::MCOPIES + (size_t)NC * G::NCOL * G::CS * G::CCOPIES + (size_t)NC * G::KG * G::NCOL * G::VS + NGP) + sizeof(int) * NiGP + 16;
	}

	template <int SLOTS, bool MMA>
	static int enqueue(const mpgos_coupled_launch* L, const PerBlockArgs<P, UD, NE>& a, cudaStream_t stream)
	{
		auto kernel = perblock_kernel<Sys, P, UPS, UD, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, DENSE, ALG, SLOTS, MMA, MMA ? 1 : UPT>;
		using G = typename MmaGeometryFor<P, MMA ? UPS : 32, MMA ? SLOTS : 8>::type;
		const size_t extra = MMA ? (size_t)NC * G::NCOL * G::CS * G::CCOPIES + (size_t)NC * G::KG * G::NCOL * G::VS : 0;
		constexpr bool BIG = !MMA && UPT > 1;
		const size_t big = BIG ? sizeof(P) * (size_t)SLOTS * 2 * NC * UPSP + 16 : 0;      // coupling terms of a system of more than 1024 units
		size_t smem = sizeof(P) * ((size_t)packed_matrix_elements(UPS, NC, CBW, CCI) * (MMA ? G::MCOPIES : 1) + extra + NGP) + sizeof(int) * NiGP + 16 + big;
		PerBlockArgs<P, UD, NE> args = a;
		args.MatrixInGlobal = 0;
		if (smem > kDynamicSharedLimit)
		{
			// the matrices do not fit next to the block's other shared data: the tensor-core variant needs them there (launch() does not
			// come here with it), the scalar variant reads them from global memory instead
			if (MMA) return (int)cudaErrorInvalidConfiguration;
			smem -= sizeof(P) * (size_t)packed_matrix_elements(UPS, NC, CBW, CCI);
			args.MatrixInGlobal = 1;
		}
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return (int)e;
		int per_sm = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SLOTS * (MMA ? UPSP : TPS), smem);
		if (e != cudaSuccess) return (int)e;
		if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
		int grid = (a.Active + SLOTS - 1) / SLOTS;
		const int resident = per_sm * (L->MultiProcessorCount > 0 ? L->MultiProcessorCount : 148);
		if (grid > resident) grid = resident;
		e = cudaMemsetAsync(L->d_WorkQueue, 0, sizeof(unsigned int), stream);
		if (e != cudaSuccess) return (int)e;
		// queue order: with the previous phase's step counts at hand and more systems than two rounds of slots, the longest first
		// (C5, whose expensive systems sit at the end of the amplitude sweep: 18.0 instead of 19.4 ms)
		args.DispatchOrder = nullptr;  args.SliceMode = 0;
		const bool many = (a.Active > 2 * grid * SLOTS) || L->DispatchHistory >= 2 || L->DispatchPilot >= 2;      // >= 2: forced (tests)
		const bool can_order = L->d_DispatchOrder && L->d_DispatchScratch && many;
		auto order_by_cost = [&]() -> cudaError_t
		{
			unsigned int* bucket_of = L->d_DispatchScratch;
			unsigned int* hist = L->d_DispatchScratch + L->desc.NS;
			cudaError_t err = cudaMemsetAsync(hist, 0, 256 * sizeof(unsigned int), stream);
			if (err != cudaSuccess) return err;
			dispatch_cost_histogram<<<(a.Active + 7) / 8, 256, 0, stream>>>(L->d_LastPhaseSteps, a.Active, 1, a.Active, bucket_of, hist);
			dispatch_bucket_starts<<<1, 256, 0, stream>>>(hist);
			dispatch_scatter<<<(a.Active + 255) / 256, 256, 0, stream>>>(bucket_of, hist, a.Active, L->d_DispatchOrder);
			return cudaGetLastError();
		};
		if (L->DispatchHistory && can_order)
		{
			e = order_by_cost();
			if (e != cudaSuccess) return (int)e;
			args.DispatchOrder = L->d_DispatchOrder;
		}
		else if (L->DispatchPilot && can_order && L->d_SliceState && L->d_SliceParameters)
		{
			// no history yet: the first tenth of the phase in index order tells what each system costs (for sweeps whose cost is set by a
			// parameter -- C5's pressure amplitude), the rest follows in that order.  Emulated with two phases before it was built:
			// 18.3 instead of 19.4 ms on C5 (tools/probe_time_slices_coupled.py).
			P* base = (P*)L->d_SliceState;
			args.SliceDt = base;  args.SliceLastDt = base + L->desc.NS;  args.SliceDenseNext = base + 2 * (size_t)L->desc.NS;
			args.SliceSkipped = (int*)(base + 3 * (size_t)L->desc.NS);  args.SlicePaused = args.SliceSkipped + L->desc.NS;
			P* pbase = (P*)L->d_SliceParameters;
			args.SliceUnitParameters = pbase;  args.SliceSystemParameters = pbase + (size_t)NUP * L->desc.NS * UPS;
			args.SliceCouplingStrength = args.SliceSystemParameters + (size_t)NSP * L->desc.NS;
			args.SliceMode = 1;  args.SliceFraction = P(0.1);
			kernel<<<grid, SLOTS * (MMA ? UPSP : TPS), smem, stream>>>(args);
			e = cudaGetLastError();
			if (e == cudaSuccess) e = order_by_cost();
			if (e == cudaSuccess) e = cudaMemsetAsync(L->d_WorkQueue, 0, sizeof(unsigned int), stream);
			if (e != cudaSuccess) return (int)e;
			args.SliceMode = 2;  args.DispatchOrder = L->d_DispatchOrder;
		}
		kernel<<<grid, SLOTS * (MMA ? UPSP : TPS), smem, stream>>>(args);
		return (int)cudaGetLastError();
	}

	static int launch(const mpgos_coupled_launch* L, void* stream_v)
	{
		cudaStream_t stream = (cudaStream_t)stream_v;
		if (L->AbiVersion != MPGOS_B200_ABI_VERSION) return (int)cudaErrorInvalidValue;      // translation unit built against another header
		const mpgos_coupled_desc& d = L->desc;
		if (d.UPS != UPS || d.UD != UD || d.NC != NC || d.CBW != CBW || d.CCI != CCI || d.NUP != NUP || d.NSP != NSP || d.NGP != NGP || d.NiGP != NiGP ||
		    d.NUA != NUA || d.NiUA != NiUA || d.NSA != NSA || d.NiSA != NiSA || d.NE != NE || d.Algorithm != ALG || d.PrecisionBytes != (int)sizeof(P) || (d.NDO > 0) != DENSE)
			return (int)cudaErrorInvalidValue;
		PerBlockArgs<P, UD, NE> a;
		a.TimeDomain = (P*)L->d_TimeDomain;  a.ActualState = (P*)L->d_ActualState;  a.ActualTime = (P*)L->d_ActualTime;  a.UnitParameters = (P*)L->d_UnitParameters;
		a.SystemParameters = (P*)L->d_SystemParameters;  a.GlobalParameters = (P*)L->d_GlobalParameters;  a.IntegerGlobalParameters = L->d_IntegerGlobalParameters;
		a.UnitAccessories = (P*)L->d_UnitAccessories;  a.IntegerUnitAccessories = L->d_IntegerUnitAccessories;  a.SystemAccessories = (P*)L->d_SystemAccessories;
		a.IntegerSystemAccessories = L->d_IntegerSystemAccessories;  a.CouplingMatrix = (P*)L->d_CouplingMatrix;  a.CouplingStrength = (P*)L->d_CouplingStrength;
		a.CouplingIndex = L->d_CouplingIndex;  a.DenseOutputIndex = L->d_DenseOutputIndex;  a.DenseOutputTimeInstances = (P*)L->d_DenseOutputTimeInstances;
		a.DenseOutputStates = (P*)L->d_DenseOutputStates;  a.StepsAttempted = L->d_StepsAttempted;  a.StepsAccepted = L->d_StepsAccepted;  a.WorkQueue = L->d_WorkQueue;
		a.LastPhaseSteps = L->d_LastPhaseSteps;  a.DispatchOrder = nullptr;
		a.SliceMode = 0;  a.SliceFraction = P(0);  a.SliceDt = a.SliceLastDt = a.SliceDenseNext = nullptr;  a.SliceSkipped = a.SlicePaused = nullptr;
		a.SliceUnitParameters = a.SliceSystemParameters = a.SliceCouplingStrength = nullptr;
		a.MatrixInGlobal = 0;
		a.ShardIndex = d.ShardIndex;  a.ShardCount = d.ShardCount;  a.ShardTile = d.ShardTile > 0 ? d.ShardTile : 1;
		a.NS = d.NS;  a.Active = L->ActiveSystems < d.NS ? L->ActiveSystems : d.NS;  a.SystemOffset = d.SystemOffset;  a.DenseCapacity = d.NDO;
		a.DenseOutputSaveFrequency = L->DenseOutputSaveFrequency;  a.DeviceMessages = L->DeviceMessages;
		a.InitialTimeStep = (P)L->InitialTimeStep;  a.MaximumTimeStep = (P)L->MaximumTimeStep;  a.MinimumTimeStep = (P)L->MinimumTimeStep;
		a.TimeStepGrowLimit = (P)L->TimeStepGrowLimit;  a.TimeStepShrinkLimit = (P)L->TimeStepShrinkLimit;  a.DenseOutputMinimumTimeStep = (P)L->DenseOutputMinimumTimeStep;
		a.MinimumTimeStepGuard = a.MinimumTimeStep * P(1.01);
		for (int i = 0; i < UD; i++) { a.RelativeTolerance[i] = (P)L->RelativeTolerance[i];  a.AbsoluteTolerance[i] = (P)L->AbsoluteTolerance[i]; }
		for (int i = 0; i < AtLeastOne<NE>::value; i++) { a.EventTolerance[i] = (i < NE) ? (P)L->EventTolerance[i] : P(0);  a.EventDirection[i] = (i < NE) ? L->EventDirection[i] : 0; }
		if (a.Active <= 0) return (int)cudaSuccess;
		if constexpr (kCanMMA)
		{
			if (L->TensorCoreCoupling && mma_shared_bytes<kMmaSlots>() <= kDynamicSharedLimit) return enqueue<kMmaSlots, true>(L, a, stream);
		}
		return enqueue<SPBK, false>(L, a, stream);
	}
};
#endif // __CUDACC__

} // namespace mpgos_b200
#endif
