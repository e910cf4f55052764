// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// Runs the REFERENCE's own block-per-system kernels (SourceCodes/CoupledSystmes_PerBlock_Solver.cuh with the
// CoupledSystems_PerBlock_* numerics headers and a tutorial system definition) on the host: one std::thread per
// CUDA thread of a block, blocks in sequence, the same kernel variant the reference's Solve() would launch
// (CoupledSystems_PerBlock_Interface.cuh:1949-1965).  Built by oracle/Makefile (target refcoupled).
//
// Build-time macros: REF_SYSDEF, REF_NS (capacity), REF_UPS, REF_UD, REF_TPB, REF_SPB, REF_NC, REF_CBW, REF_CCI, REF_NUP,
// REF_NSP, REF_NGP, REF_NiGP, REF_NUA, REF_NiUA, REF_NSA, REF_NiSA, REF_NE, REF_NDO, REF_ALGO.
#include "ref_coupled_shim.h"
#include "oracle_coupled_abi.h"
#include <thread>
#include <vector>

#include "ref_coupled_structs.generated.h"    // enum Algorithms + kernel-argument structs (sed-extracted at build time)
#include REF_SYSDEF
#include "CoupledSolver.patched.cuh"          // the reference kernels (sed-patched copy in oracle/_ref, see Makefile)

#ifndef REF_PRECISION
#define REF_PRECISION double
#endif
typedef REF_PRECISION P;
#define TARGS REF_NS, REF_UPS, REF_UD, REF_TPB, REF_SPB, REF_NC, REF_CBW, REF_CCI, REF_NUP, REF_NSP, REF_NGP, REF_NiGP, REF_NUA, REF_NiUA, REF_NSA, REF_NiSA, REF_NE, REF_NDO, REF_ALGO, P

static void launch(Struct_ThreadConfiguration tc, Struct_GlobalVariables<P> g, Struct_SharedMemoryUsage sm, Struct_SolverOptions<P> so)
{
	for (int block = 0; block < tc.GridSize; block++)
	{
		g_block_barrier.reset(tc.BlockSize);
		std::vector<std::thread> threads;
		for (int t = 0; t < tc.BlockSize; t++)
			threads.emplace_back([=]() {
				threadIdx.x = t; blockIdx.x = block; blockDim.x = tc.BlockSize; gridDim.x = tc.GridSize;
				if ((REF_SPB > 1) && (REF_SPB * REF_UPS > REF_TPB))   CoupledSystems_PerBlock_MultipleSystems_MultipleBlockLaunches<TARGS>(tc, g, sm, so);
				if ((REF_SPB == 1) && (REF_SPB * REF_UPS > REF_TPB))  CoupledSystems_PerBlock_SingleSystem_MultipleBlockLaunches<TARGS>(tc, g, sm, so);
				if ((REF_SPB > 1) && (REF_SPB * REF_UPS <= REF_TPB))  CoupledSystems_PerBlock_MultipleSystems_SingleBlockLaunch<TARGS>(tc, g, sm, so);
				if ((REF_SPB == 1) && (REF_SPB * REF_UPS <= REF_TPB)) CoupledSystems_PerBlock_SingleSystem_SingleBlockLaunch<TARGS>(tc, g, sm, so);
			});
		for (auto& th : threads) th.join();
	}
}

extern "C" void ref_coupled_shape(oracle_coupled_shape* s)
{
	s->UPS = REF_UPS; s->UD = REF_UD; s->NC = REF_NC; s->CBW = REF_CBW; s->CCI = REF_CCI; s->NUP = REF_NUP; s->NSP = REF_NSP; s->NGP = REF_NGP; s->NiGP = REF_NiGP;
	s->NUA = REF_NUA; s->NiUA = REF_NiUA; s->NSA = REF_NSA; s->NiSA = REF_NiSA; s->NE = REF_NE; s->NDO = REF_NDO; s->Algorithm = (REF_ALGO == RK4 ? 0 : 1); s->Precision = (int)sizeof(P);
}
extern "C" int ref_coupled_capacity() { return REF_NS; }

extern "C" int ref_coupled_run(const oracle_coupled_buffers* b, const oracle_coupled_options* o)
{
	const int n = b->NS;
	if (n > REF_NS) return 1;
	const int NS = REF_NS, UPS = REF_UPS, UD = REF_UD;
	// thread configuration and layouts exactly as the reference's constructor (Interface.cuh:386-391, 875-879)
	Struct_ThreadConfiguration tc;
	tc.LogicalThreadsPerBlock = REF_SPB * UPS;
	tc.NumberOfBlockLaunches = tc.LogicalThreadsPerBlock / REF_TPB + (tc.LogicalThreadsPerBlock % REF_TPB == 0 ? 0 : 1);
	tc.ThreadPaddingPerBlock = tc.NumberOfBlockLaunches * REF_TPB - tc.LogicalThreadsPerBlock;
	tc.BlockSize = REF_TPB;
	tc.GridSize = NS / REF_SPB + (NS % REF_SPB == 0 ? 0 : 1);
	tc.TotalLogicalThreads = (tc.LogicalThreadsPerBlock + tc.ThreadPaddingPerBlock) * tc.GridSize;
	const size_t TL = tc.TotalLogicalThreads;
	auto lid = [&](int sys, int unit) { return (size_t)(sys / REF_SPB) * (tc.LogicalThreadsPerBlock + tc.ThreadPaddingPerBlock) + (size_t)(sys % REF_SPB) * UPS + unit; };

	size_t msize = (REF_CCI == 0 && REF_CBW == 0) ? (size_t)REF_NC * UPS * UPS : (REF_CCI == 0) ? (size_t)REF_NC * UPS * (2 * REF_CBW + 1) : (REF_CBW > 0) ? (size_t)REF_NC * (2 * REF_CBW + 1) : (size_t)REF_NC * UPS;
	std::vector<P> TD(2 * NS), X(TL * UD), T(NS), UP(TL * (REF_NUP + 1)), SP((size_t)NS * (REF_NSP + 1)), UA(TL * (REF_NUA + 1)), SA((size_t)NS * (REF_NSA + 1)), CS((size_t)NS * REF_NC), M(msize, 0);
	std::vector<int> IUA(TL * (REF_NiUA + 1)), ISA((size_t)NS * (REF_NiSA + 1)), DI(NS, 0);
	const int slack = o->Phases + 1;   // the first point of every phase is stored without an index check
	std::vector<P> DT((size_t)NS * (REF_NDO + slack)), DS(TL * UD * (REF_NDO + slack));
	const P* iTD = (const P*)b->TimeDomain; const P* iX = (const P*)b->ActualState; const P* iT = (const P*)b->ActualTime; const P* iUP = (const P*)b->UnitParameters;
	const P* iSP = (const P*)b->SystemParameters; const P* iUA = (const P*)b->UnitAccessories; const P* iSA = (const P*)b->SystemAccessories; const P* iCS = (const P*)b->CouplingStrength;
	const P* iM = (const P*)b->CouplingMatrix;
	const size_t NU = (size_t)n * UPS;
	for (int s = 0; s < n; s++)
	{
		for (int i = 0; i < 2; i++) TD[s + i * NS] = iTD[s + i * n];
		T[s] = iT[s];  DI[s] = b->DenseOutputIndex[s];
		for (int i = 0; i < REF_NSP; i++) SP[s + i * NS] = iSP[s + i * n];
		for (int i = 0; i < REF_NSA; i++) SA[s + i * NS] = iSA[s + i * n];
		for (int i = 0; i < REF_NiSA; i++) ISA[s + i * NS] = b->IntegerSystemAccessories[s + i * n];
		for (int i = 0; i < REF_NC; i++) CS[s + i * NS] = iCS[s + i * n];
		for (int u = 0; u < UPS; u++)
		{
			const size_t g = u + (size_t)s * UPS, r = lid(s, u);
			for (int i = 0; i < UD; i++) X[r + i * TL] = iX[g + i * NU];
			for (int i = 0; i < REF_NUP; i++) UP[r + i * TL] = iUP[g + i * NU];
			for (int i = 0; i < REF_NUA; i++) UA[r + i * TL] = iUA[g + i * NU];
			for (int i = 0; i < REF_NiUA; i++) IUA[r + i * TL] = b->IntegerUnitAccessories[g + i * NU];
		}
	}
	// coupling matrix packing as SetHost (Interface.cuh:993-1081)
	for (int k = 0; k < REF_NC; k++) for (int r = 0; r < UPS; r++) for (int c = 0; c < UPS; c++)
	{
		P v = iM[r + c * UPS + k * UPS * UPS];
		if (REF_CCI == 0 && REF_CBW == 0) { M[r + c * UPS + k * UPS * UPS] = v; continue; }
		int mc = c - r + REF_CBW; if (mc >= UPS) mc -= UPS; if (mc < 0) mc += UPS;
		int lim = (REF_CBW > 0) ? 2 * REF_CBW + 1 : UPS;
		if (mc >= lim) continue;
		if (REF_CCI == 0) M[r + mc * UPS + k * UPS * lim] = v; else if (r == 0) M[mc + k * lim] = v;
	}

	Struct_GlobalVariables<P> g;
	g.d_TimeDomain = TD.data(); g.d_ActualState = X.data(); g.d_ActualTime = T.data(); g.d_UnitParameters = UP.data(); g.d_SystemParameters = SP.data();
	g.d_GlobalParameters = (P*)b->GlobalParameters; g.d_IntegerGlobalParameters = b->IntegerGlobalParameters; g.d_UnitAccessories = UA.data();
	g.d_IntegerUnitAccessories = IUA.data(); g.d_SystemAccessories = SA.data(); g.d_IntegerSystemAccessories = ISA.data(); g.d_CouplingMatrix = M.data();
	g.d_CouplingStrength = CS.data(); g.d_CouplingIndex = b->CouplingIndex; g.d_DenseOutputIndex = DI.data(); g.d_DenseOutputTimeInstances = DT.data();
	g.d_DenseOutputStates = DS.data(); g.d_RelativeTolerance = (P*)b->RelativeTolerance; g.d_AbsoluteTolerance = (P*)b->AbsoluteTolerance;
	g.d_EventTolerance = (P*)b->EventTolerance; g.d_EventDirection = b->EventDirection;
	Struct_SharedMemoryUsage sm; sm.GlobalVariables = 0; sm.CouplingMatrices = 0; sm.SingleCouplingMatrixSize = (int)(msize / REF_NC); sm.IsAdaptive = (REF_ALGO == RK4 ? 0 : 1);
	Struct_SolverOptions<P> so;
	so.InitialTimeStep = o->InitialTimeStep; so.ActiveSystems = std::min(o->ActiveSystems, n); so.MaximumTimeStep = o->MaximumTimeStep; so.MinimumTimeStep = o->MinimumTimeStep;
	so.TimeStepGrowLimit = o->TimeStepGrowLimit; so.TimeStepShrinkLimit = o->TimeStepShrinkLimit; so.DenseOutputSaveFrequency = o->DenseOutputSaveFrequency;
	so.DenseOutputMinimumTimeStep = o->DenseOutputMinimumTimeStep;
	// the kernels derive the unit-scope stride from gridDim.x, so the whole grid is launched (inactive blocks return at once)
	for (int p = 0; p < o->Phases; p++) launch(tc, g, sm, so);

	P* oTD = (P*)b->TimeDomain; P* oX = (P*)b->ActualState; P* oT = (P*)b->ActualTime; P* oUA = (P*)b->UnitAccessories; P* oSA = (P*)b->SystemAccessories;
	P* oDT = (P*)b->DenseOutputTimeInstances; P* oDS = (P*)b->DenseOutputStates;
	for (int s = 0; s < n; s++)
	{
		for (int i = 0; i < 2; i++) oTD[s + i * n] = TD[s + i * NS];
		oT[s] = T[s];  b->DenseOutputIndex[s] = DI[s];
		for (int i = 0; i < REF_NSA; i++) oSA[s + i * n] = SA[s + i * NS];
		for (int i = 0; i < REF_NiSA; i++) b->IntegerSystemAccessories[s + i * n] = ISA[s + i * NS];
		for (int k = 0; k < REF_NDO; k++) oDT[s + (size_t)k * n] = DT[s + (size_t)k * NS];
		for (int u = 0; u < UPS; u++)
		{
			const size_t gi = u + (size_t)s * UPS, r = lid(s, u);
			for (int i = 0; i < UD; i++) oX[gi + i * NU] = X[r + i * TL];
			for (int i = 0; i < REF_NUA; i++) oUA[gi + i * NU] = UA[r + i * TL];
			for (int i = 0; i < REF_NiUA; i++) b->IntegerUnitAccessories[gi + i * NU] = IUA[r + i * TL];
			for (int k = 0; k < REF_NDO; k++) for (int i = 0; i < UD; i++) oDS[gi + i * NU + (size_t)k * NU * UD] = DS[r + i * TL + (size_t)k * TL * UD];
		}
	}
	return 0;
}
