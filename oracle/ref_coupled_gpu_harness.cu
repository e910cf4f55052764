// TEST / BASELINE INFRASTRUCTURE ONLY -- never part of the product path.
// The REFERENCE MPGOS coupled module itself, unmodified, compiled by nvcc for sm_100a from /root/reference
// (CoupledSystems_PerBlock_Interface.cuh + its kernels + a tutorial system definition) behind the checker ABI of
// oracle_coupled_abi.h: GPU-vs-GPU parity and the "reference MPGOS build on the same B200" baseline of config C5.
// Build-time macros as in ref_coupled_harness.cpp (REF_NS is the capacity; runs use ActiveSystems).
#include <cuda_runtime.h>
#define cudaDeviceSetSharedMemConfig(x) cudaSuccess     // deprecated call in the reference constructor (Interface.cuh:373-377)

#include REF_SYSDEF
#include "CoupledSystems_PerBlock_Interface.cuh"
#include "oracle_coupled_abi.h"
#include <sstream>

typedef double P;
typedef ProblemSolver<REF_NS, REF_UPS, REF_UD, REF_TPB, REF_SPB, REF_NC, REF_CBW, REF_CCI, REF_NUP, REF_NSP, REF_NGP, REF_NiGP, REF_NUA, REF_NiUA, REF_NSA, REF_NiSA, REF_NE, REF_NDO, REF_ALGO, P> RefSolver;
static RefSolver* g_solver = nullptr;

static RefSolver& solver()
{
	if (!g_solver)
	{
		std::streambuf* keep = std::cout.rdbuf();
		std::ostringstream sink;
		std::cout.rdbuf(sink.rdbuf());
		g_solver = new RefSolver(0);
		std::cout.rdbuf(keep);
	}
	return *g_solver;
}

extern "C" void ref_coupled_shape(oracle_coupled_shape* s)
{
	s->UPS = REF_UPS; s->UD = REF_UD; s->NC = REF_NC; s->CBW = REF_CBW; s->CCI = REF_CCI; s->NUP = REF_NUP; s->NSP = REF_NSP; s->NGP = REF_NGP; s->NiGP = REF_NiGP;
	s->NUA = REF_NUA; s->NiUA = REF_NiUA; s->NSA = REF_NSA; s->NiSA = REF_NiSA; s->NE = REF_NE; s->NDO = REF_NDO; s->Algorithm = (REF_ALGO == RK4 ? 0 : 1); s->Precision = 8;
}
extern "C" int ref_coupled_capacity() { return REF_NS; }

extern "C" int ref_coupled_gpu_run(const oracle_coupled_buffers* b, const oracle_coupled_options* o, int shared_matrices, int repeats, float* kernel_ms)
{
	const int n = b->NS;
	if (n > REF_NS) return 1;
	RefSolver& S = solver();
	const size_t NU = (size_t)n * REF_UPS;
	const P* iTD = (const P*)b->TimeDomain; const P* iX = (const P*)b->ActualState; const P* iT = (const P*)b->ActualTime; const P* iUP = (const P*)b->UnitParameters;
	const P* iSP = (const P*)b->SystemParameters; const P* iUA = (const P*)b->UnitAccessories; const P* iSA = (const P*)b->SystemAccessories; const P* iCS = (const P*)b->CouplingStrength;
	const P* iM = (const P*)b->CouplingMatrix;

	S.SolverOption(SharedCouplingMatrices, shared_matrices);
	S.SolverOption(InitialTimeStep, o->InitialTimeStep);
	S.SolverOption(ActiveSystems, n < o->ActiveSystems ? n : o->ActiveSystems);
	S.SolverOption(MaximumTimeStep, o->MaximumTimeStep);
	S.SolverOption(MinimumTimeStep, o->MinimumTimeStep);
	S.SolverOption(TimeStepGrowLimit, o->TimeStepGrowLimit);
	S.SolverOption(TimeStepShrinkLimit, o->TimeStepShrinkLimit);
	S.SolverOption(DenseOutputMinimumTimeStep, o->DenseOutputMinimumTimeStep);
	S.SolverOption(DenseOutputSaveFrequency, o->DenseOutputSaveFrequency);
	for (int i = 0; i < REF_UD; i++) { S.SolverOption(RelativeTolerance, i, ((const P*)b->RelativeTolerance)[i]); S.SolverOption(AbsoluteTolerance, i, ((const P*)b->AbsoluteTolerance)[i]); }
	for (int i = 0; i < REF_NE; i++) { S.SolverOption(EventTolerance, i, ((const P*)b->EventTolerance)[i]); S.SolverOption(EventDirection, i, b->EventDirection[i]); }

	for (int k = 0; k < REF_NC; k++) for (int r = 0; r < REF_UPS; r++) for (int c = 0; c < REF_UPS; c++)
	{
		if (REF_CBW > 0)   // only entries inside the band may be set (the reference's BoundCheck exits otherwise)
		{
			int d1 = (c - r + REF_UPS) % REF_UPS, d2 = (r - c + REF_UPS) % REF_UPS;
			if ((d1 < d2 ? d1 : d2) > REF_CBW) continue;
		}
		S.SetHost(k, CouplingMatrix, r, c, iM[r + c * REF_UPS + k * REF_UPS * REF_UPS]);
	}
	for (int i = 0; i < REF_NC; i++) S.SetHost(CouplingIndex, i, b->CouplingIndex[i]);
	for (int i = 0; i < REF_NGP; i++) S.SetHost(GlobalParameters, i, ((const P*)b->GlobalParameters)[i]);
	for (int i = 0; i < REF_NiGP; i++) S.SetHost(IntegerGlobalParameters, i, b->IntegerGlobalParameters[i]);

	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);  cudaEventCreate(&e1);
	float best = -1.0f;
	for (int rep = 0; rep < (repeats > 0 ? repeats : 1); rep++)
	{
		for (int s = 0; s < n; s++)
		{
			for (int i = 0; i < 2; i++) S.SetHost(s, TimeDomain, i, iTD[s + i * n]);
			S.SetHost(s, ActualTime, iT[s]);
			if (REF_NDO > 0) S.SetHost(s, DenseIndex, b->DenseOutputIndex[s]);
			for (int i = 0; i < REF_NSP; i++) S.SetHost(s, SystemParameters, i, iSP[s + i * n]);
			for (int i = 0; i < REF_NSA; i++) S.SetHost(s, SystemAccessories, i, iSA[s + i * n]);
			for (int i = 0; i < REF_NiSA; i++) S.SetHost(s, IntegerSystemAccessories, i, b->IntegerSystemAccessories[s + i * n]);
			for (int i = 0; i < REF_NC; i++) S.SetHost(s, CouplingStrength, i, iCS[s + i * n]);
			for (int u = 0; u < REF_UPS; u++)
			{
				const size_t g = u + (size_t)s * REF_UPS;
				for (int i = 0; i < REF_UD; i++) S.SetHost(s, u, ActualState, i, iX[g + i * NU]);
				for (int i = 0; i < REF_NUP; i++) S.SetHost(s, u, UnitParameters, i, iUP[g + i * NU]);
				for (int i = 0; i < REF_NUA; i++) S.SetHost(s, u, UnitAccessories, i, iUA[g + i * NU]);
				for (int i = 0; i < REF_NiUA; i++) S.SetHost(s, u, IntegerUnitAccessories, i, b->IntegerUnitAccessories[g + i * NU]);
			}
		}
		S.SynchroniseFromHostToDevice(All);
		S.SynchroniseDevice();
		cudaEventRecord(e0, 0);
		for (int ph = 0; ph < o->Phases; ph++) S.Solve();
		cudaEventRecord(e1, 0);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		if (best < 0 || ms < best) best = ms;
	}
	cudaEventDestroy(e0);  cudaEventDestroy(e1);
	if (kernel_ms) *kernel_ms = best;

	S.SynchroniseFromDeviceToHost(All);
	S.InsertSynchronisationPoint();
	S.SynchroniseSolver();
	P* oTD = (P*)b->TimeDomain; P* oX = (P*)b->ActualState; P* oT = (P*)b->ActualTime; P* oUA = (P*)b->UnitAccessories; P* oSA = (P*)b->SystemAccessories;
	P* oDT = (P*)b->DenseOutputTimeInstances; P* oDS = (P*)b->DenseOutputStates;
	for (int s = 0; s < n; s++)
	{
		for (int i = 0; i < 2; i++) oTD[s + i * n] = S.GetHost<P>(s, TimeDomain, i);
		oT[s] = S.GetHost<P>(s, ActualTime);
		b->DenseOutputIndex[s] = S.GetHost<int>(s, DenseIndex);
		for (int i = 0; i < REF_NSA; i++) oSA[s + i * n] = S.GetHost<P>(s, SystemAccessories, i);
		for (int i = 0; i < REF_NiSA; i++) b->IntegerSystemAccessories[s + i * n] = S.GetHost<int>(s, IntegerSystemAccessories, i);
		for (int k = 0; k < REF_NDO; k++) oDT[s + (size_t)k * n] = S.GetHost<P>(s, DenseTime, k);
		for (int u = 0; u < REF_UPS; u++)
		{
			const size_t g = u + (size_t)s * REF_UPS;
			for (int i = 0; i < REF_UD; i++) oX[g + i * NU] = S.GetHost<P>(s, u, ActualState, i);
			for (int i = 0; i < REF_NUA; i++) oUA[g + i * NU] = S.GetHost<P>(s, u, UnitAccessories, i);
			for (int i = 0; i < REF_NiUA; i++) b->IntegerUnitAccessories[g + i * NU] = S.GetHost<int>(s, u, IntegerUnitAccessories, i);
			for (int k = 0; k < REF_NDO; k++) for (int i = 0; i < REF_UD; i++) oDS[g + i * NU + (size_t)k * NU * REF_UD] = S.GetHost<P>(s, u, DenseState, i, k);
		}
	}
	return 0;
}
