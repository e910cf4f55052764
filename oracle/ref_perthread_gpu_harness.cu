// TEST / BASELINE INFRASTRUCTURE ONLY -- never part of the product path.
//
// The REFERENCE MPGOS itself, unmodified, compiled by nvcc for sm_100a from the sources where they
// lie under /root/reference (its own interface header, solver kernel and tutorial system
// definition), behind the checker ABI of oracle_abi.h.  Used for
//   * GPU-vs-GPU parity of our kernels against the reference's kernels on the same inputs, and
//   * the "reference MPGOS build on the same B200" baseline the bench reports.
// Built by oracle/Makefile (target refgpu) into oracle/_ref/libref_gpu_*.so; nothing is copied.
//
// Build-time macros: REF_SYSDEF, REF_NT (capacity; runs use the first n <= REF_NT instances through
// ActiveNumberOfThreads), REF_SD ... REF_NDO, REF_ALGO, REF_PREC, optional REF_RESET_DENSE_INDEX
// (see ref_perthread_harness.cpp).
#include <cuda_runtime.h>

// deprecated bank-size call (reference Interface.cuh:332-336): a no-op on sm_100; make sure a
// "not supported" return cannot trip the reference's exit-on-error macro
#define cudaDeviceSetSharedMemConfig(x) cudaSuccess

#define PerThread_Initialization RefSys_Initialization
#include REF_SYSDEF
#undef PerThread_Initialization

template <class P>
__forceinline__ __device__ void PerThread_Initialization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
{
	RefSys_Initialization<P>(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi);
#ifdef REF_RESET_DENSE_INDEX
	DOIDX = 0;
#endif
}

#include "SingleSystem_PerThread_Interface.cuh"   // the reference's (-I /root/reference/SourceCodes)
#include "oracle_abi.h"
#include <sstream>

typedef REF_PREC P;
typedef ProblemSolver<REF_NT, REF_SD, REF_NCP, REF_NSP, REF_NISP, REF_NE, REF_NA, REF_NIA, REF_NDO, REF_ALGO, P> RefSolver;

static RefSolver* g_solver = nullptr;

static RefSolver& solver()
{
	if (!g_solver)
	{
		std::streambuf* keep = std::cout.rdbuf();
		std::ostringstream sink;
		std::cout.rdbuf(sink.rdbuf());        // the constructor prints several screens
		g_solver = new RefSolver(0);
		std::cout.rdbuf(keep);
	}
	return *g_solver;
}

extern "C" void ref_perthread_shape(oracle_perthread_shape* s)
{
	s->SD = REF_SD; s->NCP = REF_NCP; s->NSP = REF_NSP; s->NISP = REF_NISP; s->NE = REF_NE; s->NA = REF_NA; s->NIA = REF_NIA; s->NDO = REF_NDO;
	s->Algorithm = (REF_ALGO == RK4 ? 0 : 1);
	s->Precision = (int)sizeof(P);
}

extern "C" int ref_perthread_gpu_capacity() { return REF_NT; }

// Integrates o->Phases phases of the first b->NT instances; `repeats` > 1 re-uploads the inputs and
// repeats the whole run (timing); *kernel_ms receives the best device time of the Solve() calls of
// one run (CUDA events, copies excluded).
extern "C" int ref_perthread_gpu_run(const oracle_perthread_buffers* b, const oracle_perthread_options* o, int threads_per_block, int repeats, float* kernel_ms)
{
	const int n = b->NT;
	if (n > REF_NT) return 1;
	RefSolver& S = solver();
	const P* TD = (const P*)b->TimeDomain;  const P* X = (const P*)b->ActualState;  const P* T = (const P*)b->ActualTime;
	const P* CP = (const P*)b->ControlParameters;  const P* ACC = (const P*)b->Accessories;

	S.SolverOption(ThreadsPerBlock, threads_per_block);
	S.SolverOption(InitialTimeStep, (P)o->InitialTimeStep);
	S.SolverOption(ActiveNumberOfThreads, n < o->NumberOfActiveThreads ? n : o->NumberOfActiveThreads);
	S.SolverOption(MaximumTimeStep, (P)o->MaximumTimeStep);
	S.SolverOption(MinimumTimeStep, (P)o->MinimumTimeStep);
	S.SolverOption(TimeStepGrowLimit, (P)o->TimeStepGrowLimit);
	S.SolverOption(TimeStepShrinkLimit, (P)o->TimeStepShrinkLimit);
	S.SolverOption(DenseOutputMinimumTimeStep, (P)o->DenseOutputMinimumTimeStep);
	S.SolverOption(DenseOutputSaveFrequency, o->DenseOutputSaveFrequency);
	for (int i = 0; i < REF_SD; i++)
	{
		S.SolverOption(RelativeTolerance, i, ((const P*)b->RelativeTolerance)[i]);
		S.SolverOption(AbsoluteTolerance, i, ((const P*)b->AbsoluteTolerance)[i]);
	}
	for (int i = 0; i < REF_NE; i++)
	{
		S.SolverOption(EventTolerance, i, ((const P*)b->EventTolerance)[i]);
		S.SolverOption(EventDirection, i, b->EventDirection[i]);
	}

	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);  cudaEventCreate(&e1);
	float best = -1.0f;
	for (int rep = 0; rep < (repeats > 0 ? repeats : 1); rep++)
	{
		for (int p = 0; p < n; p++)
		{
			for (int i = 0; i < 2; i++)       S.SetHost(p, TimeDomain, i, TD[p + i * n]);
			for (int i = 0; i < REF_SD; i++)  S.SetHost(p, ActualState, i, X[p + i * n]);
			for (int i = 0; i < REF_NCP; i++) S.SetHost(p, ControlParameters, i, CP[p + i * n]);
			for (int i = 0; i < REF_NA; i++)  S.SetHost(p, Accessories, i, ACC[p + i * n]);
			for (int i = 0; i < REF_NIA; i++) S.SetHost(p, IntegerAccessories, i, b->IntegerAccessories[p + i * n]);
			S.SetHost(p, ActualTime, T[p]);
			if (REF_NDO > 0) S.SetHost(p, DenseIndex, b->DenseOutputIndex[p]);
		}
		for (int i = 0; i < REF_NSP; i++)  S.SetHost(SharedParameters, i, ((const P*)b->SharedParameters)[i]);
		for (int i = 0; i < REF_NISP; i++) S.SetHost(IntegerSharedParameters, i, b->IntegerSharedParameters[i]);
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
	P* oTD = (P*)b->TimeDomain;  P* oX = (P*)b->ActualState;  P* oT = (P*)b->ActualTime;
	P* oCP = (P*)b->ControlParameters;  P* oACC = (P*)b->Accessories;
	P* oDT = (P*)b->DenseOutputTimeInstances;  P* oDS = (P*)b->DenseOutputStates;
	for (int p = 0; p < n; p++)
	{
		for (int i = 0; i < 2; i++)       oTD[p + i * n]  = S.GetHost<P>(p, TimeDomain, i);
		for (int i = 0; i < REF_SD; i++)  oX[p + i * n]   = S.GetHost<P>(p, ActualState, i);
		for (int i = 0; i < REF_NCP; i++) oCP[p + i * n]  = S.GetHost<P>(p, ControlParameters, i);
		for (int i = 0; i < REF_NA; i++)  oACC[p + i * n] = S.GetHost<P>(p, Accessories, i);
		for (int i = 0; i < REF_NIA; i++) b->IntegerAccessories[p + i * n] = S.GetHost<int>(p, IntegerAccessories, i);
		oT[p] = S.GetHost<P>(p, ActualTime);
		b->DenseOutputIndex[p] = S.GetHost<int>(p, DenseIndex);
		for (int k = 0; k < REF_NDO; k++)
		{
			oDT[p + (size_t)k * n] = S.GetHost<P>(p, DenseTime, k);
			for (int i = 0; i < REF_SD; i++) oDS[p + (size_t)i * n + (size_t)k * n * REF_SD] = S.GetHost<P>(p, DenseState, i, k);
		}
	}
	return 0;
}
