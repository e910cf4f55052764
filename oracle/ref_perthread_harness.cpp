// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// Runs the REFERENCE's own thread-per-system kernel body
//     /root/reference/SourceCodes/SingleSystem_PerThread_Solver.cuh:11-437
// (with its steppers, error controllers, event handling and dense output headers) on the host,
// one call per instance, with the reference's own tutorial system definition.  Built by
// oracle/Makefile from the sources where they lie; outputs go to oracle/_ref/ only.
//
// Build-time macros:
//   REF_SYSDEF   "…_SystemDefinition.cuh" of the tutorial (found through -I)
//   REF_SD, REF_NCP, REF_NSP, REF_NISP, REF_NE, REF_NA, REF_NIA, REF_NDO   template integers
//   REF_ALGO     RK4 | RKCK45
//   REF_PREC     double | float
//   REF_RESET_DENSE_INDEX  (optional) Initialization additionally sets DOIDX = 0, as the T1
//                tutorial does (Reference_SystemDefinition.cuh:59-60); used by config C4 which
//                adds dense output to the Keller-Miksis tutorial.
//
// The kernel is instantiated with NT = 1 and called with blockIdx.x = threadIdx.x = 0, so
// `tid + i*NT` addresses a private one-instance buffer; the harness gathers/scatters between
// that buffer and the caller's SoA arrays.  None of the tutorial callbacks reads tid or NT.
#include "ref_shim.h"
#include "oracle_abi.h"
#include <vector>
#include <thread>

#include "ref_structs.generated.h"   // enum Algorithms + the four kernel-argument structs (sed-extracted)

// Count stepper invocations without touching the reference sources: rename the user's
// callbacks while their definition is parsed, then provide counting forwarders under the
// names the reference kernel looks up.
static thread_local long long g_rhs_calls = 0;
static thread_local long long g_accepted  = 0;

#define PerThread_OdeFunction                   RefSys_OdeFunction
#define PerThread_ActionAfterSuccessfulTimeStep RefSys_ActionAfterSuccessfulTimeStep
#define PerThread_Initialization                RefSys_Initialization
#include REF_SYSDEF
#undef PerThread_OdeFunction
#undef PerThread_ActionAfterSuccessfulTimeStep
#undef PerThread_Initialization

template <class P>
inline void PerThread_OdeFunction(int tid, int NT, P* F, P* X, P T, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
{
	++g_rhs_calls;
	RefSys_OdeFunction<P>(tid, NT, F, X, T, cPAR, sPAR, sPARi, ACC, ACCi);
}

template <class P>
inline void PerThread_ActionAfterSuccessfulTimeStep(int tid, int NT, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
{
	++g_accepted;
	RefSys_ActionAfterSuccessfulTimeStep<P>(tid, NT, UDT, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi);
}

template <class P>
inline void PerThread_Initialization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
{
	RefSys_Initialization<P>(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi);
#ifdef REF_RESET_DENSE_INDEX
	DOIDX = 0;
#endif
}

#include "SingleSystem_PerThread_Solver.cuh"   // the reference kernel itself (-I /root/reference/SourceCodes)

typedef REF_PREC P;
static const int SD = REF_SD, NCP = REF_NCP, NSP = REF_NSP, NISP = REF_NISP;
static const int NE = REF_NE, NA = REF_NA, NIA = REF_NIA, NDO = REF_NDO;
static const int RHS_PER_STEP = (REF_ALGO == RK4 ? 4 : 6);

template <class T> static inline T* nz(std::vector<T>& v) { return v.empty() ? (T*)nullptr : v.data(); }

static void run_range(const oracle_perthread_buffers* b, const oracle_perthread_options* o, int lo, int hi)
{
	const int NT = b->NT;
	P* gTD = (P*)b->TimeDomain;   P* gX = (P*)b->ActualState;  P* gT = (P*)b->ActualTime;
	P* gCP = (P*)b->ControlParameters;  P* gACC = (P*)b->Accessories;
	P* gDT = (P*)b->DenseOutputTimeInstances;  P* gDS = (P*)b->DenseOutputStates;

	std::vector<P> td(2), x(SD), t(1), cp(NCP > 0 ? NCP : 1), acc(NA > 0 ? NA : 1);
	std::vector<int> acci(NIA > 0 ? NIA : 1), didx(1);
	// The reference stores the first point of every phase without checking the dense index
	// (Solver.cuh:179-193), i.e. it writes one slot past NDO per phase once the buffer is full;
	// give the private buffers that much slack so the harness stays memory-safe.
	const int slack = o->Phases + 1;
	std::vector<P> dt(NDO > 0 ? NDO + slack : 1), ds(NDO > 0 ? (NDO + slack) * SD : 1);

	Struct_ThreadConfiguration tc;  tc.GridSize = 1;  tc.BlockSize = 1;  tc.NumberOfActiveThreads = 1;
	Struct_SharedMemoryUsage   sm;  sm.PreferSharedMemory = 0;  sm.IsAdaptive = (REF_ALGO == RK4 ? 0 : 1);
	Struct_SolverOptions<P>    so;
	so.InitialTimeStep = (P)o->InitialTimeStep;  so.MaximumTimeStep = (P)o->MaximumTimeStep;
	so.MinimumTimeStep = (P)o->MinimumTimeStep;  so.TimeStepGrowLimit = (P)o->TimeStepGrowLimit;
	so.TimeStepShrinkLimit = (P)o->TimeStepShrinkLimit;
	so.DenseOutputSaveFrequency = o->DenseOutputSaveFrequency;
	so.DenseOutputMinimumTimeStep = (P)o->DenseOutputMinimumTimeStep;

	Struct_GlobalVariables<P> g;
	g.d_TimeDomain = td.data();  g.d_ActualState = x.data();  g.d_ActualTime = t.data();
	g.d_ControlParameters = cp.data();  g.d_SharedParameters = (P*)b->SharedParameters;
	g.d_IntegerSharedParameters = b->IntegerSharedParameters;
	g.d_Accessories = acc.data();  g.d_IntegerAccessories = acci.data();
	g.d_RelativeTolerance = (P*)b->RelativeTolerance;  g.d_AbsoluteTolerance = (P*)b->AbsoluteTolerance;
	g.d_EventTolerance = (P*)b->EventTolerance;  g.d_EventDirection = b->EventDirection;
	g.d_DenseOutputIndex = didx.data();  g.d_DenseOutputTimeInstances = dt.data();  g.d_DenseOutputStates = ds.data();

	threadIdx.x = 0;  blockIdx.x = 0;  blockDim.x = 1;  gridDim.x = 1;

	for (int tid = lo; tid < hi; tid++)
	{
		for (int i = 0; i < 2; i++)   td[i]   = gTD[tid + i * NT];
		for (int i = 0; i < SD; i++)  x[i]    = gX[tid + i * NT];
		for (int i = 0; i < NCP; i++) cp[i]   = gCP[tid + i * NT];
		for (int i = 0; i < NA; i++)  acc[i]  = gACC[tid + i * NT];
		for (int i = 0; i < NIA; i++) acci[i] = b->IntegerAccessories[tid + i * NT];
		t[0] = gT[tid];
		didx[0] = b->DenseOutputIndex[tid];
		for (int k = 0; k < NDO; k++)
		{
			dt[k] = gDT[tid + k * NT];
			for (int i = 0; i < SD; i++) ds[i + k * SD] = gDS[tid + i * NT + (size_t)k * NT * SD];
		}

		g_rhs_calls = 0;  g_accepted = 0;
		for (int p = 0; p < o->Phases; p++)
			SingleSystem_PerThread<1, REF_SD, REF_NCP, REF_NSP, REF_NISP, REF_NE, REF_NA, REF_NIA, REF_NDO, REF_ALGO, P>(tc, g, sm, so);

		for (int i = 0; i < 2; i++)   gTD[tid + i * NT]  = td[i];
		for (int i = 0; i < SD; i++)  gX[tid + i * NT]   = x[i];
		for (int i = 0; i < NCP; i++) gCP[tid + i * NT]  = cp[i];
		for (int i = 0; i < NA; i++)  gACC[tid + i * NT] = acc[i];
		for (int i = 0; i < NIA; i++) b->IntegerAccessories[tid + i * NT] = acci[i];
		gT[tid] = t[0];
		b->DenseOutputIndex[tid] = didx[0];
		for (int k = 0; k < NDO; k++)
		{
			gDT[tid + k * NT] = dt[k];
			for (int i = 0; i < SD; i++) gDS[tid + i * NT + (size_t)k * NT * SD] = ds[i + k * SD];
		}
		if (b->steps_attempted) b->steps_attempted[tid] += g_rhs_calls / RHS_PER_STEP;
		if (b->steps_accepted)  b->steps_accepted[tid]  += g_accepted;
	}
}

extern "C" void ref_perthread_shape(oracle_perthread_shape* s)
{
	s->SD = SD; s->NCP = NCP; s->NSP = NSP; s->NISP = NISP; s->NE = NE; s->NA = NA; s->NIA = NIA; s->NDO = NDO;
	s->Algorithm = (REF_ALGO == RK4 ? 0 : 1);
	s->Precision = (int)sizeof(P);
}

extern "C" int ref_perthread_run(const oracle_perthread_buffers* b, const oracle_perthread_options* o)
{
	int active = o->NumberOfActiveThreads;
	if (active > b->NT) active = b->NT;
	int nthreads = o->HostThreads > 0 ? o->HostThreads : 1;
	if (nthreads > active) nthreads = active > 0 ? active : 1;
	if (nthreads == 1) { run_range(b, o, 0, active); return 0; }

	// interleaved chunks so that smooth cost gradients along a sweep are balanced over threads
	std::vector<std::thread> pool;
	const int chunk = 64;
	for (int w = 0; w < nthreads; w++)
		pool.emplace_back([=]() {
			for (int lo = w * chunk; lo < active; lo += nthreads * chunk)
				run_range(b, o, lo, std::min(lo + chunk, active));
		});
	for (auto& th : pool) th.join();
	return 0;
}
