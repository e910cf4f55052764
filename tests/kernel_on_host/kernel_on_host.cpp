// TEST INFRASTRUCTURE ONLY -- the product's perthread_kernel<> body compiled for the host (cuda_on_host.h) behind the checker's
// C-ABI (oracle/oracle_abi.h), one library per (system, algorithm, precision, dense):
//   -DKOH_SYSTEM=Lorenz -DKOH_PREC=double -DKOH_ALG=1 -DKOH_DENSE=0 -DKOH_NDO=0
// Each instance is one call of the kernel function with blockIdx.x = instance (direct mapping, one thread per block).
#ifdef KOH_REFILL      // the persistent-warp lane mapping needs real warps: one std::thread per CUDA thread (cuda_block_on_host.h)
#include "cuda_block_on_host.h"
#include <thread>
#else
#include "cuda_on_host.h"
#endif
#include "../../include/mpgos_b200/perthread_kernel.cuh"
#include "../../massively-parallel-gpu-ode-solver_b200/csrc/builtin_systems.cuh"
#include "../../oracle/oracle_abi.h"
#include <vector>

using P = KOH_PREC;
#ifdef KOH_USER_SYSDEF
// A user's system definition, unmodified: the six PerThread_* callbacks of a reference tutorial, included where it lies under
// /root/reference at build time (never copied), bound exactly as include/SingleSystem_PerThread_Interface.cuh binds them
// (mpgos_b200::UserCallbacks; that header also holds the host object, which needs nvcc and the runtime library).
#include KOH_USER_SYSDEF
struct Sys
{
	static inline void OdeFunction(int tid, int NT, P* F, P* X, P T, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_OdeFunction(tid, NT, F, X, T, cPAR, sPAR, sPARi, ACC, ACCi); }
	static inline void EventFunction(int tid, int NT, P* EF, P T, P dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_EventFunction(tid, NT, EF, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static inline void ActionAfterEventDetection(int tid, int NT, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_ActionAfterEventDetection(tid, NT, IDX, UDT, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static inline void ActionAfterSuccessfulTimeStep(int tid, int NT, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_ActionAfterSuccessfulTimeStep(tid, NT, UDT, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static inline void Initialization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_Initialization(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static inline void Finalization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_Finalization(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
};
enum { SD = KOH_SD, NCP = KOH_NCP, NSP = KOH_NSP, NISP = KOH_NISP, NE = KOH_NE, NA = KOH_NA, NIA = KOH_NIA };
#else
using Sys = mpgos_b200::builtin::KOH_SYSTEM;
enum { SD = Sys::SD, NCP = Sys::NCP, NSP = Sys::NSP, NISP = Sys::NISP, NE = Sys::NE, NA = Sys::NA, NIA = Sys::NIA };
#endif

extern "C" void ref_perthread_shape(oracle_perthread_shape* s)
{
	s->SD = SD; s->NCP = NCP; s->NSP = NSP; s->NISP = NISP; s->NE = NE; s->NA = NA; s->NIA = NIA; s->NDO = KOH_NDO;
	s->Algorithm = KOH_ALG;  s->Precision = (int)sizeof(P);
}

extern "C" int ref_perthread_run(const oracle_perthread_buffers* b, const oracle_perthread_options* o)
{
	using namespace mpgos_b200;
	const int NT = b->NT;
	std::vector<unsigned long long> att(NT, 0), acc(NT, 0);
	std::vector<unsigned int> last(NT, 0);
	unsigned int queue = 0;
	PerThreadArgs<P, SD, NE, NSP, NISP> a;
	std::memset(&a, 0, sizeof(a));
	a.TimeDomain = (P*)b->TimeDomain;  a.ActualState = (P*)b->ActualState;  a.ActualTime = (P*)b->ActualTime;
	a.ControlParameters = (P*)b->ControlParameters;  a.SharedParameters = (P*)b->SharedParameters;
	a.IntegerSharedParameters = b->IntegerSharedParameters;  a.Accessories = (P*)b->Accessories;  a.IntegerAccessories = b->IntegerAccessories;
	a.DenseOutputIndex = b->DenseOutputIndex;  a.DenseOutputTimeInstances = (P*)b->DenseOutputTimeInstances;  a.DenseOutputStates = (P*)b->DenseOutputStates;
	a.StepsAttempted = att.data();  a.StepsAccepted = acc.data();  a.WorkQueue = &queue;  a.LastPhaseSteps = last.data();  a.DispatchOrder = nullptr;
	a.Stride = NT;  a.Active = o->NumberOfActiveThreads < NT ? o->NumberOfActiveThreads : NT;  a.InstanceOffset = 0;  a.GlobalInstances = NT;
	a.DenseCapacity = KOH_NDO;  a.DenseOutputSaveFrequency = o->DenseOutputSaveFrequency;  a.DeviceMessages = 0;  a.RefillThreshold = 1;
	// exactly PerThreadLauncher::launch (perthread_kernel.cuh), which g++ cannot see (it sits behind __CUDACC__)
	a.InitialTimeStep = (P)o->InitialTimeStep;  a.MaximumTimeStep = (P)o->MaximumTimeStep;  a.MinimumTimeStep = (P)o->MinimumTimeStep;
	a.TimeStepGrowLimit = (P)o->TimeStepGrowLimit;  a.TimeStepShrinkLimit = (P)o->TimeStepShrinkLimit;
	a.DenseOutputMinimumTimeStep = (P)o->DenseOutputMinimumTimeStep;
	a.MinimumTimeStepGuard = a.MinimumTimeStep * P(1.01);
	a.GuardAbove   = (float)((double)a.MinimumTimeStepGuard * (1.0 + 1e-4));
	a.ShrinkAbove  = (float)((double)a.TimeStepShrinkLimit * (1.0 + 1e-4));
	a.GrowBelow    = (float)((double)a.TimeStepGrowLimit * (1.0 - 1e-4));
	a.MinimumAbove = (float)((double)a.MinimumTimeStep * (1.0 + 1e-4));
	a.MaximumBelow = (float)((double)a.MaximumTimeStep * (1.0 - 1e-4));
	for (int i = 0; i < SD; i++) { a.RelativeTolerance[i] = ((const P*)b->RelativeTolerance)[i];  a.AbsoluteTolerance[i] = ((const P*)b->AbsoluteTolerance)[i]; }
	for (int i = 0; i < NE; i++) { a.EventTolerance[i] = ((const P*)b->EventTolerance)[i];  a.EventDirection[i] = b->EventDirection[i]; }
	a.EventScreen = 1;
	for (int i = 0; i < NE; i++) if (!(a.EventTolerance[i] >= P(0))) a.EventScreen = 0;
	if (NSP <= kSharedByValueLimit && NISP <= kSharedByValueLimit)
	{
		for (int i = 0; i < NSP; i++)  a.SharedValues[i] = ((const P*)b->SharedParameters)[i];
		for (int i = 0; i < NISP; i++) a.IntegerSharedValues[i] = b->IntegerSharedParameters[i];
	}

#ifdef KOH_REFILL
	// LaneRefill = KOH_REFILL idle lanes per warp trigger a refill; 2 blocks of 64 threads (2 warps each) share the instance queue
	a.RefillThreshold = KOH_REFILL;
	for (int phase = 0; phase < o->Phases; phase++)
	{
		queue = 0;
		for (int block = 0; block < 2; block++)
		{
			std::vector<std::thread> pool;
			for (int t = 0; t < 64; t++)
				pool.emplace_back([=]() {
					host_threadIdx.x = (unsigned)t;  host_blockIdx.x = (unsigned)block;  host_blockDim.x = 64;  host_gridDim.x = 2;
					perthread_kernel<Sys, P, SD, NCP, NSP, NISP, NE, NA, NIA, (KOH_DENSE != 0), KOH_ALG, true, 2>(a);
				});
			for (auto& th : pool) th.join();
		}
	}
#else
	host_blockDim.x = 1;  host_threadIdx.x = 0;  host_gridDim.x = (unsigned)NT;
	for (int phase = 0; phase < o->Phases; phase++)
		for (int inst = 0; inst < NT; inst++)
		{
			host_blockIdx.x = (unsigned)inst;
			perthread_kernel<Sys, P, SD, NCP, NSP, NISP, NE, NA, NIA, (KOH_DENSE != 0), KOH_ALG, false, 2>(a);
		}
#endif
	for (int i = 0; i < NT; i++)
	{
		if (b->steps_attempted) b->steps_attempted[i] += (long long)att[i];
		if (b->steps_accepted)  b->steps_accepted[i]  += (long long)acc[i];
	}
	return 0;
}
