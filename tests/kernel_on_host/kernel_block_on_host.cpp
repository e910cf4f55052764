// TEST INFRASTRUCTURE ONLY -- the product's perblock_kernel<> body compiled for the host (cuda_block_on_host.h) behind the checker's
// coupled C-ABI (oracle/oracle_coupled_abi.h), one library per build:
//   -DKOH_SYSTEM=DuffingRing -DKOH_PREC=double -DKOH_UPS=14 -DKOH_CBW=1 -DKOH_CCI=1 -DKOH_ALG=1 -DKOH_NDO=100 -DKOH_SLOTS=2 -DKOH_MMA=0
// One std::thread per CUDA thread of a block (KOH_SLOTS x UPSP threads); KOH_BLOCKS blocks share the system queue, one after the other.
// Environment KOH_SLICE=1 runs every phase as pilot slice + continuation in cost order (SliceMode 1 then 2, PerBlockLauncher::enqueue).
#include "cuda_block_on_host.h"
#include "../../include/mpgos_b200/perblock_kernel.cuh"
#include "../../massively-parallel-gpu-ode-solver_b200/csrc/builtin_coupled_systems.cuh"
#include "../../oracle/oracle_coupled_abi.h"
#include <algorithm>
#include <numeric>
#include <vector>

using P = KOH_PREC;
#ifdef KOH_USER_SYSDEF
// A user's coupled system definition, unmodified: the CoupledSystems_PerBlock_* callbacks of a reference tutorial, included where it
// lies under /root/reference at build time (never copied), bound as include/CoupledSystems_PerBlock_Interface.cuh binds them
// (mpgos_b200::UserCoupledCallbacks).  -DKOH_SHAPE=UD,NC,NUP,NSP,NGP,NiGP,NUA,NiUA,NSA,NiSA,NE
#include KOH_USER_SYSDEF
struct Sys
{
	static inline void OdeFunction(int sid, int uid, P* F, P* X, P T, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC, P* CPT, P* CPF)
	{ CoupledSystems_PerBlock_OdeFunction(sid, uid, F, X, T, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC, CPT, CPF); }
	static inline void EventFunction(int sid, int uid, P* EF, P T, P dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_EventFunction(sid, uid, EF, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static inline void ActionAfterEventDetection(int sid, int uid, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_ActionAfterEventDetection(sid, uid, IDX, UDT, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static inline void ActionAfterSuccessfulTimeStep(int sid, int uid, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_ActionAfterSuccessfulTimeStep(sid, uid, UDT, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static inline void Initialization(int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_Initialization(sid, uid, DOIDX, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static inline void Finalization(int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_Finalization(sid, uid, DOIDX, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
};
static constexpr int kShape[11] = { KOH_SHAPE };
enum { UPS = KOH_UPS, CBW = KOH_CBW, CCI = KOH_CCI, UD = kShape[0], NC = kShape[1], NUP = kShape[2], NSP = kShape[3], NGP = kShape[4], NiGP = kShape[5],
       NUA = kShape[6], NiUA = kShape[7], NSA = kShape[8], NiSA = kShape[9], NE = kShape[10], UPSP = ((UPS + 31) / 32) * 32 };
#else
using Sys = mpgos_b200::builtin::KOH_SYSTEM<P>;
enum { UPS = KOH_UPS, CBW = KOH_CBW, CCI = KOH_CCI, UD = Sys::UD, NC = Sys::NC, NUP = Sys::NUP, NSP = Sys::NSP, NGP = Sys::NGP, NiGP = Sys::NiGP,
       NUA = Sys::NUA, NiUA = Sys::NiUA, NSA = Sys::NSA, NiSA = Sys::NiSA, NE = Sys::NE, UPSP = ((UPS + 31) / 32) * 32 };
#endif
#ifndef KOH_BLOCKS
#define KOH_BLOCKS 2
#endif
#ifndef KOH_UPT
#define KOH_UPT 1      // units per thread (the kernel's UPT: systems of more than 1024 units on the device; here any UPS)
#endif

extern "C" void ref_coupled_shape(oracle_coupled_shape* s)
{
	s->UPS = UPS; s->UD = UD; s->NC = NC; s->CBW = CBW; s->CCI = CCI; s->NUP = NUP; s->NSP = NSP; s->NGP = NGP; s->NiGP = NiGP;
	s->NUA = NUA; s->NiUA = NiUA; s->NSA = NSA; s->NiSA = NiSA; s->NE = NE; s->NDO = KOH_NDO; s->Algorithm = KOH_ALG; s->Precision = (int)sizeof(P);
}
extern "C" int ref_coupled_capacity() { return 1 << 20; }

static void launch(const mpgos_b200::PerBlockArgs<P, UD, NE>& args)
{
	using namespace mpgos_b200;
	const int threads = KOH_SLOTS * slot_threads(UPS, KOH_UPT);
	*args.WorkQueue = 0;
	for (int block = 0; block < KOH_BLOCKS; block++)
	{
		std::vector<std::thread> pool;
		for (int t = 0; t < threads; t++)
			pool.emplace_back([=]() {
				host_threadIdx.x = (unsigned)t;  host_blockIdx.x = (unsigned)block;  host_blockDim.x = (unsigned)threads;  host_gridDim.x = KOH_BLOCKS;
				perblock_kernel<Sys, P, UPS, UD, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, (KOH_NDO > 0), KOH_ALG, KOH_SLOTS, (KOH_MMA != 0), KOH_UPT>(args);
			});
		for (auto& th : pool) th.join();
	}
}

extern "C" int ref_coupled_run(const oracle_coupled_buffers* b, const oracle_coupled_options* o)
{
	using namespace mpgos_b200;
	const int NS = b->NS;
	// coupling matrix packed by type as the runtime's set_matrix_entry does (csrc/runtime_coupled.cu packed_index; CP/Interface.cuh:993-1081)
	const P* iM = (const P*)b->CouplingMatrix;
	std::vector<P> M((size_t)packed_matrix_elements(UPS, NC, CBW, CCI), P(0));
	for (int k = 0; k < NC; k++) for (int r = 0; r < UPS; r++) for (int c = 0; c < UPS; c++)
	{
		const P v = iM[r + c * UPS + k * UPS * UPS];
		if (CCI == 0 && CBW == 0) { M[r + c * UPS + k * UPS * UPS] = v; continue; }
		int mc = c - r + CBW;  if (mc >= UPS) mc -= UPS;  if (mc < 0) mc += UPS;
		const int lim = (CBW > 0) ? 2 * CBW + 1 : UPS;
		if (mc >= lim) continue;
		if (CCI == 0) M[r + mc * UPS + k * UPS * lim] = v; else if (r == 0) M[mc + k * lim] = v;
	}
	std::vector<unsigned long long> att(NS, 0), acc(NS, 0);
	std::vector<unsigned int> last(NS, 0);
	std::vector<int> order(NS), skipped(NS, 0), paused(NS, 0);
	std::vector<P> sdt(NS), sldt(NS), sdn(NS);
	std::vector<P> sup((size_t)NS * UPS * (NUP > 0 ? NUP : 1)), ssp((size_t)NS * (NSP > 0 ? NSP : 1)), scs((size_t)NS * (NC > 0 ? NC : 1));
	unsigned int queue = 0;
	PerBlockArgs<P, UD, NE> a;
	std::memset(&a, 0, sizeof(a));
	a.TimeDomain = (P*)b->TimeDomain;  a.ActualState = (P*)b->ActualState;  a.ActualTime = (P*)b->ActualTime;  a.UnitParameters = (P*)b->UnitParameters;
	a.SystemParameters = (P*)b->SystemParameters;  a.GlobalParameters = (P*)b->GlobalParameters;  a.IntegerGlobalParameters = b->IntegerGlobalParameters;
	a.UnitAccessories = (P*)b->UnitAccessories;  a.IntegerUnitAccessories = b->IntegerUnitAccessories;  a.SystemAccessories = (P*)b->SystemAccessories;
	a.IntegerSystemAccessories = b->IntegerSystemAccessories;  a.CouplingMatrix = M.data();  a.CouplingStrength = (P*)b->CouplingStrength;
	a.CouplingIndex = b->CouplingIndex;  a.DenseOutputIndex = b->DenseOutputIndex;  a.DenseOutputTimeInstances = (P*)b->DenseOutputTimeInstances;
	a.DenseOutputStates = (P*)b->DenseOutputStates;  a.StepsAttempted = att.data();  a.StepsAccepted = acc.data();  a.WorkQueue = &queue;
	a.LastPhaseSteps = last.data();  a.DispatchOrder = nullptr;
	a.SliceMode = 0;  a.SliceFraction = P(0);  a.SliceDt = sdt.data();  a.SliceLastDt = sldt.data();  a.SliceDenseNext = sdn.data();
	a.SliceSkipped = skipped.data();  a.SlicePaused = paused.data();
	a.SliceUnitParameters = sup.data();  a.SliceSystemParameters = ssp.data();  a.SliceCouplingStrength = scs.data();
	a.NS = NS;  a.Active = o->ActiveSystems < NS ? o->ActiveSystems : NS;  a.SystemOffset = 0;  a.DenseCapacity = KOH_NDO;
	a.DenseOutputSaveFrequency = o->DenseOutputSaveFrequency;  a.DeviceMessages = 0;
	// exactly PerBlockLauncher::launch (perblock_kernel.cuh), which g++ cannot see (it sits behind __CUDACC__)
	a.InitialTimeStep = (P)o->InitialTimeStep;  a.MaximumTimeStep = (P)o->MaximumTimeStep;  a.MinimumTimeStep = (P)o->MinimumTimeStep;
	a.TimeStepGrowLimit = (P)o->TimeStepGrowLimit;  a.TimeStepShrinkLimit = (P)o->TimeStepShrinkLimit;  a.DenseOutputMinimumTimeStep = (P)o->DenseOutputMinimumTimeStep;
	a.MinimumTimeStepGuard = a.MinimumTimeStep * P(1.01);
	for (int i = 0; i < UD; i++) { a.RelativeTolerance[i] = ((const P*)b->RelativeTolerance)[i];  a.AbsoluteTolerance[i] = ((const P*)b->AbsoluteTolerance)[i]; }
	for (int i = 0; i < NE; i++) { a.EventTolerance[i] = ((const P*)b->EventTolerance)[i];  a.EventDirection[i] = b->EventDirection[i]; }
	if (a.Active <= 0) return 0;
	{
		// dynamic shared memory the launcher would ask for (PerBlockLauncher::enqueue); the shim's buffer must hold it
		using G = typename MmaGeometryFor<P, (KOH_MMA != 0) ? UPS : 32, (KOH_MMA != 0) ? KOH_SLOTS : 8>::type;
		const size_t extra = (KOH_MMA != 0) ? (size_t)NC * G::NCOL * G::CS * G::CCOPIES + (size_t)NC * G::KG * G::NCOL * G::VS : 0;
		const size_t smem = sizeof(P) * ((size_t)packed_matrix_elements(UPS, NC, CBW, CCI) * ((KOH_MMA != 0) ? G::MCOPIES : 1) + extra + NGP) + sizeof(int) * NiGP + 16
		                  + ((KOH_UPT > 1) ? sizeof(P) * (size_t)KOH_SLOTS * 2 * NC * UPSP + 16 : 0);
		if (smem > sizeof(g_dynamic_shared)) return 2;
	}

	const char* slice = std::getenv("KOH_SLICE");
	for (int phase = 0; phase < o->Phases; phase++)
	{
		PerBlockArgs<P, UD, NE> args = a;
		if (slice && slice[0] == '1')
		{
			args.SliceMode = 1;  args.SliceFraction = P(0.1);
			launch(args);
			// the queue order of the continuation: decreasing cost of the pilot (the device builds it with a counting sort over logarithmic buckets)
			std::iota(order.begin(), order.begin() + a.Active, 0);
			std::stable_sort(order.begin(), order.begin() + a.Active, [&](int x, int y) { return last[x] > last[y]; });
			args.SliceMode = 2;  args.DispatchOrder = order.data();
		}
		launch(args);
	}
	for (int i = 0; i < NS; i++)
	{
		if (b->steps_attempted) b->steps_attempted[i] += (long long)att[i];
		if (b->steps_accepted)  b->steps_accepted[i]  += (long long)acc[i];
	}
	return 0;
}
