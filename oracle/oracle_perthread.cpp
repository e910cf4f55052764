// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU restatement (see oracle_perthread_core.h,
// oracle_systems.h, oracle_abi.h).  Dispatches (system, algorithm, precision) to a template
// instantiation and gathers / scatters between the caller's SoA arrays and per-instance state.
#include "oracle_abi.h"
#include "oracle_perthread_core.h"
#include "oracle_systems.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace {

using namespace oracle;

template <class Sys, class P, int ALG>
void run_range(const oracle_perthread_buffers* b, const oracle_perthread_options* o, int ndo, int lo, int hi)
{
	enum { SD = Sys::SD, NCP = Sys::NCP, NA = Sys::NA, NIA = Sys::NIA };
	const int NT = b->NT;
	P* gTD = (P*)b->TimeDomain;  P* gX = (P*)b->ActualState;  P* gT = (P*)b->ActualTime;
	P* gCP = (P*)b->ControlParameters;  P* gACC = (P*)b->Accessories;
	P* gDT = (P*)b->DenseOutputTimeInstances;  P* gDS = (P*)b->DenseOutputStates;

	PhaseOptions<P> po;
	po.dt_initial = (P)o->InitialTimeStep;  po.dt_max = (P)o->MaximumTimeStep;  po.dt_min = (P)o->MinimumTimeStep;
	po.grow_limit = (P)o->TimeStepGrowLimit;  po.shrink_limit = (P)o->TimeStepShrinkLimit;
	po.dense_min_dt = (P)o->DenseOutputMinimumTimeStep;  po.dense_save_frequency = o->DenseOutputSaveFrequency;
	po.rtol = (const P*)b->RelativeTolerance;  po.atol = (const P*)b->AbsoluteTolerance;
	po.ev_tol = (const P*)b->EventTolerance;  po.ev_dir = b->EventDirection;
	po.ndo = ndo;  po.verbose = false;

	std::vector<P> dense_time(ndo > 0 ? ndo : 1), dense_state(ndo > 0 ? ndo * SD : 1);
	Instance<Sys, P> s;
	s.dense_time = dense_time.data();
	s.dense_state = dense_state.data();

	for (int tid = lo; tid < hi; tid++)
	{
		for (int i = 0; i < 2; i++)   s.td[i]   = gTD[tid + i * NT];
		for (int i = 0; i < SD; i++)  s.x[i]    = gX[tid + i * NT];
		for (int i = 0; i < NCP; i++) s.cpar[i] = gCP[tid + i * NT];
		for (int i = 0; i < NA; i++)  s.acc[i]  = gACC[tid + i * NT];
		for (int i = 0; i < NIA; i++) s.acci[i] = b->IntegerAccessories[tid + i * NT];
		s.t = gT[tid];
		s.dense_index = b->DenseOutputIndex[tid];
		for (int k = 0; k < ndo; k++)
		{
			dense_time[k] = gDT[tid + (size_t)k * NT];
			for (int i = 0; i < SD; i++) dense_state[k * SD + i] = gDS[tid + (size_t)i * NT + (size_t)k * NT * SD];
		}
		s.attempted = s.accepted = 0;

		for (int p = 0; p < o->Phases; p++)
			integrate_phase<Sys, P, ALG>(tid, NT, s, (P*)b->SharedParameters, b->IntegerSharedParameters, po);

		for (int i = 0; i < 2; i++)   gTD[tid + i * NT]  = s.td[i];
		for (int i = 0; i < SD; i++)  gX[tid + i * NT]   = s.x[i];
		for (int i = 0; i < NCP; i++) gCP[tid + i * NT]  = s.cpar[i];
		for (int i = 0; i < NA; i++)  gACC[tid + i * NT] = s.acc[i];
		for (int i = 0; i < NIA; i++) b->IntegerAccessories[tid + i * NT] = s.acci[i];
		gT[tid] = s.t;
		b->DenseOutputIndex[tid] = s.dense_index;
		for (int k = 0; k < ndo; k++)
		{
			gDT[tid + (size_t)k * NT] = dense_time[k];
			for (int i = 0; i < SD; i++) gDS[tid + (size_t)i * NT + (size_t)k * NT * SD] = dense_state[k * SD + i];
		}
		if (b->steps_attempted) b->steps_attempted[tid] += s.attempted;
		if (b->steps_accepted)  b->steps_accepted[tid]  += s.accepted;
	}
}

template <class Sys, class P, int ALG>
int run_all(const oracle_perthread_buffers* b, const oracle_perthread_options* o, int ndo)
{
	int active = std::min(o->NumberOfActiveThreads, b->NT);
	int nthreads = std::max(1, std::min(o->HostThreads, active));
	if (nthreads == 1) { run_range<Sys, P, ALG>(b, o, ndo, 0, active); return 0; }
	std::vector<std::thread> pool;
	const int chunk = 64;
	for (int w = 0; w < nthreads; w++)
		pool.emplace_back([=]() {
			for (int lo = w * chunk; lo < active; lo += nthreads * chunk)
				run_range<Sys, P, ALG>(b, o, ndo, lo, std::min(lo + chunk, active));
		});
	for (auto& th : pool) th.join();
	return 0;
}

template <class Sys>
void fill_shape(oracle_perthread_shape* s, int algorithm, int precision, int ndo)
{
	s->SD = Sys::SD; s->NCP = Sys::NCP; s->NSP = Sys::NSP; s->NISP = Sys::NISP;
	s->NE = Sys::NE; s->NA = Sys::NA; s->NIA = Sys::NIA; s->NDO = ndo;
	s->Algorithm = algorithm;  s->Precision = precision < 0 ? -precision : precision;
}

// precision: 8 = double; 4 = float with float literals ("true FP32");
//           -4 = float with double literals (what the reference's FP32 instantiation computes)
template <template <class, class> class Sys>
int dispatch(int algorithm, int precision, int ndo, const oracle_perthread_buffers* b, const oracle_perthread_options* o, oracle_perthread_shape* shape)
{
	if (shape) { fill_shape<Sys<double, double> >(shape, algorithm, precision, ndo); return 0; }
	if (precision == 8)
		return algorithm == ALG_RK4 ? run_all<Sys<double, double>, double, ALG_RK4>(b, o, ndo) : run_all<Sys<double, double>, double, ALG_RKCK45>(b, o, ndo);
	if (precision == 4)
		return algorithm == ALG_RK4 ? run_all<Sys<float, float>, float, ALG_RK4>(b, o, ndo) : run_all<Sys<float, float>, float, ALG_RKCK45>(b, o, ndo);
	if (precision == -4)
		return algorithm == ALG_RK4 ? run_all<Sys<float, double>, float, ALG_RK4>(b, o, ndo) : run_all<Sys<float, double>, float, ALG_RKCK45>(b, o, ndo);
	return 2;
}

template <class P, class L> using KellerMiksisPlain = KellerMiksis<P, L, false>;
template <class P, class L> using KellerMiksisDense = KellerMiksis<P, L, true>;

int route(const char* system, int algorithm, int precision, int ndo, const oracle_perthread_buffers* b, const oracle_perthread_options* o, oracle_perthread_shape* shape)
{
	if (algorithm != ALG_RK4 && algorithm != ALG_RKCK45) return 3;
	if (!std::strcmp(system, "lorenz"))             return dispatch<Lorenz>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "poincare"))           return dispatch<Poincare>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing"))            return dispatch<Duffing>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "impact"))             return dispatch<Impact>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "kellermiksis"))       return dispatch<KellerMiksisPlain>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "kellermiksis_dense")) return dispatch<KellerMiksisDense>(algorithm, precision, ndo, b, o, shape);
	if (!std::strcmp(system, "vanderpol"))          return dispatch<VanDerPol>(algorithm, precision, ndo, b, o, shape);
	return 1;
}

} // namespace

extern "C" int oracle_perthread_query_shape(const char* system, int algorithm, int precision, int ndo, oracle_perthread_shape* shape)
{
	return route(system, algorithm, precision, ndo, nullptr, nullptr, shape);
}

extern "C" int oracle_perthread_run(const char* system, int algorithm, int precision, int ndo,
                                    const oracle_perthread_buffers* b, const oracle_perthread_options* o)
{
	return route(system, algorithm, precision, ndo, b, o, nullptr);
}
