// TEST INFRASTRUCTURE ONLY -- C entry points of the coupled CPU restatement (oracle_coupled_core.h).
#include "oracle_coupled_abi.h"
#include "oracle_coupled_core.h"
#include "oracle_coupled_systems.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

namespace {
using namespace oracle_coupled;

// Pack the original coupling matrix as the reference's SetHost does (Interface.cuh:993-1081).
template <class Sys, class P>
std::vector<P> pack_matrix(const P* orig)
{
	enum { UPS = Sys::UPS, NC = Sys::NC, CBW = Sys::CBW, CCI = Sys::CCI };
	size_t n = (CCI == 0 && CBW == 0) ? (size_t)NC * UPS * UPS : (CCI == 0) ? (size_t)NC * UPS * (2 * CBW + 1) : (CBW > 0) ? (size_t)NC * (2 * CBW + 1) : (size_t)NC * UPS;
	std::vector<P> M(n, 0);
	for (int k = 0; k < NC; k++)
		for (int r = 0; r < UPS; r++)
			for (int c = 0; c < UPS; c++)
			{
				P v = orig[r + c * UPS + k * UPS * UPS];
				if (CCI == 0 && CBW == 0) { M[r + c * UPS + k * UPS * UPS] = v; continue; }
				int mc = c - r + CBW; if (mc >= UPS) mc -= UPS; if (mc < 0) mc += UPS;
				int lim = (CBW > 0) ? 2 * CBW + 1 : UPS;
				if (mc >= lim) continue;                       // outside the band: the reference rejects such entries
				if (CCI == 0) M[r + mc * UPS + k * UPS * lim] = v;
				else if (r == 0 || true) { if (r == 0) M[mc + k * lim] = v; }   // circulant: row 0 defines the matrix
			}
	return M;
}

template <class Sys, class P, int ALG>
void run_range(const oracle_coupled_buffers* b, const oracle_coupled_options* o, int ndo, const P* M, int lo, int hi)
{
	enum { UPS = Sys::UPS, UD = Sys::UD, NC = Sys::NC, NUP = Sys::NUP, NSP = Sys::NSP, NUA = Sys::NUA, NiUA = Sys::NiUA, NSA = Sys::NSA, NiSA = Sys::NiSA };
	const size_t NS = b->NS, NU = NS * UPS;
	P* gTD = (P*)b->TimeDomain; P* gX = (P*)b->ActualState; P* gT = (P*)b->ActualTime; P* gUP = (P*)b->UnitParameters; P* gSP = (P*)b->SystemParameters;
	P* gUA = (P*)b->UnitAccessories; P* gSA = (P*)b->SystemAccessories; P* gCS = (P*)b->CouplingStrength;
	P* gDT = (P*)b->DenseOutputTimeInstances; P* gDS = (P*)b->DenseOutputStates;
	Options<P> po;
	po.dt_initial = (P)o->InitialTimeStep; po.dt_max = (P)o->MaximumTimeStep; po.dt_min = (P)o->MinimumTimeStep; po.grow = (P)o->TimeStepGrowLimit;
	po.shrink = (P)o->TimeStepShrinkLimit; po.dense_min_dt = (P)o->DenseOutputMinimumTimeStep; po.dense_freq = o->DenseOutputSaveFrequency;
	po.rtol = (const P*)b->RelativeTolerance; po.atol = (const P*)b->AbsoluteTolerance; po.ev_tol = (const P*)b->EventTolerance; po.ev_dir = b->EventDirection; po.ndo = ndo;
	std::vector<P> dtime(ndo > 0 ? ndo : 1), dstate(ndo > 0 ? (size_t)ndo * UPS * UD : 1);
	System<Sys, P>* sp = new System<Sys, P>();
	System<Sys, P>& s = *sp;
	s.dense_time = dtime.data(); s.dense_state = dstate.data();
	for (int sid = lo; sid < hi; sid++)
	{
		for (int i = 0; i < 2; i++) s.td[i] = gTD[sid + i * NS];
		s.t = gT[sid];
		for (int i = 0; i < NSP; i++) s.spar[i] = gSP[sid + i * NS];
		for (int i = 0; i < NSA; i++) s.sacc[i] = gSA[sid + i * NS];
		for (int i = 0; i < NiSA; i++) s.isacc[i] = b->IntegerSystemAccessories[sid + i * NS];
		for (int i = 0; i < NC; i++) s.strength[i] = gCS[sid + i * NS];
		for (int u = 0; u < UPS; u++)
		{
			const size_t g = u + (size_t)sid * UPS;
			for (int i = 0; i < UD; i++) s.x[u][i] = gX[g + i * NU];
			for (int i = 0; i < NUP; i++) s.upar[u][i] = gUP[g + i * NU];
			for (int i = 0; i < NUA; i++) s.uacc[u][i] = gUA[g + i * NU];
			for (int i = 0; i < NiUA; i++) s.iuacc[u][i] = b->IntegerUnitAccessories[g + i * NU];
		}
		s.dense_index = b->DenseOutputIndex[sid];
		for (int k = 0; k < ndo; k++)
		{
			dtime[k] = gDT[sid + (size_t)k * NS];
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) dstate[((size_t)k * UPS + u) * UD + i] = gDS[u + (size_t)sid * UPS + i * NU + (size_t)k * NU * UD];
		}
		s.attempted = s.accepted = 0;
		for (int p = 0; p < o->Phases; p++)
			integrate_phase<Sys, P, ALG>(sid, s, (P*)b->GlobalParameters, b->IntegerGlobalParameters, M, b->CouplingIndex, po);
		for (int i = 0; i < 2; i++) gTD[sid + i * NS] = s.td[i];
		gT[sid] = s.t;
		for (int i = 0; i < NSA; i++) gSA[sid + i * NS] = s.sacc[i];
		for (int i = 0; i < NiSA; i++) b->IntegerSystemAccessories[sid + i * NS] = s.isacc[i];
		for (int u = 0; u < UPS; u++)
		{
			const size_t g = u + (size_t)sid * UPS;
			for (int i = 0; i < UD; i++) gX[g + i * NU] = s.x[u][i];
			for (int i = 0; i < NUA; i++) gUA[g + i * NU] = s.uacc[u][i];
			for (int i = 0; i < NiUA; i++) b->IntegerUnitAccessories[g + i * NU] = s.iuacc[u][i];
		}
		b->DenseOutputIndex[sid] = s.dense_index;
		for (int k = 0; k < ndo; k++)
		{
			gDT[sid + (size_t)k * NS] = dtime[k];
			for (int u = 0; u < UPS; u++) for (int i = 0; i < UD; i++) gDS[u + (size_t)sid * UPS + i * NU + (size_t)k * NU * UD] = dstate[((size_t)k * UPS + u) * UD + i];
		}
		if (b->steps_attempted) b->steps_attempted[sid] += s.attempted;
		if (b->steps_accepted) b->steps_accepted[sid] += s.accepted;
	}
	delete sp;
}

template <class Sys, class P, int ALG>
int run_all(const oracle_coupled_buffers* b, const oracle_coupled_options* o, int ndo)
{
	std::vector<P> M = pack_matrix<Sys, P>((const P*)b->CouplingMatrix);
	int active = std::min(o->ActiveSystems, b->NS);
	int nthreads = std::max(1, std::min(o->HostThreads, active));
	if (nthreads == 1) { run_range<Sys, P, ALG>(b, o, ndo, M.data(), 0, active); return 0; }
	std::vector<std::thread> pool;
	for (int w = 0; w < nthreads; w++)
		pool.emplace_back([=, &M]() { for (int s = w; s < active; s += nthreads) run_range<Sys, P, ALG>(b, o, ndo, M.data(), s, s + 1); });
	for (auto& t : pool) t.join();
	return 0;
}

template <class Sys, class P>
int dispatch(int algorithm, int ndo, const oracle_coupled_buffers* b, const oracle_coupled_options* o, oracle_coupled_shape* shape)
{
	if (shape)
	{
		shape->UPS = Sys::UPS; shape->UD = Sys::UD; shape->NC = Sys::NC; shape->CBW = Sys::CBW; shape->CCI = Sys::CCI; shape->NUP = Sys::NUP; shape->NSP = Sys::NSP;
		shape->NGP = Sys::NGP; shape->NiGP = Sys::NiGP; shape->NUA = Sys::NUA; shape->NiUA = Sys::NiUA; shape->NSA = Sys::NSA; shape->NiSA = Sys::NiSA; shape->NE = Sys::NE;
		shape->NDO = ndo; shape->Algorithm = algorithm; shape->Precision = (int)sizeof(P);
		return 0;
	}
	return algorithm == ALG_RK4 ? run_all<Sys, P, ALG_RK4>(b, o, ndo) : run_all<Sys, P, ALG_RKCK45>(b, o, ndo);
}

// P = float instantiates the same restatement in single precision, with the system definitions' double literals
// promoting as they do in the reference's own float instantiation (pinned against the reference's ring kernel in float,
// tests/test_oracle_coupled.py; the reference's bubble tutorial does not compile with Precision = float).
template <class P>
int route(const char* system, int ups, int cbw, int cci, int algorithm, int ndo, const oracle_coupled_buffers* b, const oracle_coupled_options* o, oracle_coupled_shape* shape)
{
	if (!std::strcmp(system, "bubble_ensemble"))
	{
		if (ups == 8)  return dispatch<BubbleEnsemble<P, 8>, P>(algorithm, ndo, b, o, shape);
		if (ups == 16) return dispatch<BubbleEnsemble<P, 16>, P>(algorithm, ndo, b, o, shape);
		if (ups == 32) return dispatch<BubbleEnsemble<P, 32>, P>(algorithm, ndo, b, o, shape);
		if (ups == 64) return dispatch<BubbleEnsemble<P, 64>, P>(algorithm, ndo, b, o, shape);
		if (ups == 128) return dispatch<BubbleEnsemble<P, 128>, P>(algorithm, ndo, b, o, shape);
	}
	if (!std::strcmp(system, "duffing_ring") && ups == 32 && cbw == 0 && cci == 0) return dispatch<DuffingRing<P, 32, 0, 0>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 192 && cbw == 0 && cci == 0) return dispatch<DuffingRing<P, 192, 0, 0>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 50 && cbw == 1 && cci == 1) return dispatch<DuffingRing<P, 50, 1, 1>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 70 && cbw == 2 && cci == 0) return dispatch<DuffingRing<P, 70, 2, 0>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 1500 && cbw == 1 && cci == 1) return dispatch<DuffingRing<P, 1500, 1, 1>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 2500 && cbw == 2 && cci == 0) return dispatch<DuffingRing<P, 2500, 2, 0>, P>(algorithm, ndo, b, o, shape);
	if (!std::strcmp(system, "duffing_ring") && ups == 14)
	{
		if (cbw == 1 && cci == 1) return dispatch<DuffingRing<P, 14, 1, 1>, P>(algorithm, ndo, b, o, shape);
		if (cbw == 1 && cci == 0) return dispatch<DuffingRing<P, 14, 1, 0>, P>(algorithm, ndo, b, o, shape);
		if (cbw == 0 && cci == 1) return dispatch<DuffingRing<P, 14, 0, 1>, P>(algorithm, ndo, b, o, shape);
		if (cbw == 0 && cci == 0) return dispatch<DuffingRing<P, 14, 0, 0>, P>(algorithm, ndo, b, o, shape);
	}
	if (!std::strcmp(system, "swing_ring") && ups == 40 && cbw == 2 && cci == 0) return dispatch<SwingRing<P, 40, 2, 0>, P>(algorithm, ndo, b, o, shape);
	return 1;
}
} // namespace

extern "C" int oracle_coupled_query_shape(const char* system, int ups, int cbw, int cci, int algorithm, int ndo, oracle_coupled_shape* shape)
{ return route<double>(system, ups, cbw, cci, algorithm, ndo, nullptr, nullptr, shape); }

extern "C" int oracle_coupled_run(const char* system, int ups, int cbw, int cci, int algorithm, int ndo, const oracle_coupled_buffers* b, const oracle_coupled_options* o)
{ return route<double>(system, ups, cbw, cci, algorithm, ndo, b, o, nullptr); }

// the same with the precision stated in bytes (8 or 4)
extern "C" int oracle_coupled_run_p(const char* system, int ups, int cbw, int cci, int algorithm, int ndo, int precision, const oracle_coupled_buffers* b, const oracle_coupled_options* o)
{
	if (precision == 8) return route<double>(system, ups, cbw, cci, algorithm, ndo, b, o, nullptr);
	if (precision == 4) return route<float>(system, ups, cbw, cci, algorithm, ndo, b, o, nullptr);
	return 2;
}
