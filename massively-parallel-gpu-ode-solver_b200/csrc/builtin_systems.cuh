// builtin_systems.cuh -- the tutorial systems the BASELINE configurations use, compiled into the
// library so that callers without a CUDA compiler (Python, the bench) can run them by name.
// Each struct supplies the six callbacks of one reference tutorial as static device functions
// with the reference's argument lists; the file:line each one follows is cited.  Floating
// literals take the run's precision (true FP32 in float runs; the reference's sources write
// double literals and therefore mix FP64 into their FP32 instantiation -- see DESIGN.md).
#ifndef MPGOS_B200_BUILTIN_SYSTEMS_CUH
#define MPGOS_B200_BUILTIN_SYSTEMS_CUH

#include <cuda_runtime.h>
#include <math.h>
#include "../../include/mpgos_b200/fast_math.cuh"

#ifndef MPGOS_B200_FM_BUNDLE
#define MPGOS_B200_FM_BUNDLE 1      // A/B switch (tools/build_variant.sh): 0 = the fast-math forms make three separate subroutine calls, as the CUDA-library forms do
#endif

namespace mpgos_b200 {
namespace builtin {

#define MPGOS_ODE_ARGS(P)    int tid, int NT, P* F, P* X, P T, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi
#define MPGOS_EVENT_ARGS(P)  int tid, int NT, P* EF, P T, P dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi
#define MPGOS_EVACT_ARGS(P)  int tid, int NT, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi
#define MPGOS_STEP_ARGS(P)   int tid, int NT, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi
#define MPGOS_INIT_ARGS(P)   int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi
#define MPGOS_DEV            static __device__ __forceinline__ void

template <class P>
struct Empty   // callbacks a system leaves empty
{
	MPGOS_DEV EventFunction(MPGOS_EVENT_ARGS(P)) {}
	MPGOS_DEV ActionAfterEventDetection(MPGOS_EVACT_ARGS(P)) {}
	MPGOS_DEV ActionAfterSuccessfulTimeStep(MPGOS_STEP_ARGS(P)) {}
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P)) {}
	MPGOS_DEV Finalization(MPGOS_INIT_ARGS(P)) {}
};

// Tutorials_SingleSystem_PerThread/T2_Lorenz/Lorenz_SystemDefinition.cuh:6-14
template <class P>
struct Lorenz : Empty<P>
{
	enum { SD = 3, NCP = 1, NSP = 0, NISP = 0, NE = 0, NA = 0, NIA = 0 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P))
	{
		F[0] = P(10.0) * (X[1] - X[0]);
		F[1] = cPAR[0] * X[0] - X[1] - X[0] * X[2];
		F[2] = X[0] * X[1] - P(2.666) * X[2];
	}
};

// Not a tutorial system: relaxation towards tid / NT.  The only built-in system whose right-hand side and accessories depend on
// the instance id and on NT as the callbacks see them (ST/Solver.cuh:148-160 hands both to every callback; README.md:35 of the
// reference advertises indexing by them) -- used to check that a sweep dealt out over several solver parts / devices still
// shows every callback the instance's id in the WHOLE sweep.
template <class P>
struct TidProbe : Empty<P>
{
	enum { SD = 1, NCP = 1, NSP = 0, NISP = 0, NE = 0, NA = 1, NIA = 2 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P)) { F[0] = cPAR[0] * ((P)tid / (P)NT - X[0]); }
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P)) { T = TD[0];  ACCi[0] = tid; }
	MPGOS_DEV Finalization(MPGOS_INIT_ARGS(P)) { ACCi[1] = NT;  ACC[0] = (P)tid / (P)NT; }
};

// T3_PoincareSection/Poincare_SystemDefinition.cuh:6-13, 44-51 -- Duffing without events
template <class P>
struct Poincare : Empty<P>
{
	enum { SD = 2, NCP = 1, NSP = 1, NISP = 0, NE = 0, NA = 0, NIA = 0 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P))
	{
		F[0] = X[1];
		F[1] = X[0] - X[0] * X[0] * X[0] - cPAR[0] * X[1] + sPAR[0] * cos(T);
	}
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P)) { T = TD[0]; }
};

// T1_Reference/Reference_SystemDefinition.cuh:6-76 -- Duffing, 2 events, accessories, dense output
template <class P>
struct Duffing : Empty<P>
{
	enum { SD = 2, NCP = 1, NSP = 1, NISP = 0, NE = 2, NA = 3, NIA = 1 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P))
	{
		F[0] = X[1];
		F[1] = X[0] - X[0] * X[0] * X[0] - cPAR[0] * X[1] + sPAR[0] * cos(T);
	}
	MPGOS_DEV EventFunction(MPGOS_EVENT_ARGS(P))
	{
		EF[0] = X[1];
		EF[1] = X[0];
	}
	MPGOS_DEV ActionAfterEventDetection(MPGOS_EVACT_ARGS(P))
	{
		if (X[0] > ACC[0]) ACC[0] = X[0];
		if (IDX == 1) ACCi[0]++;
		if ((IDX == 1) && (ACCi[0] == 2)) ACC[1] = X[1];
	}
	MPGOS_DEV ActionAfterSuccessfulTimeStep(MPGOS_STEP_ARGS(P))
	{
		if (X[0] > ACC[2]) ACC[2] = X[0];
	}
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P))
	{
		T = TD[0];
		DOIDX = 0;
		ACC[0] = X[0];
		ACC[1] = X[1];
		ACC[2] = X[0];
		ACCi[0] = 0;
	}
};

// T6_ImpactDynamics/ImpactDynamics_SystemDefinition.cuh:7-94 -- pressure relief valve, impact law
template <class P>
struct Impact : Empty<P>
{
	enum { SD = 3, NCP = 1, NSP = 4, NISP = 0, NE = 2, NA = 3, NIA = 2 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P))
	{
		F[0] = X[1];
		F[1] = -sPAR[0] * X[1] - (X[0] + sPAR[1]) + X[2];
		F[2] = sPAR[2] * (cPAR[0] - X[0] * sqrt(X[2]));
	}
	MPGOS_DEV EventFunction(MPGOS_EVENT_ARGS(P))
	{
		EF[0] = X[1];   // Poincare section
		EF[1] = X[0];   // valve seat
	}
	MPGOS_DEV ActionAfterEventDetection(MPGOS_EVACT_ARGS(P))
	{
		if (IDX == 0) ACCi[0]++;
		if ((ACCi[0] == 1) && (IDX == 0)) UDT = 1;
		if (IDX == 1) X[1] = -sPAR[3] * X[1];
	}
	MPGOS_DEV ActionAfterSuccessfulTimeStep(MPGOS_STEP_ARGS(P))
	{
		if (X[0] < ACC[0]) ACC[0] = X[0];
		if (X[0] > ACC[1]) ACC[1] = X[0];
		if (fabs(X[0] - ACC[2]) < P(1e-6))
			ACCi[1]++;
		else
			ACC[2] = X[0];
		if (ACCi[1] == 50) UDT = 1;
	}
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P))
	{
		T = TD[0];
		ACC[0] = X[0];
		ACC[1] = X[0];
		ACCi[0] = 0;
		ACC[2] = X[0];
		ACCi[1] = 0;
	}
};

// T5_LargeTimeScaleDifferences/KellerMiksis_SystemDefinition.cuh:7-84 -- dual-frequency
// Keller-Miksis bubble.  RESET_DOIDX adds T1's `DOIDX = 0` (config C4 adds dense output).  FAST: the same right-hand side with
// pow / sincospi / sin / cos from include/mpgos_b200/fast_math.cuh instead of the CUDA math library (the "_fm" built-in names).
template <class P, bool RESET_DOIDX, bool FAST = false>
struct KellerMiksisT : Empty<P>
{
	using M = fm::Math<FAST>;
	enum { SD = 2, NCP = 13, NSP = 0, NISP = 0, NE = 0, NA = 1, NIA = 0 };
	MPGOS_DEV OdeFunction(MPGOS_ODE_ARGS(P))
	{
		const P PI = P(3.14159265358979323846);
		P rx1, rD, p, s1, c1, s2, c2, N;
		const P D = X[0] - cPAR[9] * X[0] * X[1] + cPAR[4] * cPAR[9];
#if MPGOS_B200_FM_BUNDLE
		if (FAST)
		{
			// fast-math form: the two reciprocals side by side, ONE subroutine for the three transcendental evaluations (fast_math.cuh),
			// and the seven terms of the numerator summed as a tree: the four that need no transcendental ahead of the call, the other
			// three side by side after it (the reference's left-to-right sum is a chain of seven dependent additions)
			M::rcp2_(X[0], D, &rx1, &rD);
			const P w = P(1.0) + cPAR[9] * X[1];
			const P Q = -(cPAR[2] * w) - cPAR[3] * rx1 - cPAR[4] * X[1] * rx1 - P(1.5) * (P(1.0) - cPAR[9] * X[1] * (P(1.0) / P(3.0))) * X[1] * X[1];
			M::pow_sincospi_sincos_(rx1, cPAR[10], P(2.0) * T, P(2.0) * cPAR[11] * PI * T + cPAR[12], &p, &s1, &c1, &s2, &c2);
			const P u = (cPAR[0] + cPAR[1] * X[1]) * p + Q;
			const P v = (cPAR[5] * s1 + cPAR[6] * s2) * w;
			const P g = X[0] * (cPAR[7] * c1 + cPAR[8] * c2);
			N = (u - v) - g;
		}
		else
#endif
		{
			rx1 = M::rcp_(X[0]);
			p = M::pow_(rx1, cPAR[10]);
			M::sincospi_(P(2.0) * T, &s1, &c1);
			M::sincos_(P(2.0) * cPAR[11] * PI * T + cPAR[12], &s2, &c2);
			N = (cPAR[0] + cPAR[1] * X[1]) * p - cPAR[2] * (P(1.0) + cPAR[9] * X[1]) - cPAR[3] * rx1 - cPAR[4] * X[1] * rx1
			  - P(1.5) * (P(1.0) - cPAR[9] * X[1] * (P(1.0) / P(3.0))) * X[1] * X[1]
			  - (cPAR[5] * s1 + cPAR[6] * s2) * (P(1.0) + cPAR[9] * X[1]) - X[0] * (cPAR[7] * c1 + cPAR[8] * c2);
			rD = M::rcp_(D);
		}
		F[0] = X[1];
		F[1] = N * rD;
	}
	MPGOS_DEV ActionAfterSuccessfulTimeStep(MPGOS_STEP_ARGS(P))
	{
		if (X[0] > ACC[0]) ACC[0] = X[0];
	}
	MPGOS_DEV Initialization(MPGOS_INIT_ARGS(P))
	{
		T = TD[0];
		ACC[0] = X[0];
		if (RESET_DOIDX) DOIDX = 0;
	}
};
template <class P> using KellerMiksis        = KellerMiksisT<P, false>;
template <class P> using KellerMiksisDense   = KellerMiksisT<P, true>;
template <class P> using KellerMiksisFm      = KellerMiksisT<P, false, true>;
template <class P> using KellerMiksisDenseFm = KellerMiksisT<P, true, true>;

} // namespace builtin
} // namespace mpgos_b200
#endif
