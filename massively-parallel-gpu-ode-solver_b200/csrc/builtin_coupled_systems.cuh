// builtin_coupled_systems.cuh -- the coupled tutorial systems compiled into the library.  Static device functions with
// the argument lists of the reference's CoupledSystems_PerBlock_* callbacks; literals in the run's precision.
//   BubbleEnsemble : Tutorials_CoupledSystems_PerBlock/T2_BubbleEnsemble/BubbleEnsemble_SystemDefinition.cuh:7-95
//   DuffingRing    : Tutorials_CoupledSystems_PerBlock/T1_RingOfDuffingOscillators/RingOfDuffingOscillators_SystemDefinition.cuh:7-105
#ifndef MPGOS_B200_BUILTIN_COUPLED_SYSTEMS_CUH
#define MPGOS_B200_BUILTIN_COUPLED_SYSTEMS_CUH

#include <cuda_runtime.h>
#include <math.h>
#include "../../include/mpgos_b200/fast_math.cuh"

#ifndef MPGOS_B200_FM_BUNDLE
#define MPGOS_B200_FM_BUNDLE 1      // A/B switch (tools/build_variant.sh): 0 = the fast-math forms make three separate subroutine calls, as the CUDA-library forms do
#endif

namespace mpgos_b200 {
namespace builtin {

#define MPGOS_C_ODE_ARGS(P)   int sid, int uid, P* F, P* X, P T, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC, P* CPT, P* CPF
#define MPGOS_C_EV_ARGS(P)    int sid, int uid, P* EF, P T, P dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define MPGOS_C_EVACT_ARGS(P) int sid, int uid, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define MPGOS_C_STEP_ARGS(P)  int sid, int uid, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define MPGOS_C_INIT_ARGS(P)  int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define MPGOS_C_DEV           static __device__ __forceinline__ void

// FAST: pow / sincospi / sin / cos from include/mpgos_b200/fast_math.cuh instead of the CUDA math library ("bubble_ensemble_fm")
template <class P, bool FAST>
struct BubbleEnsembleT
{
	using M = fm::Math<FAST>;
	enum { UD = 2, NC = 1, NUP = 21, NSP = 0, NGP = 0, NiGP = 0, NUA = 0, NiUA = 0, NSA = 0, NiSA = 0, NE = 0 };
	MPGOS_C_DEV OdeFunction(MPGOS_C_ODE_ARGS(P))
	{
		const P PI = P(3.14159265358979323846);
		P rx1, rD, p, s1, c1, s2, c2, N;
		const P D = X[0] - uPAR[9] * X[0] * X[1] + uPAR[4] * uPAR[9];
#if MPGOS_B200_FM_BUNDLE
		if (FAST)
		{
			// fast-math form: the two reciprocals side by side, ONE subroutine for the three transcendental evaluations (fast_math.cuh),
			// and the seven terms of the numerator summed as a tree: the four that need no transcendental ahead of the call, the other
			// three side by side after it (the reference's left-to-right sum is a chain of seven dependent additions)
			M::rcp2_(X[0], D, &rx1, &rD);
			const P w = P(1.0) + uPAR[9] * X[1];
			const P Q = -(uPAR[2] * w) - uPAR[3] * rx1 - uPAR[4] * X[1] * rx1 - P(1.5) * (P(1.0) - uPAR[9] * X[1] / P(3.0)) * X[1] * X[1];
			M::pow_sincospi_sincos_(rx1, uPAR[10], P(2.0) * T, P(2.0) * uPAR[11] * PI * T + uPAR[12], &p, &s1, &c1, &s2, &c2);
			const P u = (uPAR[0] + uPAR[1] * X[1]) * p + Q;
			const P v = (uPAR[5] * s1 + uPAR[6] * s2) * w;
			const P g = X[0] * (uPAR[7] * c1 + uPAR[8] * c2);
			N = (u - v) - g;
		}
		else
#endif
		{
			rx1 = M::rcp_(X[0]);
			p = M::pow_(rx1, uPAR[10]);
			M::sincospi_(P(2.0) * T, &s1, &c1);
			M::sincos_(P(2.0) * uPAR[11] * PI * T + uPAR[12], &s2, &c2);
			N = (uPAR[0] + uPAR[1] * X[1]) * p - uPAR[2] * (P(1.0) + uPAR[9] * X[1]) - uPAR[3] * rx1 - uPAR[4] * X[1] * rx1
			  - P(1.5) * (P(1.0) - uPAR[9] * X[1] / P(3.0)) * X[1] * X[1]
			  - (uPAR[5] * s1 + uPAR[6] * s2) * (P(1.0) + uPAR[9] * X[1]) - X[0] * (uPAR[7] * c1 + uPAR[8] * c2);
			rD = M::rcp_(D);
		}
		F[0] = X[1];
		F[1] = N * rD;
		CPT[0] = -(2 * X[0] * X[1] * X[1] + X[0] * X[0] * N * rD);
		CPF[0] = rD;
	}
	MPGOS_C_DEV EventFunction(MPGOS_C_EV_ARGS(P)) {}
	MPGOS_C_DEV ActionAfterEventDetection(MPGOS_C_EVACT_ARGS(P)) {}
	MPGOS_C_DEV ActionAfterSuccessfulTimeStep(MPGOS_C_STEP_ARGS(P)) {}
	MPGOS_C_DEV Initialization(MPGOS_C_INIT_ARGS(P)) { if (uid == 0) T = TD[0]; }
	MPGOS_C_DEV Finalization(MPGOS_C_INIT_ARGS(P)) {}
};

template <class P> using BubbleEnsemble   = BubbleEnsembleT<P, false>;
template <class P> using BubbleEnsembleFm = BubbleEnsembleT<P, true>;

template <class P>
struct DuffingRing
{
	enum { UD = 2, NC = 2, NUP = 3, NSP = 0, NGP = 1, NiGP = 0, NUA = 3, NiUA = 1, NSA = 2, NiSA = 1, NE = 1 };
	MPGOS_C_DEV OdeFunction(MPGOS_C_ODE_ARGS(P))
	{
		F[0] = X[1] - gPAR[0] * X[0];
		F[1] = -uPAR[1] * X[1] + X[0] - X[0] * X[0] * X[0] + uPAR[2] * cos(uPAR[0] * T) - gPAR[0] * X[1];
		CPT[0] = X[0];
		CPF[0] = P(1.0);
		CPT[1] = X[1];
		CPF[1] = P(1.0);
	}
	MPGOS_C_DEV EventFunction(MPGOS_C_EV_ARGS(P)) { EF[0] = X[1]; }
	MPGOS_C_DEV ActionAfterEventDetection(MPGOS_C_EVACT_ARGS(P))
	{
		iuACC[0]++;
		if (iuACC[0] == 2) { UDT = 1; uACC[2] = X[0]; }
	}
	MPGOS_C_DEV ActionAfterSuccessfulTimeStep(MPGOS_C_STEP_ARGS(P))
	{
		if (uid == 0)
		{
			if (sACC[0] > dT) sACC[0] = dT;
			if (sACC[1] < dT) sACC[1] = dT;
			isACC[0]++;
		}
	}
	MPGOS_C_DEV Initialization(MPGOS_C_INIT_ARGS(P))
	{
		if (uid == 0) { sACC[0] = dT; sACC[1] = dT; isACC[0] = 0; T = TD[0]; DOIDX = 0; }
		uACC[2] = X[0];
		iuACC[0] = 0;
	}
	MPGOS_C_DEV Finalization(MPGOS_C_INIT_ARGS(P)) {}
};

// Not a tutorial system: callbacks that WRITE the parameter arrays (unit parameters in every unit, the system parameter and the
// coupling strength... by unit 0), as the reference allows -- there the arrays live in registers / shared memory for the whole
// phase.  Used to check that pausing a system in mid-phase (pilot slice) and resuming it later loses none of those writes.
template <class P>
struct ParameterDrift
{
	enum { UD = 1, NC = 1, NUP = 1, NSP = 1, NGP = 0, NiGP = 0, NUA = 0, NiUA = 0, NSA = 0, NiSA = 0, NE = 0 };
	MPGOS_C_DEV OdeFunction(MPGOS_C_ODE_ARGS(P))
	{
		F[0] = -uPAR[0] * X[0] + sPAR[0];
		CPT[0] = X[0];
		CPF[0] = P(1.0);
	}
	MPGOS_C_DEV EventFunction(MPGOS_C_EV_ARGS(P)) {}
	MPGOS_C_DEV ActionAfterEventDetection(MPGOS_C_EVACT_ARGS(P)) {}
	MPGOS_C_DEV ActionAfterSuccessfulTimeStep(MPGOS_C_STEP_ARGS(P))
	{
		uPAR[0] *= P(1.0) + P(0.05) * dT;
		if (uid == 0) sPAR[0] += P(0.1) * dT;
	}
	MPGOS_C_DEV Initialization(MPGOS_C_INIT_ARGS(P)) { if (uid == 0) T = TD[0]; }
	MPGOS_C_DEV Finalization(MPGOS_C_INIT_ARGS(P)) {}
};

} // namespace builtin
} // namespace mpgos_b200
#endif
