// TEST INFRASTRUCTURE ONLY -- CPU restatement of the coupled tutorial system definitions.
//   BubbleEnsemble : Tutorials_CoupledSystems_PerBlock/T2_BubbleEnsemble/BubbleEnsemble_SystemDefinition.cuh:7-95
//   DuffingRing    : Tutorials_CoupledSystems_PerBlock/T1_RingOfDuffingOscillators/RingOfDuffingOscillators_SystemDefinition.cuh:7-105
//   SwingRing      : examples/swing_ring/SwingRing_SystemDefinition.cuh (the user-defined system of the coupled drop-in example)
// Static members with the argument meaning of the reference's CoupledSystems_PerBlock_* callbacks.
#ifndef MPGOS_B200_ORACLE_COUPLED_SYSTEMS_H
#define MPGOS_B200_ORACLE_COUPLED_SYSTEMS_H

#include <cmath>

namespace oracle_coupled {

static inline void host_sincospi(double x, double* s, double* c)
{
	double r = std::remainder(x, 2.0);
	*s = std::sin(3.14159265358979323846 * r);
	*c = std::cos(3.14159265358979323846 * r);
}

#define OC_ODE_ARGS   int sid, int uid, P* F, P* X, P T, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC, P* CPT, P* CPF
#define OC_EV_ARGS    int sid, int uid, P* EF, P T, P dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define OC_EVACT_ARGS int sid, int uid, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define OC_STEP_ARGS  int sid, int uid, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC
#define OC_INIT_ARGS  int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC

template <class P, int UPS_>
struct BubbleEnsemble
{
	enum { UPS = UPS_, UD = 2, NC = 1, CBW = 0, CCI = 0, NUP = 21, NSP = 0, NGP = 0, NiGP = 0, NUA = 0, NiUA = 0, NSA = 0, NiSA = 0, NE = 0 };
	static void ode_function(OC_ODE_ARGS)
	{
		const double PI = 3.14159265358979323846;
		P rx1 = 1.0 / X[0];
		P p = std::pow(rx1, uPAR[10]);
		P s1, c1;
		{ double s, c; host_sincospi((double)(2.0 * T), &s, &c); s1 = (P)s; c1 = (P)c; }
		P s2 = std::sin(2.0 * uPAR[11] * PI * T + uPAR[12]);
		P c2 = std::cos(2.0 * uPAR[11] * PI * T + uPAR[12]);
		P N, D, rD;
		N = (uPAR[0] + uPAR[1] * X[1]) * p - uPAR[2] * (1.0 + uPAR[9] * X[1]) - uPAR[3] * rx1 - uPAR[4] * X[1] * rx1
		    - 1.5 * (1.0 - uPAR[9] * X[1] / 3.0) * X[1] * X[1] - (uPAR[5] * s1 + uPAR[6] * s2) * (1.0 + uPAR[9] * X[1]) - X[0] * (uPAR[7] * c1 + uPAR[8] * c2);
		D = X[0] - uPAR[9] * X[0] * X[1] + uPAR[4] * uPAR[9];
		rD = 1.0 / D;
		F[0] = X[1];
		F[1] = N * rD;
		CPT[0] = -(2 * X[0] * X[1] * X[1] + X[0] * X[0] * N * rD);
		CPF[0] = rD;
	}
	static void event_function(OC_EV_ARGS) {}
	static void action_after_event_detection(OC_EVACT_ARGS) {}
	static void action_after_successful_time_step(OC_STEP_ARGS) {}
	static void initialization(OC_INIT_ARGS) { if (uid == 0) T = TD[0]; }
	static void finalization(OC_INIT_ARGS) {}
};

template <class P, int UPS_, int CBW_, int CCI_>
struct DuffingRing
{
	enum { UPS = UPS_, UD = 2, NC = 2, CBW = CBW_, CCI = CCI_, NUP = 3, NSP = 0, NGP = 1, NiGP = 0, NUA = 3, NiUA = 1, NSA = 2, NiSA = 1, NE = 1 };
	static void ode_function(OC_ODE_ARGS)
	{
		F[0] = X[1] - gPAR[0] * X[0];
		F[1] = -uPAR[1] * X[1] + X[0] - X[0] * X[0] * X[0] + uPAR[2] * std::cos(uPAR[0] * T) - gPAR[0] * X[1];
		CPT[0] = X[0];
		CPF[0] = 1.0;
		CPT[1] = X[1];
		CPF[1] = 1.0;
	}
	static void event_function(OC_EV_ARGS) { EF[0] = X[1]; }
	static void action_after_event_detection(OC_EVACT_ARGS)
	{
		iuACC[0]++;
		if (iuACC[0] == 2) { UDT = 1; uACC[2] = X[0]; }
	}
	static void action_after_successful_time_step(OC_STEP_ARGS)
	{
		if (uid == 0)
		{
			if (sACC[0] > dT) sACC[0] = dT;
			if (sACC[1] < dT) sACC[1] = dT;
			isACC[0]++;
		}
	}
	static void initialization(OC_INIT_ARGS)
	{
		if (uid == 0) { sACC[0] = dT; sACC[1] = dT; isACC[0] = 0; T = TD[0]; DOIDX = 0; }
		uACC[2] = X[0];
		iuACC[0] = 0;
	}
	static void finalization(OC_INIT_ARGS) {}
};

// examples/swing_ring/SwingRing_SystemDefinition.cuh -- the user-defined system of the coupled drop-in example
template <class P, int UPS_, int CBW_, int CCI_>
struct SwingRing
{
	enum { UPS = UPS_, UD = 2, NC = 2, CBW = CBW_, CCI = CCI_, NUP = 2, NSP = 1, NGP = 1, NiGP = 1, NUA = 1, NiUA = 1, NSA = 1, NiSA = 1, NE = 1 };
	static void ode_function(OC_ODE_ARGS)
	{
		const P rm = 1.0 / gPAR[0];
		const P a = igPAR[0] * X[0];
		const P s = std::sin(a);
		const P c = std::cos(a);
		F[0] = X[1];
		F[1] = (uPAR[0] + sPAR[0] - uPAR[1] * X[1]) * rm;
		CPT[0] = s;  CPF[0] = c * rm;
		CPT[1] = c;  CPF[1] = -s * rm;
	}
	static void event_function(OC_EV_ARGS) { EF[0] = X[1]; }
	static void action_after_event_detection(OC_EVACT_ARGS)
	{
		iuACC[0]++;
		if (uid == 0 && iuACC[0] == 3) UDT = 1;
	}
	static void action_after_successful_time_step(OC_STEP_ARGS)
	{
		if (std::fabs(X[1]) > uACC[0]) uACC[0] = std::fabs(X[1]);
		if (uid == 0) { isACC[0]++; sACC[0] += dT; }
	}
	static void initialization(OC_INIT_ARGS)
	{
		if (uid == 0) { T = TD[0]; DOIDX = 0; isACC[0] = 0; sACC[0] = 0.0; }
		uACC[0] = std::fabs(X[1]);
		iuACC[0] = 0;
	}
	static void finalization(OC_INIT_ARGS)
	{
		if (uid == 0) { const P span = TD[1] - TD[0]; TD[0] = T; TD[1] = T + span; }
	}
};

} // namespace oracle_coupled
#endif
