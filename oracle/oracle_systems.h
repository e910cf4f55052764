// TEST INFRASTRUCTURE ONLY -- CPU restatement of the tutorial system definitions the BASELINE
// configurations use.  Each struct restates the six user callbacks of one reference tutorial as
// static members (same argument meaning as the reference's PerThread_* functions), citing the
// file:line it follows.  `L` is the type floating literals take: the reference writes plain
// double literals, so its FP32 instantiation silently evaluates those sub-expressions in double
// (L = double reproduces that bit for bit); L = P gives the "true FP32" arithmetic the CUDA
// built-in systems use.
#ifndef MPGOS_B200_ORACLE_SYSTEMS_H
#define MPGOS_B200_ORACLE_SYSTEMS_H

#include <cmath>

namespace oracle {

static inline void host_sincospi(double x, double* s, double* c)
{
	double r = std::remainder(x, 2.0);
	*s = std::sin(3.14159265358979323846 * r);
	*c = std::cos(3.14159265358979323846 * r);
}

// Callbacks a system does not need.
template <class P>
struct NoCallbacks
{
	static void event_function(int, int, P*, P, P, P*, P*, P*, P*, int*, P*, int*) {}
	static void action_after_event_detection(int, int, int, int&, P&, P&, P*, P*, P*, P*, int*, P*, int*) {}
	static void action_after_successful_time_step(int, int, int&, P&, P&, P*, P*, P*, P*, int*, P*, int*) {}
	static void initialization(int, int, int&, P&, P&, P*, P*, P*, P*, int*, P*, int*) {}
	static void finalization(int, int, int&, P&, P&, P*, P*, P*, P*, int*, P*, int*) {}
};

// Tutorials_SingleSystem_PerThread/T2_Lorenz/Lorenz_SystemDefinition.cuh:6-14 (all other callbacks empty)
template <class P, class L>
struct Lorenz : NoCallbacks<P>
{
	enum { SD = 3, NCP = 1, NSP = 0, NISP = 0, NE = 0, NA = 0, NIA = 0 };
	static void ode_function(int, int, P* F, P* X, P, P* cPAR, P*, int*, P*, int*)
	{
		F[0] = L(10.0) * (X[1] - X[0]);
		F[1] = cPAR[0] * X[0] - X[1] - X[0] * X[2];
		F[2] = X[0] * X[1] - L(2.666) * X[2];
	}
};

// T3_PoincareSection/Poincare_SystemDefinition.cuh:6-13 (RHS), :44-51 (Initialization); no events
template <class P, class L>
struct Poincare : NoCallbacks<P>
{
	enum { SD = 2, NCP = 1, NSP = 1, NISP = 0, NE = 0, NA = 0, NIA = 0 };
	static void ode_function(int, int, P* F, P* X, P T, P* cPAR, P* sPAR, int*, P*, int*)
	{
		F[0] = X[1];
		F[1] = X[0] - X[0] * X[0] * X[0] - cPAR[0] * X[1] + sPAR[0] * std::cos(T);
	}
	static void initialization(int, int, int&, P& T, P&, P* TD, P*, P*, P*, int*, P*, int*) { T = TD[0]; }
};

// T1_Reference/Reference_SystemDefinition.cuh:6-76
template <class P, class L>
struct Duffing : NoCallbacks<P>
{
	enum { SD = 2, NCP = 1, NSP = 1, NISP = 0, NE = 2, NA = 3, NIA = 1 };
	static void ode_function(int, int, P* F, P* X, P T, P* cPAR, P* sPAR, int*, P*, int*)
	{
		F[0] = X[1];
		F[1] = X[0] - X[0] * X[0] * X[0] - cPAR[0] * X[1] + sPAR[0] * std::cos(T);
	}
	static void event_function(int, int, P* EF, P, P, P*, P* X, P*, P*, int*, P*, int*)
	{
		EF[0] = X[1];
		EF[1] = X[0];
	}
	static void action_after_event_detection(int, int, int IDX, int&, P&, P&, P*, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		if (X[0] > ACC[0]) ACC[0] = X[0];
		if (IDX == 1) ACCi[0]++;
		if ((IDX == 1) && (ACCi[0] == 2)) ACC[1] = X[1];
	}
	static void action_after_successful_time_step(int, int, int&, P&, P&, P*, P* X, P*, P*, int*, P* ACC, int*)
	{
		if (X[0] > ACC[2]) ACC[2] = X[0];
	}
	static void initialization(int, int, int& DOIDX, P& T, P&, P* TD, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		T = TD[0];
		DOIDX = 0;
		ACC[0] = X[0];
		ACC[1] = X[1];
		ACC[2] = X[0];
		ACCi[0] = 0;
	}
};

// T6_ImpactDynamics/ImpactDynamics_SystemDefinition.cuh:7-94 (pressure relief valve with impact law)
template <class P, class L>
struct Impact : NoCallbacks<P>
{
	enum { SD = 3, NCP = 1, NSP = 4, NISP = 0, NE = 2, NA = 3, NIA = 2 };
	static void ode_function(int, int, P* F, P* X, P, P* cPAR, P* sPAR, int*, P*, int*)
	{
		F[0] = X[1];
		F[1] = -sPAR[0] * X[1] - (X[0] + sPAR[1]) + X[2];
		F[2] = sPAR[2] * (cPAR[0] - X[0] * std::sqrt(X[2]));
	}
	static void event_function(int, int, P* EF, P, P, P*, P* X, P*, P*, int*, P*, int*)
	{
		EF[0] = X[1];   // Poincare section
		EF[1] = X[0];   // impact detection
	}
	static void action_after_event_detection(int, int, int IDX, int& UDT, P&, P&, P*, P* X, P*, P* sPAR, int*, P*, int* ACCi)
	{
		if (IDX == 0) ACCi[0]++;
		if ((ACCi[0] == 1) && (IDX == 0)) UDT = 1;
		if (IDX == 1) X[1] = -sPAR[3] * X[1];
	}
	static void action_after_successful_time_step(int, int, int& UDT, P&, P&, P*, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		if (X[0] < ACC[0]) ACC[0] = X[0];
		if (X[0] > ACC[1]) ACC[1] = X[0];
		if (std::abs(X[0] - ACC[2]) < L(1e-6))
			ACCi[1]++;
		else
			ACC[2] = X[0];
		if (ACCi[1] == 50) UDT = 1;
	}
	static void initialization(int, int, int&, P& T, P&, P* TD, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		T = TD[0];
		ACC[0] = X[0];
		ACC[1] = X[0];
		ACCi[0] = 0;
		ACC[2] = X[0];
		ACCi[1] = 0;
	}
};

// T5_LargeTimeScaleDifferences/KellerMiksis_SystemDefinition.cuh:7-84.  RESET_DOIDX adds the
// `DOIDX = 0` of T1's Initialization (config C4 adds dense output to this tutorial).
template <class P, class L, bool RESET_DOIDX>
struct KellerMiksis : NoCallbacks<P>
{
	enum { SD = 2, NCP = 13, NSP = 0, NISP = 0, NE = 0, NA = 1, NIA = 0 };
	static void ode_function(int, int, P* F, P* X, P T, P* cPAR, P*, int*, P*, int*)
	{
		const L PI = L(3.14159265358979323846);
		P rx1 = L(1.0) / X[0];
		P p = std::pow(rx1, cPAR[10]);
		P s1, c1;
		{ double s, c; host_sincospi((double)(L(2.0) * T), &s, &c); s1 = (P)s; c1 = (P)c; }
		P s2 = std::sin(L(2.0) * cPAR[11] * PI * T + cPAR[12]);
		P c2 = std::cos(L(2.0) * cPAR[11] * PI * T + cPAR[12]);
		P N, D, rD;
		N = (cPAR[0] + cPAR[1] * X[1]) * p - cPAR[2] * (L(1.0) + cPAR[9] * X[1]) - cPAR[3] * rx1 - cPAR[4] * X[1] * rx1
		    - L(1.5) * (L(1.0) - cPAR[9] * X[1] * (L(1.0) / L(3.0))) * X[1] * X[1]
		    - (cPAR[5] * s1 + cPAR[6] * s2) * (L(1.0) + cPAR[9] * X[1]) - X[0] * (cPAR[7] * c1 + cPAR[8] * c2);
		D = X[0] - cPAR[9] * X[0] * X[1] + cPAR[4] * cPAR[9];
		rD = L(1.0) / D;
		F[0] = X[1];
		F[1] = N * rD;
	}
	static void action_after_successful_time_step(int, int, int&, P&, P&, P*, P* X, P*, P*, int*, P* ACC, int*)
	{
		if (X[0] > ACC[0]) ACC[0] = X[0];
	}
	static void initialization(int, int, int& DOIDX, P& T, P&, P* TD, P* X, P*, P*, int*, P* ACC, int*)
	{
		T = TD[0];
		ACC[0] = X[0];
		if (RESET_DOIDX) DOIDX = 0;
	}
};

// examples/vanderpol/VanDerPol_SystemDefinition.cuh -- a user-written system (not from the reference)
// used by the drop-in test: forced Van der Pol oscillator with two events, accessories, integer
// shared parameter, dense output and a Finalization that moves the time domain.
template <class P, class L>
struct VanDerPol : NoCallbacks<P>
{
	enum { SD = 2, NCP = 1, NSP = 2, NISP = 1, NE = 2, NA = 2, NIA = 2 };
	static void ode_function(int, int, P* F, P* X, P T, P* cPAR, P* sPAR, int*, P*, int*)
	{
		F[0] = X[1];
		F[1] = cPAR[0] * (L(1.0) - X[0] * X[0]) * X[1] - X[0] + sPAR[0] * std::cos(sPAR[1] * T);
	}
	static void event_function(int, int, P* EF, P, P, P*, P* X, P*, P*, int*, P*, int*)
	{
		EF[0] = X[0];
		EF[1] = X[1];
	}
	static void action_after_event_detection(int, int, int IDX, int& UDT, P& T, P&, P*, P*, P*, P*, int* sPARi, P* ACC, int* ACCi)
	{
		if (IDX == 0)
		{
			ACCi[0]++;
			ACC[1] = T;
			if (ACCi[0] == sPARi[0]) UDT = 1;
		}
	}
	static void action_after_successful_time_step(int, int, int&, P&, P&, P*, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		if (X[0] > ACC[0]) ACC[0] = X[0];
		ACCi[1]++;
	}
	static void initialization(int, int, int& DOIDX, P& T, P&, P* TD, P* X, P*, P*, int*, P* ACC, int* ACCi)
	{
		T = TD[0];
		DOIDX = 0;
		ACC[0] = X[0];
		ACC[1] = TD[0];
		ACCi[0] = 0;
	}
	static void finalization(int, int, int&, P& T, P&, P* TD, P*, P*, P*, int*, P*, int*) { TD[0] = T; }
};

} // namespace oracle
#endif
