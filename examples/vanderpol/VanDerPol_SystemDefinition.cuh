// A user-written system definition in the reference's plugin format (six device function templates
// of fixed names, found by the solver header through ordinary name lookup).  NOT one of the systems
// compiled into the library: it exercises the drop-in path end to end -- callbacks inlined into the
// kernel instantiated in this translation unit.
//
// Forced Van der Pol oscillator   x0' = x1,   x1' = mu (1 - x0^2) x1 - x0 + A cos(w t)
//   control parameter cPAR[0] = mu; shared parameters sPAR[0] = A, sPAR[1] = w; integer shared sPARi[0] = crossings to count
//   event 0: x0 = 0 downwards (counted; the phase stops after sPARi[0] crossings); event 1: x1 = 0 either way (extrema)
//   accessories: ACC[0] = max x0, ACC[1] = time of the last event-0 crossing; ACCi[0] = crossings, ACCi[1] = accepted steps
#ifndef VANDERPOL_SYSTEMDEFINITION_H
#define VANDERPOL_SYSTEMDEFINITION_H

template <class Precision>
__forceinline__ __device__ void PerThread_OdeFunction(int tid, int NT, Precision* F, Precision* X, Precision T,
                                                      Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	F[0] = X[1];
	F[1] = cPAR[0] * (1.0 - X[0] * X[0]) * X[1] - X[0] + sPAR[0] * cos(sPAR[1] * T);
}

template <class Precision>
__forceinline__ __device__ void PerThread_EventFunction(int tid, int NT, Precision* EF, Precision T, Precision dT, Precision* TD, Precision* X,
                                                        Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	EF[0] = X[0];
	EF[1] = X[1];
}

template <class Precision>
__forceinline__ __device__ void PerThread_ActionAfterEventDetection(int tid, int NT, int IDX, int& UDT, Precision& T, Precision& dT, Precision* TD, Precision* X,
                                                                    Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	if (IDX == 0)
	{
		ACCi[0]++;
		ACC[1] = T;
		if (ACCi[0] == sPARi[0]) UDT = 1;
	}
}

template <class Precision>
__forceinline__ __device__ void PerThread_ActionAfterSuccessfulTimeStep(int tid, int NT, int& UDT, Precision& T, Precision& dT, Precision* TD, Precision* X,
                                                                        Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	if (X[0] > ACC[0]) ACC[0] = X[0];
	ACCi[1]++;
}

template <class Precision>
__forceinline__ __device__ void PerThread_Initialization(int tid, int NT, int& DOIDX, Precision& T, Precision& dT, Precision* TD, Precision* X,
                                                         Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	T = TD[0];
	DOIDX = 0;
	ACC[0] = X[0];
	ACC[1] = TD[0];
	ACCi[0] = 0;
}

template <class Precision>
__forceinline__ __device__ void PerThread_Finalization(int tid, int NT, int& DOIDX, Precision& T, Precision& dT, Precision* TD, Precision* X,
                                                       Precision* cPAR, Precision* sPAR, int* sPARi, Precision* ACC, int* ACCi)
{
	TD[0] = T;     // continue from here in the next phase
}

#endif
