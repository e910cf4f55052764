// A user-defined COUPLED system written in the style of the reference's coupled tutorials (six
// CoupledSystems_PerBlock_* device callbacks of fixed names, included BEFORE the solver header).
// It is not one of the systems compiled into libmpgos_b200.so.
//
// Ring of second-order phase oscillators (swing equations): for unit r of a system
//     theta_r' = v_r
//     v_r'     = (omega_r - d_r v_r) / m  +  K/m  sum_c A[r][c] sin(n (theta_c - theta_r))
// The sine of the difference is split into two linear couplings of the MPGOS form
// F[index_k] += strength_k * CPF_k[r] * sum_c A_k[r][c] * CPT_k[c]:
//     k = 0: CPT = sin(n theta_c), CPF =  cos(n theta_r) / m
//     k = 1: CPT = cos(n theta_c), CPF = -sin(n theta_r) / m
// unit parameters: omega, d;  system parameter: (unused offset added to omega);  global parameter: m;
// integer global parameter: n.  One event per unit (v crosses zero downwards = local maximum of theta);
// the third such event of unit 0 ends the phase.
#ifndef SWINGRING_SYSTEMDEFINITION_H
#define SWINGRING_SYSTEMDEFINITION_H

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_OdeFunction(
		int sid, int uid, Precision* F, Precision* X, Precision T,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC, Precision* CPT, Precision* CPF)
{
	const Precision rm = 1.0 / gPAR[0];
	const Precision a  = igPAR[0] * X[0];
	const Precision s  = sin(a);
	const Precision c  = cos(a);
	F[0] = X[1];
	F[1] = (uPAR[0] + sPAR[0] - uPAR[1] * X[1]) * rm;
	CPT[0] = s;  CPF[0] = c * rm;
	CPT[1] = c;  CPF[1] = -s * rm;
}

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_EventFunction(
		int sid, int uid, Precision* EF, Precision T, Precision dT, Precision* TD, Precision* X,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC)
{
	EF[0] = X[1];
}

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_ActionAfterEventDetection(
		int sid, int uid, int IDX, int& UDT, Precision& T, Precision& dT, Precision* TD, Precision* X,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC)
{
	iuACC[0]++;
	if (uid == 0 && iuACC[0] == 3)
		UDT = 1;
}

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_ActionAfterSuccessfulTimeStep(
		int sid, int uid, int& UDT, Precision& T, Precision& dT, Precision* TD, Precision* X,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC)
{
	if (fabs(X[1]) > uACC[0])
		uACC[0] = fabs(X[1]);
	if (uid == 0)
	{
		isACC[0]++;
		sACC[0] += dT;
	}
}

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_Initialization(
		int sid, int uid, int& DOIDX, Precision& T, Precision& dT, Precision* TD, Precision* X,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC)
{
	if (uid == 0)
	{
		T = TD[0];
		DOIDX = 0;
		isACC[0] = 0;
		sACC[0] = 0.0;
	}
	uACC[0] = fabs(X[1]);
	iuACC[0] = 0;
}

template <class Precision>
__forceinline__ __device__ void CoupledSystems_PerBlock_Finalization(
		int sid, int uid, int& DOIDX, Precision& T, Precision& dT, Precision* TD, Precision* X,
		Precision* uPAR, Precision* sPAR, Precision* gPAR, int* igPAR,
		Precision* uACC, int* iuACC, Precision* sACC, int* isACC)
{
	if (uid == 0)
	{
		const Precision span = TD[1] - TD[0];   // the next phase continues where this one ended
		TD[0] = T;
		TD[1] = T + span;
	}
}

#endif
