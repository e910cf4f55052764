// TEST INFRASTRUCTURE ONLY -- checker ABI of the coupled (block-per-system) module, shared by
//   * oracle/_ref/libref_coupled_*.so     : the reference's own coupled kernel source run on the host with one
//                                           std::thread per CUDA thread (oracle/ref_coupled_harness.cpp),
//   * oracle/_ref/libref_gpu_coupled_*.so : the reference compiled by nvcc for sm_100a,
//   * oracle/liboracle_coupled.so         : our CPU restatement (oracle_coupled_core.h).
// Layout of every array (all three and the product use it at this boundary):
//   unit scope    (system s, unit u, serial i)  ->  [u + s*UPS + i*NS*UPS]
//   system scope  (system s, serial i)          ->  [s + i*NS]
//   coupling matrix k, row r, col c (ORIGINAL, unpacked)  ->  [r + c*UPS + k*UPS*UPS]
//   dense state (s, u, comp i, point p) -> [u + s*UPS + i*NS*UPS + p*NS*UPS*UD];  dense time (s, p) -> [s + p*NS]
#ifndef MPGOS_B200_ORACLE_COUPLED_ABI_H
#define MPGOS_B200_ORACLE_COUPLED_ABI_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_coupled_shape
{
	int UPS, UD, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, NDO;
	int Algorithm;   // 0 RK4, 1 RKCK45
	int Precision;   // bytes
} oracle_coupled_shape;

typedef struct oracle_coupled_options
{
	double InitialTimeStep, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit, DenseOutputMinimumTimeStep;
	int    DenseOutputSaveFrequency;
	int    ActiveSystems;
	int    Phases;
	int    HostThreads;
} oracle_coupled_options;

typedef struct oracle_coupled_buffers
{
	int    NS;
	void*  TimeDomain;               // P[2*NS]
	void*  ActualState;              // P[UD*NS*UPS]
	void*  ActualTime;               // P[NS]
	void*  UnitParameters;           // P[NUP*NS*UPS]
	void*  SystemParameters;         // P[NSP*NS]
	void*  GlobalParameters;         // P[NGP]
	int*   IntegerGlobalParameters;  // int[NiGP]
	void*  UnitAccessories;          // P[NUA*NS*UPS]
	int*   IntegerUnitAccessories;   // int[NiUA*NS*UPS]
	void*  SystemAccessories;        // P[NSA*NS]
	int*   IntegerSystemAccessories; // int[NiSA*NS]
	void*  CouplingMatrix;           // P[NC*UPS*UPS]  original (unpacked) form
	void*  CouplingStrength;         // P[NC*NS]
	int*   CouplingIndex;            // int[NC]
	void*  RelativeTolerance;        // P[UD]
	void*  AbsoluteTolerance;        // P[UD]
	void*  EventTolerance;           // P[NE]
	int*   EventDirection;           // int[NE]
	int*   DenseOutputIndex;         // int[NS]
	void*  DenseOutputTimeInstances; // P[NDO*NS]
	void*  DenseOutputStates;        // P[NDO*UD*NS*UPS]
	long long* steps_attempted;      // [NS] or NULL
	long long* steps_accepted;       // [NS] or NULL
} oracle_coupled_buffers;

#ifdef __cplusplus
}
#endif
#endif
