// TEST INFRASTRUCTURE ONLY -- the checker's C-ABI, shared by
//   * oracle/_ref/libref_*.so   (the reference's own kernel source compiled for the host), and
//   * oracle/liboracle_perthread.so (our CPU restatement of the same algorithm).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load these libraries.  The product (include/, massively-parallel-gpu-ode-solver_b200/) never does.
#ifndef MPGOS_B200_ORACLE_ABI_H
#define MPGOS_B200_ORACLE_ABI_H

#ifdef __cplusplus
extern "C" {
#endif

// Mirrors Struct_SolverOptions + Struct_ThreadConfiguration of the reference
// (SourceCodes/SingleSystem_PerThread_Interface.cuh:72-77,105-115); always doubles here,
// narrowed to the run's precision by the callee.
typedef struct oracle_perthread_options
{
	double InitialTimeStep;
	double MaximumTimeStep;
	double MinimumTimeStep;
	double TimeStepGrowLimit;
	double TimeStepShrinkLimit;
	double DenseOutputMinimumTimeStep;
	int    DenseOutputSaveFrequency;
	int    NumberOfActiveThreads;   // instances [0, active) are integrated, the rest untouched
	int    Phases;                  // number of consecutive Solve() calls
	int    HostThreads;             // worker threads over instances (instances are independent)
} oracle_perthread_options;

// All arrays use the reference's device layout: structure of arrays, instance index fastest,
// element (instance, serial) at [instance + serial*NT]; dense state (instance, comp, point) at
// [instance + comp*NT + point*NT*SD] (reference Interface.cuh:712, 814).  P is double or float
// as the library was built.  steps_attempted / steps_accepted (length NT, may be NULL)
// accumulate stepper invocations / accepted steps per instance over all phases.
typedef struct oracle_perthread_buffers
{
	int    NT;
	void*  TimeDomain;              // P[2*NT]
	void*  ActualState;             // P[SD*NT]
	void*  ActualTime;              // P[NT]
	void*  ControlParameters;       // P[NCP*NT]
	void*  SharedParameters;        // P[NSP]
	int*   IntegerSharedParameters; // int[NISP]
	void*  Accessories;             // P[NA*NT]
	int*   IntegerAccessories;      // int[NIA*NT]
	void*  RelativeTolerance;       // P[SD]
	void*  AbsoluteTolerance;       // P[SD]
	void*  EventTolerance;          // P[NE]
	int*   EventDirection;          // int[NE]
	int*   DenseOutputIndex;        // int[NT]
	void*  DenseOutputTimeInstances;// P[NDO*NT]
	void*  DenseOutputStates;       // P[NDO*SD*NT]
	long long* steps_attempted;     // [NT] or NULL
	long long* steps_accepted;      // [NT] or NULL
} oracle_perthread_buffers;

// Compile-time shape of the library (template arguments of the reference's ProblemSolver).
typedef struct oracle_perthread_shape
{
	int SD, NCP, NSP, NISP, NE, NA, NIA, NDO;
	int Algorithm;   // 0 = RK4, 1 = RKCK45
	int Precision;   // 8 = double, 4 = float
} oracle_perthread_shape;

// Exported by every libref_*.so:
//   void ref_perthread_shape(oracle_perthread_shape*);
//   int  ref_perthread_run(const oracle_perthread_buffers*, const oracle_perthread_options*);
// Exported by liboracle_perthread.so:
//   int  oracle_perthread_query_shape(const char* system, int algorithm, int precision, int ndo, oracle_perthread_shape*);
//   int  oracle_perthread_run(const char* system, int algorithm, int precision, int ndo,
//                             const oracle_perthread_buffers*, const oracle_perthread_options*);

#ifdef __cplusplus
}
#endif
#endif
