/* mpgos_b200_coupled.h -- C-ABI of the block-per-system (coupled systems) module.
 *
 * Replaces the host half of the reference's coupled ProblemSolver
 * (SourceCodes/CoupledSystems_PerBlock_Interface.cuh, abbreviated CP/Interface.cuh): one solver object = NS systems of
 * UPS coupled units each, UD state components per unit, NC coupling terms through coupling matrices shared by all
 * systems.  Same conventions as mpgos_b200.h (return codes, mpgos_last_error(), asynchronous on the object's stream,
 * no CPU fallback).
 *
 * Layouts at this boundary (host mirrors == device buffers):
 *   unit scope   (system s, unit u, serial i)   [u + s*UPS + i*NS*UPS]     (the reference pads this index per block
 *                                                 from TPB/SPB, CP/Interface.cuh:875-879; hidden behind SetHost/GetHost)
 *   system scope (system s, serial i)           [s + i*NS]                 (CP/Interface.cuh:918-919)
 *   coupling matrix k, PACKED by type exactly as CP/Interface.cuh:993-1081:
 *       full (CBW=0,CCI=0)            [row + col*UPS + k*UPS*UPS]
 *       banded (CBW>0,CCI=0)          [row + diag*UPS + k*UPS*(2CBW+1)],  diag = (col-row+CBW) mod UPS
 *       circulant banded (CBW>0,CCI=1)[diag + k*(2CBW+1)]
 *       full circulant (CBW=0,CCI=1)  [diag + k*UPS]
 *   dense state (s, u, comp i, point p) [u + s*UPS + i*NS*UPS + p*NS*UPS*UD];  dense time (s, p) [s + p*NS]
 */
#ifndef MPGOS_B200_COUPLED_H
#define MPGOS_B200_COUPLED_H

#include <stddef.h>
#include "mpgos_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* enum ListOfVariables of the coupled module, CP/Interface.cuh:35-53 (same order, same values) */
typedef enum mpgos_coupled_variable
{
	MPGOS_CVAR_All = 0, MPGOS_CVAR_TimeDomain, MPGOS_CVAR_ActualTime, MPGOS_CVAR_ActualState, MPGOS_CVAR_UnitParameters,
	MPGOS_CVAR_SystemParameters, MPGOS_CVAR_GlobalParameters, MPGOS_CVAR_IntegerGlobalParameters, MPGOS_CVAR_UnitAccessories,
	MPGOS_CVAR_IntegerUnitAccessories, MPGOS_CVAR_SystemAccessories, MPGOS_CVAR_IntegerSystemAccessories, MPGOS_CVAR_CouplingMatrix,
	MPGOS_CVAR_CouplingStrength, MPGOS_CVAR_CouplingIndex, MPGOS_CVAR_DenseOutput, MPGOS_CVAR_DenseIndex, MPGOS_CVAR_DenseTime,
	MPGOS_CVAR_DenseState, MPGOS_CVAR_COUNT
} mpgos_coupled_variable;

/* enum ListOfSolverOptions of the coupled module, CP/Interface.cuh:55-68 (same order, same values), plus ours */
typedef enum mpgos_coupled_option
{
	MPGOS_COPT_InitialTimeStep = 0, MPGOS_COPT_ActiveSystems, MPGOS_COPT_MaximumTimeStep, MPGOS_COPT_MinimumTimeStep,
	MPGOS_COPT_TimeStepGrowLimit, MPGOS_COPT_TimeStepShrinkLimit, MPGOS_COPT_RelativeTolerance, MPGOS_COPT_AbsoluteTolerance,
	MPGOS_COPT_EventTolerance, MPGOS_COPT_EventDirection, MPGOS_COPT_DenseOutputSaveFrequency, MPGOS_COPT_DenseOutputMinimumTimeStep,
	MPGOS_COPT_SharedGlobalVariables, MPGOS_COPT_SharedCouplingMatrices,
	MPGOS_COPT_DeviceMessages,
	MPGOS_COPT_TensorCoreCoupling,  /* 1 (default): dense shared coupling matrices go through tensor-core MMA where the shape allows; 0: scalar sums */
	MPGOS_COPT_DispatchOrder        /* 1 (default): the system queue hands out the systems in order of decreasing cost (longest first): from the previous
	                                   phase's step counts, or, in the first phase of a sweep, from a pilot slice (the first tenth of the phase, after which
	                                   every system is paused on an accepted step and continued); only for queues longer than two rounds of slots;
	                                   2: the same for any queue length (tests); 0: in index order, one launch per phase */
} mpgos_coupled_option;

/* Template arguments of the reference's coupled ProblemSolver (CP/Interface.cuh:135-136) as run-time data.  TPB and SPB
 * are accepted for source compatibility; the kernel chooses its own block shape. */
typedef struct mpgos_coupled_desc
{
	int NS, UPS, UD, TPB, SPB, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, NDO;
	int Algorithm;        /* mpgos_algorithm */
	int PrecisionBytes;   /* 8 (4 is not built yet) */
	int Device;
	int SystemOffset;     /* global id of local system 0 when a sweep is sharded over several solver objects */
	/* tile-cyclic shard of a multi-device solver (mpgos_coupled_create_sharded; all 0 otherwise): tiles of ShardTile consecutive
	 * systems, tile t on part t % ShardCount; local system i of part ShardIndex is system
	 * ((i / ShardTile) * ShardCount + ShardIndex) * ShardTile + i % ShardTile (+ SystemOffset) of the sweep -- the `sid` the callbacks get */
	int ShardIndex, ShardCount, ShardTile;
} mpgos_coupled_desc;

typedef struct mpgos_coupled_launch
{
	int    AbiVersion;      /* MPGOS_B200_ABI_VERSION of the runtime that filled the struct; a launcher built against another one refuses */
	mpgos_coupled_desc desc;
	void* d_TimeDomain;  void* d_ActualState;  void* d_ActualTime;  void* d_UnitParameters;  void* d_SystemParameters;
	void* d_GlobalParameters;  int* d_IntegerGlobalParameters;  void* d_UnitAccessories;  int* d_IntegerUnitAccessories;
	void* d_SystemAccessories;  int* d_IntegerSystemAccessories;  void* d_CouplingMatrix;  void* d_CouplingStrength;  int* d_CouplingIndex;
	int*  d_DenseOutputIndex;  void* d_DenseOutputTimeInstances;  void* d_DenseOutputStates;
	unsigned long long* d_StepsAttempted;   /* [NS] */
	unsigned long long* d_StepsAccepted;    /* [NS] */
	unsigned int*       d_WorkQueue;        /* [1] */
	unsigned int*       d_LastPhaseSteps;   /* [NS] stepper invocations of the phase just integrated */
	int*                d_DispatchOrder;    /* [NS] scratch: system ids in queue order */
	unsigned int*       d_DispatchScratch;  /* [NS + 768] scratch of the ordering kernels */
	int                 DispatchHistory;    /* 1: d_LastPhaseSteps describes the previous phase of these systems and cost order is wanted (2: also for short queues) */
	void*               d_SliceState;       /* [NS x 40 bytes] controller state of systems paused at the end of a pilot slice */
	void*               d_SliceParameters;  /* [(NUP x UPS + NSP + NC) x NS] unit parameters, system parameters and coupling strengths as the callbacks
	                                           left them at the pause (allocated by the runtime on the first phase that uses a pilot slice) */
	int                 DispatchPilot;      /* 1: no history: integrate a pilot slice first and order the rest by its step counts (2: also for short queues) */
	size_t CouplingMatrixElements;
	const double* RelativeTolerance;  const double* AbsoluteTolerance;  const double* EventTolerance;  const int* EventDirection;   /* host */
	double InitialTimeStep, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit, DenseOutputMinimumTimeStep;
	int    DenseOutputSaveFrequency, ActiveSystems, DeviceMessages, MultiProcessorCount;
	int    TensorCoreCoupling;
} mpgos_coupled_launch;

typedef int (*mpgos_coupled_launcher)(const mpgos_coupled_launch* args, void* stream);
typedef struct mpgos_coupled_solver mpgos_coupled_solver;

/* ProblemSolver(int) / ~ProblemSolver, CP/Interface.cuh:353-657, 661-700 */
int   mpgos_coupled_create(const mpgos_coupled_desc* desc, mpgos_coupled_solver** out);
int   mpgos_coupled_destroy(mpgos_coupled_solver* s);
/* One solver object over several devices of the node, the twin of mpgos_perthread_create_sharded (mpgos_b200.h): desc->NS systems
 * are dealt to the devices in tiles of `tile` systems (<= 0: 64), one pinned host mirror per variable in the reference layout over
 * all NS systems, strided copies per device, one launch per device from the calling thread, coupling matrices / global parameters /
 * coupling indices replicated.  A single system is never split.  Every other entry point takes the handle as usual. */
int   mpgos_coupled_create_sharded(const mpgos_coupled_desc* desc, const int* devices, int device_count, int tile, mpgos_coupled_solver** out);
int   mpgos_coupled_shard_count(mpgos_coupled_solver* s);
/* pinned host mirror of a variable (CouplingMatrix: the PACKED matrix) and its length */
void* mpgos_coupled_host_ptr(mpgos_coupled_solver* s, int variable, size_t* elements);
/* SetHost(CouplingNumber, CouplingMatrix, Row, Col, Value), CP/Interface.cuh:993-1081: packs by matrix type; entries outside
 * the band are rejected with MPGOS_ERR_BOUNDS as the reference's BoundCheck does */
int   mpgos_coupled_set_matrix_entry(mpgos_coupled_solver* s, int coupling, int row, int col, double value);
int   mpgos_coupled_get_matrix_entry(mpgos_coupled_solver* s, int coupling, int row, int col, double* value);
/* SolverOption, CP/Interface.cuh:704-790 */
int   mpgos_coupled_set_option(mpgos_coupled_solver* s, int option, double value);
int   mpgos_coupled_set_option_indexed(mpgos_coupled_solver* s, int option, int index, double value);
/* SynchroniseFromHostToDevice / DeviceToHost, CP/Interface.cuh:1215-1400 */
int   mpgos_coupled_sync_h2d(mpgos_coupled_solver* s, int variable);
int   mpgos_coupled_sync_d2h(mpgos_coupled_solver* s, int variable);
/* Solve(), CP/Interface.cuh:1949-1965 */
int   mpgos_coupled_solve(mpgos_coupled_solver* s, mpgos_coupled_launcher launcher);
int   mpgos_coupled_solve_timed(mpgos_coupled_solver* s, mpgos_coupled_launcher launcher, int phases, float* elapsed_ms);
/* InsertSynchronisationPoint / SynchroniseSolver / SynchroniseDevice, CP/Interface.cuh:1969-1992 */
int   mpgos_coupled_insert_sync_point(mpgos_coupled_solver* s);
int   mpgos_coupled_sync_solver(mpgos_coupled_solver* s);
int   mpgos_coupled_sync_device(mpgos_coupled_solver* s);
/* additions: copies between caller-owned host arrays (pinned for true asynchrony; unit scope [unit + system*UPS + serial*NS*UPS],
 * system scope [system + serial*NS]; not for CouplingMatrix, which is packed by set_matrix_entry) and the device buffers, on the solver's stream */
int   mpgos_coupled_upload(mpgos_coupled_solver* s, int variable, const void* host_src, size_t bytes);
int   mpgos_coupled_download(mpgos_coupled_solver* s, int variable, void* host_dst, size_t bytes);
/* additions: per-system step counters, device-resident checkpoint, raw stream */
int   mpgos_coupled_read_step_counters(mpgos_coupled_solver* s, unsigned long long* attempted, unsigned long long* accepted);
int   mpgos_coupled_reset_step_counters(mpgos_coupled_solver* s);
int   mpgos_coupled_save_state(mpgos_coupled_solver* s);
int   mpgos_coupled_restore_state(mpgos_coupled_solver* s);
void* mpgos_coupled_stream(mpgos_coupled_solver* s);

/* systems compiled into the library: "bubble_ensemble" (UPS 8, 16, 32, 64, 128), "duffing_ring" (UPS 14; CBW/CCI 1/1, 1/0, 0/1, 0/0; UPS 32 with full matrices) */
int   mpgos_coupled_builtin_desc(const char* system, int ups, int cbw, int cci, mpgos_coupled_desc* desc);
mpgos_coupled_launcher mpgos_coupled_builtin_launcher(const char* system, int ups, int cbw, int cci, int algorithm, int dense_output);   /* double precision */
/* the same with the precision stated (bytes: 8, or 4 for the FP32 instantiations; their dense coupling runs on the tensor cores as 3xTF32) */
mpgos_coupled_launcher mpgos_coupled_builtin_launcher_p(const char* system, int ups, int cbw, int cci, int algorithm, int dense_output, int precision_bytes);

#ifdef __cplusplus
}
#endif
#endif
