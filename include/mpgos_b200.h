/* mpgos_b200.h -- C-ABI of the B200-native ensemble ODE integrator.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes.  The reference (MPGOS) is a
 * header-only C++/CUDA template library with no binary interface, so each entry point below
 * names the member function of the reference's host object it replaces
 * (file:line relative to the reference root, SourceCodes/SingleSystem_PerThread_Interface.cuh
 * abbreviated ST/Interface.cuh).  The C++ header include/SingleSystem_PerThread_Interface.cuh
 * re-creates the reference's `ProblemSolver<...>` class on top of these calls.
 *
 * Ownership: a solver handle owns its stream, event, pinned host mirrors and device buffers
 * (as the reference object does, ST/Interface.cuh:399-425) and frees them in
 * mpgos_perthread_destroy.  All mirrors use the reference's layout: structure of arrays,
 * instance index fastest, element (instance, serial) at [instance + serial*NT]
 * (ST/Interface.cuh:712); dense state (instance, comp, point) at
 * [instance + comp*NT + point*NT*SD] (ST/Interface.cuh:814).
 *
 * Threading / ordering: like the reference, not thread-safe per handle; copies and solves are
 * asynchronous on the handle's stream and ordered by it (ST/Interface.cuh:878-1005, 1280-1312).
 *
 * Errors: every function returns 0 on success or a negative mpgos_status; the text of the last
 * error of the calling thread is available through mpgos_last_error().  (The reference prints
 * and calls exit(); the C++ drop-in header reproduces that on top of these return codes.)
 * There is no CPU fallback: without a usable CUDA device mpgos_perthread_create fails with
 * MPGOS_ERR_CUDA.
 */
#ifndef MPGOS_B200_H
#define MPGOS_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPGOS_B200_ABI_VERSION 3

typedef enum mpgos_status
{
	MPGOS_OK            =  0,
	MPGOS_ERR_ARGUMENT  = -1,   /* bad enum value, null pointer, wrong precision */
	MPGOS_ERR_BOUNDS    = -2,   /* index outside the template sizes (reference: BoundCheck, ST/Interface.cuh:566-581) */
	MPGOS_ERR_CUDA      = -3,   /* CUDA runtime error (reference: gpuErrCHK, ST/Interface.cuh:13-22) */
	MPGOS_ERR_MEMORY    = -4,   /* not enough device memory (reference: ST/Interface.cuh:368-396) */
	MPGOS_ERR_UNKNOWN_SYSTEM = -5
} mpgos_status;

/* enum Algorithms, ST/Interface.cuh:33 */
typedef enum mpgos_algorithm { MPGOS_RK4 = 0, MPGOS_RKCK45 = 1 } mpgos_algorithm;

/* enum ListOfVariables, ST/Interface.cuh:35-47 (same order, same values) */
typedef enum mpgos_variable
{
	MPGOS_VAR_All = 0, MPGOS_VAR_TimeDomain, MPGOS_VAR_ActualState, MPGOS_VAR_ActualTime,
	MPGOS_VAR_ControlParameters, MPGOS_VAR_SharedParameters, MPGOS_VAR_IntegerSharedParameters,
	MPGOS_VAR_Accessories, MPGOS_VAR_IntegerAccessories, MPGOS_VAR_DenseOutput,
	MPGOS_VAR_DenseIndex, MPGOS_VAR_DenseTime, MPGOS_VAR_DenseState
} mpgos_variable;

/* enum ListOfSolverOptions, ST/Interface.cuh:49-62 (same order, same values), plus two
 * options of this implementation appended after the reference's last value. */
typedef enum mpgos_option
{
	MPGOS_OPT_ThreadsPerBlock = 0, MPGOS_OPT_InitialTimeStep, MPGOS_OPT_ActiveNumberOfThreads,
	MPGOS_OPT_MaximumTimeStep, MPGOS_OPT_MinimumTimeStep, MPGOS_OPT_TimeStepGrowLimit,
	MPGOS_OPT_TimeStepShrinkLimit, MPGOS_OPT_RelativeTolerance, MPGOS_OPT_AbsoluteTolerance,
	MPGOS_OPT_EventTolerance, MPGOS_OPT_EventDirection, MPGOS_OPT_DenseOutputMinimumTimeStep,
	MPGOS_OPT_DenseOutputSaveFrequency, MPGOS_OPT_PreferSharedMemory,
	MPGOS_OPT_LaneRefill,       /* 0: one thread integrates one instance (reference mapping);
	                               n = 1..32: persistent warps; when a warp ballot counts >= n idle
	                               lanes they take the next n instances from a device-side queue
	                               together (1 = every finished lane at once, 32 = whole warps) */
	MPGOS_OPT_DeviceMessages,   /* 1: device-side printf of the reference's warnings; 0 (default) */
	MPGOS_OPT_DispatchOrder     /* direct mapping: 1 (default): from the second phase on, thread blocks are dispatched in order of
	                               decreasing cost of their instances in the previous phase (longest first: the tail of the launch
	                               is then made of the cheapest blocks); 0: in index order, as the reference */
} mpgos_option;

/* Template arguments of the reference's ProblemSolver (ST/Interface.cuh:117-118) as run-time data.
 * InstanceOffset / GlobalInstances support sharding a sweep over several solver objects (one per
 * GPU): user callbacks still receive the global instance id and the global NT
 * (ST/Solver.cuh:148-160 passes tid and NT to every callback). */
typedef struct mpgos_perthread_desc
{
	int NT;                 /* instances owned by this solver object */
	int SD, NCP, NSP, NISP, NE, NA, NIA, NDO;
	int Algorithm;          /* mpgos_algorithm */
	int PrecisionBytes;     /* 8 = double, 4 = float */
	int Device;             /* CUDA device ordinal (ctor argument, ST/Interface.cuh:309) */
	int InstanceOffset;     /* global id of local instance 0 (0 for an unsharded solver) */
	int GlobalInstances;    /* NT seen by the callbacks (0 = same as NT) */
	/* Tile-cyclic shard of a sweep (filled in by mpgos_perthread_create_sharded for the per-device parts of a multi-device
	 * solver; all 0 for an ordinary solver): the sweep is cut into tiles of ShardTile consecutive instances, tile t belongs to
	 * shard t % ShardCount, and local instance i of shard ShardIndex has the global id
	 *     ((i / ShardTile) * ShardCount + ShardIndex) * ShardTile + i % ShardTile  (+ InstanceOffset),
	 * which is what the callbacks receive as `tid` (ST/Solver.cuh:148-160 passes tid and NT to every callback). */
	int ShardIndex, ShardCount, ShardTile;
} mpgos_perthread_desc;

/* What a kernel launcher receives: device pointers in the run's precision and the resolved
 * options.  Plays the role of the four by-value kernel argument structs of the reference
 * (ST/Interface.cuh:72-115). */
typedef struct mpgos_perthread_launch
{
	int    AbiVersion;      /* MPGOS_B200_ABI_VERSION of the runtime that filled the struct; a launcher built against another one refuses */
	mpgos_perthread_desc desc;
	void*  d_TimeDomain;  void* d_ActualState;  void* d_ActualTime;  void* d_ControlParameters;
	void*  d_SharedParameters;  int* d_IntegerSharedParameters;
	const double* SharedParameterValues;          /* host [NSP]: the values of the last host-to-device copy (small sets travel by value) */
	const int*    IntegerSharedParameterValues;   /* host [NISP] */
	void*  d_Accessories;  int* d_IntegerAccessories;
	int*   d_DenseOutputIndex;  void* d_DenseOutputTimeInstances;  void* d_DenseOutputStates;
	unsigned long long* d_StepsAttempted;   /* [NT] accumulated stepper invocations, never NULL */
	unsigned long long* d_StepsAccepted;    /* [NT] accumulated accepted steps, never NULL */
	unsigned int*       d_WorkQueue;        /* [1] device-side instance queue head (lane refill) */
	unsigned int*       d_LastPhaseSteps;   /* [NT] stepper invocations of the phase just integrated, written at phase end */
	int*                d_DispatchOrder;    /* [NT/32 + 1] scratch: tile ids in dispatch order */
	unsigned int*       d_DispatchScratch;  /* [NT/32 + 1 + 768] scratch: cost bucket per tile, histogram, bucket starts, fill counts */
	int                 DispatchHistory;    /* 1: d_LastPhaseSteps describes the previous phase of these instances and cost order is wanted */
	const double* RelativeTolerance;        /* host [SD] */
	const double* AbsoluteTolerance;        /* host [SD] */
	const double* EventTolerance;           /* host [NE] */
	const int*    EventDirection;           /* host [NE] */
	double InitialTimeStep, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit;
	double DenseOutputMinimumTimeStep;
	int    DenseOutputSaveFrequency;
	int    ThreadsPerBlock;
	int    ActiveNumberOfThreads;
	int    PreferSharedMemory;
	int    LaneRefill;
	int    DeviceMessages;
	int    MultiProcessorCount;
} mpgos_perthread_launch;

/* A launcher enqueues ONE integration phase (the reference's kernel launch,
 * ST/Interface.cuh:1284) on `stream` (a cudaStream_t) and returns a cudaError_t as int.  It is
 * generated by include/mpgos_b200/perthread_kernel.cuh in the translation unit that defines the
 * six device callbacks, or obtained from mpgos_perthread_builtin_launcher(). */
typedef int (*mpgos_perthread_launcher)(const mpgos_perthread_launch* args, void* stream);

typedef struct mpgos_perthread_solver mpgos_perthread_solver;

const char* mpgos_last_error(void);
int   mpgos_abi_version(void);
int   mpgos_device_count(void);                                   /* ListCUDADevices, ST/Interface.cuh:218-229 */

/* ProblemSolver(int), ST/Interface.cuh:309-518: device, stream, event, pinned mirrors, device buffers, default options */
int   mpgos_perthread_create(const mpgos_perthread_desc* desc, mpgos_perthread_solver** out);
/* ~ProblemSolver, ST/Interface.cuh:522-555 */
int   mpgos_perthread_destroy(mpgos_perthread_solver* s);

/* One solver object over SEVERAL devices of one node (the reference's only multi-GPU mechanism is one ProblemSolver per device,
 * filled and driven by hand: Tutorials_SingleSystem_PerThread/T8_MultiGPUSingleNode/MultiGPUSingleNode.cu:55-56, 85-135).
 * desc->NT is the size of the whole sweep; desc->Device is ignored.  The handle behaves like an ordinary one for every call
 * below: ONE pinned host mirror per variable in the reference layout (stride NT, so SetHost/GetHost index arithmetic is
 * unchanged); the instances are dealt to the devices in tiles of `tile` consecutive instances, round-robin (cost varies
 * smoothly along a sweep, so contiguous slabs would leave most devices idle behind the most expensive one; tile <= 0 picks
 * 4096, a multiple of 32 is required); SynchroniseFromHostToDevice / FromDeviceToHost become strided 2-D copies per device
 * on that device's stream, Solve() enqueues one launch per device from the calling thread, SynchroniseSolver waits for all of
 * them, step counters are gathered in sweep order.  Callbacks receive the global instance id and the global NT.  There is no
 * collective and no device-to-device traffic: instances are independent.  mpgos_perthread_device_ptr and _stream refer to the
 * part on the first device. */
int   mpgos_perthread_create_sharded(const mpgos_perthread_desc* desc, const int* devices, int device_count, int tile,
                                     mpgos_perthread_solver** out);
/* Number of devices behind a handle (1 for an ordinary solver) and, for part `shard`, its device, its number of instances and
 * the global id of its local instance `local`. */
int   mpgos_perthread_shard_count(mpgos_perthread_solver* s);
int   mpgos_perthread_shard_info(mpgos_perthread_solver* s, int shard, int* device, int* instances);
int   mpgos_perthread_shard_global_id(mpgos_perthread_solver* s, int shard, int local);
/* The tile map itself (no device needed): instances of part `index` of `count` in a sweep of NT, and the global id of its local
 * instance `local`; -1 on a bad argument. */
int   mpgos_shard_instances(int NT, int tile, int count, int index);
int   mpgos_shard_global_id(int tile, int count, int index, int local);

/* Pinned host mirror of a variable (SetHost/GetHost do their index arithmetic on it inline,
 * ST/Interface.cuh:708-876, 1009-1081); *elements receives its length.  Variables:
 * TimeDomain..IntegerAccessories, DenseIndex, DenseTime, DenseState. */
void* mpgos_perthread_host_ptr(mpgos_perthread_solver* s, int variable, size_t* elements);

/* SolverOption(Option, Value), ST/Interface.cuh:613-664 */
int   mpgos_perthread_set_option(mpgos_perthread_solver* s, int option, double value);
/* SolverOption(Option, Index, Value), ST/Interface.cuh:668-703 */
int   mpgos_perthread_set_option_indexed(mpgos_perthread_solver* s, int option, int index, double value);
int   mpgos_perthread_get_option(mpgos_perthread_solver* s, int option, int index, double* value);

/* SynchroniseFromHostToDevice / SynchroniseFromDeviceToHost, ST/Interface.cuh:878-1005 (asynchronous) */
int   mpgos_perthread_sync_h2d(mpgos_perthread_solver* s, int variable);
int   mpgos_perthread_sync_d2h(mpgos_perthread_solver* s, int variable);

/* Same copies between the device buffer of ONE variable and a caller-owned host array of the same
 * layout and length (pinned or pageable), bypassing the solver's own mirror; asynchronous on the
 * solver's stream when the host array is pinned. */
int   mpgos_perthread_upload(mpgos_perthread_solver* s, int variable, const void* host_src, size_t bytes);
int   mpgos_perthread_download(mpgos_perthread_solver* s, int variable, void* host_dst, size_t bytes);

/* Solve(), ST/Interface.cuh:1280-1285: one integration phase, asynchronous */
int   mpgos_perthread_solve(mpgos_perthread_solver* s, mpgos_perthread_launcher launcher);

/* InsertSynchronisationPoint / SynchroniseSolver / SynchroniseDevice, ST/Interface.cuh:1289-1312 */
int   mpgos_perthread_insert_sync_point(mpgos_perthread_solver* s);
int   mpgos_perthread_sync_solver(mpgos_perthread_solver* s);
int   mpgos_perthread_sync_device(mpgos_perthread_solver* s);

/* Not in the reference (it does not count steps): per-instance stepper
 * invocations and accepted steps accumulated over all Solve() calls since the last reset.
 * Synchronises the solver's stream.  Either output may be NULL. */
int   mpgos_perthread_read_step_counters(mpgos_perthread_solver* s, unsigned long long* attempted, unsigned long long* accepted);
int   mpgos_perthread_reset_step_counters(mpgos_perthread_solver* s);
/* Sum over instances, reduced on the device; synchronises the solver's stream. */
int   mpgos_perthread_total_steps(mpgos_perthread_solver* s, unsigned long long* attempted, unsigned long long* accepted);

/* Device-resident checkpoint of every per-instance buffer (time domain, state, time, control
 * parameters, accessories, dense index; the reference's implicit checkpoint is a round trip
 * through the host mirrors).  save copies device -> device-side snapshot,
 * restore copies it back; both asynchronous on the solver's stream.  Lets a sweep be re-run from
 * the same initial condition without touching the host. */
int   mpgos_perthread_save_state(mpgos_perthread_solver* s);
int   mpgos_perthread_restore_state(mpgos_perthread_solver* s);

/* Raw handles for callers that bracket Solve() with their own CUDA events. */
void* mpgos_perthread_stream(mpgos_perthread_solver* s);     /* cudaStream_t */
int   mpgos_perthread_device(mpgos_perthread_solver* s);
/* Device buffer of a variable (same ids as host_ptr), for callers that keep inputs resident. */
void* mpgos_perthread_device_ptr(mpgos_perthread_solver* s, int variable, size_t* elements);

/* Convenience for timing loops: enqueue `phases` consecutive Solve() calls, bracketed by CUDA
 * events on the solver's stream, wait, and return the elapsed device time in milliseconds. */
int   mpgos_perthread_solve_timed(mpgos_perthread_solver* s, mpgos_perthread_launcher launcher, int phases, float* elapsed_ms);

/* Systems compiled into the library (the tutorial systems the BASELINE configurations use):
 * "lorenz", "poincare", "duffing", "impact", "kellermiksis", "kellermiksis_dense", and "tidprobe" (a test system whose
 * right-hand side depends on the instance id).
 * mpgos_perthread_builtin_desc fills SD..NIA for the named system (NT, NDO, Device are the
 * caller's); mpgos_perthread_builtin_launcher returns NULL for an unknown combination. */
int   mpgos_perthread_builtin_count(void);
const char* mpgos_perthread_builtin_name(int i);
int   mpgos_perthread_builtin_desc(const char* system, mpgos_perthread_desc* desc);
mpgos_perthread_launcher mpgos_perthread_builtin_launcher(const char* system, int algorithm, int precision_bytes, int dense_output);

/* Measurement helper (no counterpart in the reference): sustained FMA throughput of the device in
 * TFLOP/s (2 flops per FMA) for FP64 (precision_bytes 8) or FP32 (4) from a register-resident
 * micro-benchmark of independent FMA chains timed with CUDA events.  This is the denominator of
 * the roofline the bench reports. */
int   mpgos_probe_fma_peak(int device, int precision_bytes, double* tflops);
/* The same for the FP64 tensor-core instruction (mma.sync.m8n8k4.f64, 256 FMAs per warp instruction).  On B200 it has the FP64
 * pipe's FMA rate, no more; the bench takes the larger of the two probes as the FP64 peak. */
int   mpgos_probe_dmma_peak(int device, double* tflops);
/* Test support for include/mpgos_b200/fast_math.cuh: evaluates function 0 (sincos: out0 = sin, out1 = cos), 1 (sincospi) or 2 (pow(x, y):
 * out0) on the device for n host values; ref0 / ref1 receive the CUDA math library's results for the same inputs.  Functions 3, 4, 5:
 * the bundled subroutine pow_sincospi_sincos -- its pow(x, y), its sincospi(x), its sincos(y); 6: rcp2 -- 1/x and 1/y; 7: out = the CUDA
 * library's sin(x), cos(x), ref = its sincos(x). */
int   mpgos_fast_math_probe(int device, int function, int n, const double* x, const double* y, double* out0, double* out1, double* ref0, double* ref1);

#ifdef __cplusplus
}
#endif
#endif
