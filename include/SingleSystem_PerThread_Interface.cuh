// SingleSystem_PerThread_Interface.cuh -- source-level drop-in for the MPGOS header of the same name.
//
// A user program written against the reference keeps working unchanged:
//
//     #include "MySystem_SystemDefinition.cuh"          // the six PerThread_* device callbacks
//     #include "SingleSystem_PerThread_Interface.cuh"   // this file (include order as in the reference)
//     ProblemSolver<NT,SD,NCP,NSP,NISP,NE,NA,NIA,NDO,RKCK45,double> Solver(DeviceNumber);
//     Solver.SolverOption(...); Solver.SetHost(...); Solver.SynchroniseFromHostToDevice(All);
//     Solver.Solve(); Solver.InsertSynchronisationPoint(); Solver.SynchroniseSolver(); ...
//
// Same unscoped enums, same template signature, same public methods, same file formats as
// SourceCodes/SingleSystem_PerThread_Interface.cuh:33-208 of the reference.  Behind the surface
// nothing is shared with it: the host object is a thin shell over the C-ABI of libmpgos_b200.so
// (include/mpgos_b200.h: buffers, stream, copies, options) and Solve() hands that runtime a
// launcher for OUR kernel (include/mpgos_b200/perthread_kernel.cuh) instantiated here, in the
// user's translation unit, so that the callbacks inline into it.  Link with -lmpgos_b200.
//
// Deliberate differences, all documented in DESIGN.md:
//   * ThreadsPerBlock is a hint (rounded to a multiple of 32, at most 256) and PreferSharedMemory
//     is accepted and ignored: shared parameters always live in shared memory;
//   * callbacks receive the run-time instance count through `NT` exactly as before, but the
//     device-side warnings of the reference are printed only after
//     SolverOption(DeviceMessages, 1);
//   * extra options LaneRefill / DeviceMessages / DispatchOrder and the StepCounters() accessor.
#ifndef SINGLESYSTEM_PERTHREAD_INTERFACE_H
#define SINGLESYSTEM_PERTHREAD_INTERFACE_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "mpgos_b200.h"
#include "mpgos_b200/perthread_kernel.cuh"

// ---- the reference's unscoped enums (Interface.cuh:33-62); values equal the C-ABI's ------------
enum Algorithms { RK4, RKCK45 };

enum ListOfVariables { All, TimeDomain, ActualState, ActualTime, ControlParameters, SharedParameters, IntegerSharedParameters,
                       Accessories, IntegerAccessories, DenseOutput, DenseIndex, DenseTime, DenseState };

enum ListOfSolverOptions { ThreadsPerBlock, InitialTimeStep, ActiveNumberOfThreads, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit,
                           TimeStepShrinkLimit, RelativeTolerance, AbsoluteTolerance, EventTolerance, EventDirection,
                           DenseOutputMinimumTimeStep, DenseOutputSaveFrequency, PreferSharedMemory,
                           LaneRefill, DeviceMessages, DispatchOrder };

inline std::string VariablesToString(ListOfVariables v)
{
	static const char* const names[] = { "All", "TimeDomain", "ActualState", "ActualTime", "ControlParameters", "SharedParameters",
	                                     "IntegerSharedParameters", "Accessories", "IntegerAccessories", "DenseOutput", "DenseIndex",
	                                     "DenseTime", "DenseState" };
	return ((int)v >= 0 && (int)v <= (int)DenseState) ? names[(int)v] : "Non-existent variable!";
}

inline std::string SolverOptionsToString(ListOfSolverOptions o)
{
	static const char* const names[] = { "ThreadsPerBlock", "InitialTimeStep", "ActiveNumberOfThreads", "MaximumTimeStep", "MinimumTimeStep",
	                                     "TimeStepGrowLimit", "TimeStepShrinkLimit", "RelativeTolerance", "AbsoluteTolerance", "EventTolerance",
	                                     "EventDirection", "DenseOutputMinimumTimeStep", "DenseOutputSaveFrequency", "PreferSharedMemory",
	                                     "LaneRefill", "DeviceMessages", "DispatchOrder" };
	return ((int)o >= 0 && (int)o <= (int)DispatchOrder) ? names[(int)o] : "Non-existent solver option!";
}

// ---- device discovery helpers (Interface.cuh:218-304) ---------------------------------------------
inline void ListCUDADevices()
{
	int n = 0;
	cudaGetDeviceCount(&n);
	for (int i = 0; i < n; i++)
	{
		cudaDeviceProp p;
		cudaGetDeviceProperties(&p, i);
		std::cout << "Device number: " << i << "  name: " << p.name << "  compute capability: " << p.major << "." << p.minor
		          << "  multiprocessors: " << p.multiProcessorCount << "  global memory: " << (p.totalGlobalMem >> 20) << " Mb" << std::endl;
	}
	std::cout << std::endl;
}

inline int SelectDeviceByClosestRevision(int MajorRevision, int MinorRevision)
{
	cudaDeviceProp wanted;
	std::memset(&wanted, 0, sizeof(wanted));
	wanted.major = MajorRevision;
	wanted.minor = MinorRevision;
	int device = 0;
	cudaChooseDevice(&device, &wanted);
	return device;
}

inline void PrintPropertiesOfSpecificDevice(int Device)
{
	cudaDeviceProp p;
	if (cudaGetDeviceProperties(&p, Device) != cudaSuccess) { std::cerr << "ERROR: no CUDA device " << Device << std::endl; exit(EXIT_FAILURE); }
	std::cout << "Selected device number: " << Device << "\n  name: " << p.name << "\n  compute capability: " << p.major << "." << p.minor
	          << "\n  multiprocessors: " << p.multiProcessorCount << "\n  global memory: " << (p.totalGlobalMem >> 20) << " Mb"
	          << "\n  shared memory per block: " << (p.sharedMemPerBlock >> 10) << " Kb\n  registers per block: " << p.regsPerBlock
	          << "\n  warp size: " << p.warpSize << "\n  L2 cache: " << (p.l2CacheSize >> 20) << " Mb" << std::endl << std::endl;
}

// ---- binding of the user's callbacks ----------------------------------------------------------------
// The kernel takes the callbacks through a policy type; this one forwards to the functions of the
// fixed names the user defined BEFORE including this header (found by ordinary lookup here, exactly
// as the reference finds them from its kernel, Solver.cuh:148-398).
namespace mpgos_b200 {
template <class P>
struct UserCallbacks
{
	static __device__ __forceinline__ void OdeFunction(int tid, int NT, P* F, P* X, P T, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_OdeFunction(tid, NT, F, X, T, cPAR, sPAR, sPARi, ACC, ACCi); }
	static __device__ __forceinline__ void EventFunction(int tid, int NT, P* EF, P T, P dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_EventFunction(tid, NT, EF, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static __device__ __forceinline__ void ActionAfterEventDetection(int tid, int NT, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_ActionAfterEventDetection(tid, NT, IDX, UDT, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static __device__ __forceinline__ void ActionAfterSuccessfulTimeStep(int tid, int NT, int& UDT, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_ActionAfterSuccessfulTimeStep(tid, NT, UDT, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static __device__ __forceinline__ void Initialization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_Initialization(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
	static __device__ __forceinline__ void Finalization(int tid, int NT, int& DOIDX, P& T, P& dT, P* TD, P* X, P* cPAR, P* sPAR, int* sPARi, P* ACC, int* ACCi)
	{ PerThread_Finalization(tid, NT, DOIDX, T, dT, TD, X, cPAR, sPAR, sPARi, ACC, ACCi); }
};
}

// ---- the host object ---------------------------------------------------------------------------------
template <int NT, int SD, int NCP, int NSP, int NISP, int NE, int NA, int NIA, int NDO, Algorithms Algorithm, class Precision>
class ProblemSolver
{
	mpgos_perthread_solver* Handle;
	// pinned host mirrors owned by the runtime (reference: h_* members, Interface.cuh:147-157)
	Precision *h_TimeDomain, *h_ActualState, *h_ActualTime, *h_ControlParameters, *h_SharedParameters, *h_Accessories;
	Precision *h_DenseOutputTimeInstances, *h_DenseOutputStates;
	int       *h_IntegerSharedParameters, *h_IntegerAccessories, *h_DenseOutputIndex;

	using Launcher = mpgos_b200::PerThreadLauncher<mpgos_b200::UserCallbacks<Precision>, Precision, SD, NCP, NSP, NISP, NE, NA, NIA, (NDO > 0), (int)Algorithm>;

	[[noreturn]] static void Fatal(const char* where)
	{
		std::cerr << "ERROR: In solver member function " << where << "!" << std::endl << "       " << mpgos_last_error() << std::endl;
		exit(EXIT_FAILURE);
	}
	static void Check(int rc, const char* where) { if (rc != MPGOS_OK) Fatal(where); }

	// reference: BoundCheck, Interface.cuh:566-581 (message on std::cerr, then exit)
	static void BoundCheck(const char* Function, const std::string& Variable, int Value, int LowerLimit, int UpperLimit)
	{
		if (Value >= LowerLimit && Value <= UpperLimit) return;
		std::cerr << "ERROR: In solver member function " << Function << "!" << std::endl << "       Option: " << Variable << std::endl;
		if (LowerLimit > UpperLimit) std::cerr << "       Acceptable index: none" << std::endl;
		else                         std::cerr << "       Acceptable index: " << LowerLimit << "-" << UpperLimit << std::endl;
		std::cerr << "       Current index:    " << Value << std::endl;
		exit(EXIT_FAILURE);
	}
	[[noreturn]] static void NotApplicable(const char* Function, const std::string& Name)
	{
		std::cerr << "ERROR: In solver member function " << Function << "!" << std::endl << "       Option: " << Name << std::endl
		          << "       This option needs different argument configuration or not applicable!" << std::endl;
		exit(EXIT_FAILURE);
	}

	template <class T> T* Mirror(ListOfVariables v) { size_t n = 0; return static_cast<T*>(mpgos_perthread_host_ptr(Handle, (int)v, &n)); }

	// Per-instance variable with `Serials` entries per instance; nullptr when the shape is different.
	struct Slot { Precision* p; int* ip; int serials; };
	Slot PerInstance(ListOfVariables v)
	{
		switch (v)
		{
			case TimeDomain:         return { h_TimeDomain, nullptr, 2 };
			case ActualState:        return { h_ActualState, nullptr, SD };
			case ControlParameters:  return { h_ControlParameters, nullptr, NCP };
			case Accessories:        return { h_Accessories, nullptr, NA };
			case IntegerAccessories: return { nullptr, h_IntegerAccessories, NIA };
			case DenseTime:          return { h_DenseOutputTimeInstances, nullptr, NDO };
			default:                 return { nullptr, nullptr, -1 };
		}
	}

	template <class T>
	void WriteColumns(const std::string& FileName, int Rows, int Columns, const T* Data)   // reference: WriteToFileGeneral, Interface.cuh:1085-1110
	{
		std::ofstream f(FileName.c_str());
		f.precision(10);
		f.flags(std::ios::scientific);
		for (int r = 0; r < Rows; r++)
		{
			for (int c = 0; c < Columns; c++) { f.width(18); f << Data[r + c * Rows]; if (c < Columns - 1) f << ','; }
			f << '\n';
		}
	}

	void Construct(const int* Devices, int DeviceCount, int Tile)
	{
		mpgos_perthread_desc d;
		std::memset(&d, 0, sizeof(d));
		d.NT = NT; d.SD = SD; d.NCP = NCP; d.NSP = NSP; d.NISP = NISP; d.NE = NE; d.NA = NA; d.NIA = NIA; d.NDO = NDO;
		d.Algorithm = (int)Algorithm; d.PrecisionBytes = (int)sizeof(Precision); d.Device = Devices[0];
		Handle = nullptr;
		if (DeviceCount == 1) Check(mpgos_perthread_create(&d, &Handle), "ProblemSolver (constructor)");
		else                  Check(mpgos_perthread_create_sharded(&d, Devices, DeviceCount, Tile, &Handle), "ProblemSolver (constructor)");
		h_TimeDomain = Mirror<Precision>(TimeDomain);  h_ActualState = Mirror<Precision>(ActualState);  h_ActualTime = Mirror<Precision>(ActualTime);
		h_ControlParameters = Mirror<Precision>(ControlParameters);  h_SharedParameters = Mirror<Precision>(SharedParameters);
		h_IntegerSharedParameters = Mirror<int>(IntegerSharedParameters);  h_Accessories = Mirror<Precision>(Accessories);
		h_IntegerAccessories = Mirror<int>(IntegerAccessories);  h_DenseOutputIndex = Mirror<int>(DenseIndex);
		h_DenseOutputTimeInstances = Mirror<Precision>(DenseTime);  h_DenseOutputStates = Mirror<Precision>(DenseState);
		// the reference launches 64-thread blocks from its tutorials; our default is the runtime's
		std::cout << "Object for Parameters scan is successfully created on device " << Devices[0];
		for (int i = 1; i < DeviceCount; i++) std::cout << ", " << Devices[i];
		std::cout << " (B200-native solver, " << NT << " instances, " << (Algorithm == RK4 ? "RK4" : "RKCK45") << ", " << (sizeof(Precision) == 8 ? "double" : "float") << ")" << std::endl << std::endl;
	}

public:
	explicit ProblemSolver(int Device) { Construct(&Device, 1, 0); }

	// Addition of this implementation: ONE object over several devices of the node.  The NT instances are dealt to the devices in
	// tiles of `Tile` consecutive instances (0: 4096); the host side (SetHost / GetHost / Print, one mirror in the sweep's layout)
	// and every other member work as for one device; callbacks receive the instance's id in the whole sweep and NT.  The
	// reference needs one object per device, filled and driven by hand (T8_MultiGPUSingleNode/MultiGPUSingleNode.cu:85-135).
	explicit ProblemSolver(const std::vector<int>& Devices, int Tile = 0)
	{
		if (Devices.empty()) { std::cerr << "ERROR: In solver member function ProblemSolver (constructor)!" << std::endl << "       empty device list" << std::endl; exit(EXIT_FAILURE); }
		Construct(Devices.data(), (int)Devices.size(), Tile);
	}

	~ProblemSolver() { mpgos_perthread_destroy(Handle); std::cout << "Object for Parameters scan is deleted! Every memory have been deallocated!" << std::endl << std::endl; }

	ProblemSolver(const ProblemSolver&) = delete;
	ProblemSolver& operator=(const ProblemSolver&) = delete;

	// ---- options (Interface.cuh:613-703) ----
	template <typename T> void SolverOption(ListOfSolverOptions Option, T Value)
	{
		if (Option == RelativeTolerance || Option == AbsoluteTolerance || Option == EventTolerance || Option == EventDirection)
		{
			std::cerr << "ERROR: In solver member function SolverOption!" << std::endl << "       Option: " << SolverOptionsToString(Option) << std::endl
			          << "       This option needs 2 input arguments or not applicable!" << std::endl;
			exit(EXIT_FAILURE);
		}
		Check(mpgos_perthread_set_option(Handle, (int)Option, (double)Value), "SolverOption");
	}

	template <typename T> void SolverOption(ListOfSolverOptions Option, int Index, T Value)
	{
		switch (Option)
		{
			case RelativeTolerance: case AbsoluteTolerance: BoundCheck("SolverOption", SolverOptionsToString(Option), Index, 0, SD - 1); break;
			case EventTolerance:    case EventDirection:    BoundCheck("SolverOption", SolverOptionsToString(Option), Index, 0, NE - 1); break;
			default:
				std::cerr << "ERROR: In solver member function SolverOption!" << std::endl << "       Option: " << SolverOptionsToString(Option) << std::endl
				          << "       This option needs 1 input arguments or not applicable!" << std::endl;
				exit(EXIT_FAILURE);
		}
		Check(mpgos_perthread_set_option_indexed(Handle, (int)Option, Index, (double)Value), "SolverOption");
	}

	// ---- SetHost / GetHost (Interface.cuh:708-876, 1009-1081); same four overloads ----
	template <typename T> void SetHost(int ProblemNumber, ListOfVariables Variable, int SerialNumber, T Value)   // per instance + serial, DenseTime
	{
		BoundCheck("SetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		Slot s = PerInstance(Variable);
		if (s.serials < 0) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const int idx = ProblemNumber + SerialNumber * NT;
		if (s.ip) s.ip[idx] = (int)Value; else s.p[idx] = (Precision)Value;
	}

	template <typename T> void SetHost(int ProblemNumber, ListOfVariables Variable, T Value)                       // one value per instance
	{
		BoundCheck("SetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		if (Variable == ActualTime)      h_ActualTime[ProblemNumber] = (Precision)Value;
		else if (Variable == DenseIndex) { BoundCheck("SetHost", "DenseIndex-Value", (int)Value, 0, NDO - 1); h_DenseOutputIndex[ProblemNumber] = (int)Value; }   // value check as Interface.cuh:770
		else NotApplicable("SetHost", VariablesToString(Variable));
	}

	template <typename T> void SetHost(ListOfVariables Variable, int SerialNumber, T Value)                        // shared by all instances
	{
		if (Variable == SharedParameters)             { BoundCheck("SetHost", "SharedParameters", SerialNumber, 0, NSP - 1); h_SharedParameters[SerialNumber] = (Precision)Value; }
		else if (Variable == IntegerSharedParameters) { BoundCheck("SetHost", "IntegerSharedParameters", SerialNumber, 0, NISP - 1); h_IntegerSharedParameters[SerialNumber] = (int)Value; }
		else NotApplicable("SetHost", VariablesToString(Variable));
	}

	template <typename T> void SetHost(int ProblemNumber, ListOfVariables Variable, int SerialNumber, int TimeStepNumber, T Value)   // DenseState
	{
		BoundCheck("SetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		if (Variable != DenseState) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", "DenseState/ComponentNumber", SerialNumber, 0, SD - 1);
		BoundCheck("SetHost", "DenseState/TimeStepNumber", TimeStepNumber, 0, NDO - 1);
		h_DenseOutputStates[ProblemNumber + SerialNumber * NT + TimeStepNumber * NT * SD] = (Precision)Value;
	}

	template <typename T> T GetHost(int ProblemNumber, ListOfVariables Variable, int SerialNumber)
	{
		BoundCheck("GetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		Slot s = PerInstance(Variable);
		if (s.serials < 0) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const int idx = ProblemNumber + SerialNumber * NT;
		return s.ip ? (T)s.ip[idx] : (T)s.p[idx];
	}

	template <typename T> T GetHost(int ProblemNumber, ListOfVariables Variable)
	{
		BoundCheck("GetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		if (Variable == ActualTime) return (T)h_ActualTime[ProblemNumber];
		if (Variable == DenseIndex) return (T)h_DenseOutputIndex[ProblemNumber];
		NotApplicable("GetHost", VariablesToString(Variable));
	}

	template <typename T> T GetHost(ListOfVariables Variable, int SerialNumber)
	{
		if (Variable == SharedParameters)        { BoundCheck("GetHost", "SharedParameters", SerialNumber, 0, NSP - 1); return (T)h_SharedParameters[SerialNumber]; }
		if (Variable == IntegerSharedParameters) { BoundCheck("GetHost", "IntegerSharedParameters", SerialNumber, 0, NISP - 1); return (T)h_IntegerSharedParameters[SerialNumber]; }
		NotApplicable("GetHost", VariablesToString(Variable));
	}

	template <typename T> T GetHost(int ProblemNumber, ListOfVariables Variable, int SerialNumber, int TimeStepNumber)
	{
		BoundCheck("GetHost", "ProblemNumber", ProblemNumber, 0, NT - 1);
		if (Variable != DenseState) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", "DenseState/ComponentNumber", SerialNumber, 0, SD - 1);
		BoundCheck("GetHost", "DenseState/TimeStepNumber", TimeStepNumber, 0, NDO - 1);
		return (T)h_DenseOutputStates[ProblemNumber + SerialNumber * NT + TimeStepNumber * NT * SD];
	}

	// ---- copies (asynchronous on the object's stream, Interface.cuh:878-1005) ----
	void SynchroniseFromHostToDevice(ListOfVariables Variable) { Check(mpgos_perthread_sync_h2d(Handle, (int)Variable), "SynchroniseFromHostToDevice"); }
	void SynchroniseFromDeviceToHost(ListOfVariables Variable) { Check(mpgos_perthread_sync_d2h(Handle, (int)Variable), "SynchroniseFromDeviceToHost"); }

	// ---- text dumps, same file names and format as Interface.cuh:1113-1276 ----
	void Print(ListOfVariables Variable)
	{
		switch (Variable)
		{
			case TimeDomain:              WriteColumns("TimeDomainInSolverObject.txt", NT, 2, h_TimeDomain); break;
			case ActualState:             WriteColumns("ActualStateInSolverObject.txt", NT, SD, h_ActualState); break;
			case ActualTime:              WriteColumns("ActualTimeInSolverObject.txt", NT, 1, h_ActualTime); break;
			case ControlParameters:       WriteColumns("ControlParametersInSolverObject.txt", NT, NCP, h_ControlParameters); break;
			case SharedParameters:        WriteColumns("SharedParametersInSolverObject.txt", NSP, 1, h_SharedParameters); break;
			case IntegerSharedParameters: WriteColumns("IntegerSharedParametersInSolverObject.txt", NISP, 1, h_IntegerSharedParameters); break;
			case Accessories:             WriteColumns("AccessoriesInSolverObject.txt", NT, NA, h_Accessories); break;
			case IntegerAccessories:      WriteColumns("IntegerAccessoriesInSolverObject.txt", NT, NIA, h_IntegerAccessories); break;
			default: NotApplicable("Print", VariablesToString(Variable));
		}
	}

	void Print(ListOfVariables Variable, int ThreadID)
	{
		BoundCheck("Print", "Thread", ThreadID, 0, NT - 1);
		if (Variable != DenseOutput) NotApplicable("Print", VariablesToString(Variable));
		std::ofstream f(("DenseOutput_" + std::to_string(ThreadID) + ".txt").c_str());
		f.precision(10);
		f.flags(std::ios::scientific);
		auto row = [&](const char* title, int n, auto value) {
			f << title << ":\n";
			for (int i = 0; i < n; i++) { f.width(18); f << value(i); if (i < n - 1) f << ','; }
		};
		row("ControlParameters", NCP, [&](int i) { return h_ControlParameters[ThreadID + i * NT]; });        f << '\n';
		row("SharedParameters", NSP, [&](int i) { return h_SharedParameters[i]; });                           f << '\n';
		row("IntegerSharedParameters", NISP, [&](int i) { return h_IntegerSharedParameters[i]; });            f << '\n';
		row("Accessories", NA, [&](int i) { return h_Accessories[ThreadID + i * NT]; });                      f << '\n';
		row("IntegerAccessories", NIA, [&](int i) { return h_IntegerAccessories[ThreadID + i * NT]; });       f << "\n\n";
		f << "Time series:\n";
		if (NDO > 0)
			for (int k = 0; k < h_DenseOutputIndex[ThreadID] && k < NDO; k++)
			{
				f.width(18); f << h_DenseOutputTimeInstances[ThreadID + k * NT] << ',';
				for (int j = 0; j < SD; j++) { f.width(18); f << h_DenseOutputStates[ThreadID + j * NT + k * NT * SD]; if (j < SD - 1) f << ','; }
				f << '\n';
			}
	}

	// ---- integrate one phase (Interface.cuh:1280-1285) and synchronisation (:1289-1312) ----
	void Solve()                       { Check(mpgos_perthread_solve(Handle, &Launcher::launch), "Solve"); }
	void SynchroniseDevice()           { Check(mpgos_perthread_sync_device(Handle), "SynchroniseDevice"); }
	void InsertSynchronisationPoint()  { Check(mpgos_perthread_insert_sync_point(Handle), "InsertSynchronisationPoint"); }
	void SynchroniseSolver()           { Check(mpgos_perthread_sync_solver(Handle), "SynchroniseSolver"); }

	// ---- additions of this implementation ----
	// stepper invocations / accepted steps per instance, accumulated over all Solve() calls
	void StepCounters(std::vector<unsigned long long>& Attempted, std::vector<unsigned long long>& Accepted)
	{
		Attempted.resize(NT);  Accepted.resize(NT);
		Check(mpgos_perthread_read_step_counters(Handle, Attempted.data(), Accepted.data()), "StepCounters");
	}
	mpgos_perthread_solver* NativeHandle() { return Handle; }
};

#endif
