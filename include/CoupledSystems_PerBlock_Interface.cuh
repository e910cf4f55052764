// CoupledSystems_PerBlock_Interface.cuh -- source-level drop-in for the MPGOS header of the same name.
//
// A user program written against the reference's coupled module keeps working unchanged:
//
//     #include "MySystem_SystemDefinition.cuh"            // the six CoupledSystems_PerBlock_* device callbacks
//     #include "CoupledSystems_PerBlock_Interface.cuh"    // this file (include order as in the reference)
//     ProblemSolver<NS,UPS,UD,TPB,SPB,NC,CBW,CCI,NUP,NSP,NGP,NiGP,NUA,NiUA,NSA,NiSA,NE,NDO,RKCK45,double> Solver(Device);
//     Solver.SolverOption(...); Solver.SetHost(...); Solver.SynchroniseFromHostToDevice(All); Solver.Solve(); ...
//
// Same unscoped enums, template signature, public methods, error behaviour and file formats as
// SourceCodes/CoupledSystems_PerBlock_Interface.cuh:33-248 of the reference (abbreviated CP/Interface.cuh below).
// Behind the surface the host object is a thin shell over the C-ABI of libmpgos_b200.so (include/mpgos_b200_coupled.h)
// and Solve() hands that runtime a launcher for OUR kernel (include/mpgos_b200/perblock_kernel.cuh), instantiated here
// in the user's translation unit so that the callbacks inline into it.  Link with -lmpgos_b200.
//
// Deliberate differences:
//   * TPB and SPB are accepted and only shape the text dumps: the kernel chooses its own block shape (whole warps per
//     system, as many systems per block as fit 512 threads; 8 in lockstep on the FP64 tensor cores for dense matrices);
//   * SharedGlobalVariables / SharedCouplingMatrices are accepted and ignored: both always live in shared memory;
//   * device-side warnings are printed only after SolverOption(DeviceMessages, 1);
//   * extra options DeviceMessages / TensorCoreCoupling / DispatchOrder and the StepCounters() accessor;
//   * Precision = double or float; dense shared matrices are multiplied on the tensor cores in both (FP64 DMMA; FP32 as 3xTF32).
#ifndef COUPLEDSYSTEMS_PERBLOCK_INTERFACE_H
#define COUPLEDSYSTEMS_PERBLOCK_INTERFACE_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>
#include <cuda_runtime.h>

#include "mpgos_b200_coupled.h"
#include "mpgos_b200/perblock_kernel.cuh"

// ---- the reference's unscoped enums (CP/Interface.cuh:33-68); values equal the C-ABI's ----------------------------
enum Algorithms { RK4, RKCK45 };

enum ListOfVariables { All, TimeDomain, ActualTime, ActualState, UnitParameters, SystemParameters, GlobalParameters, IntegerGlobalParameters,
                       UnitAccessories, IntegerUnitAccessories, SystemAccessories, IntegerSystemAccessories, CouplingMatrix, CouplingStrength,
                       CouplingIndex, DenseOutput, DenseIndex, DenseTime, DenseState };

enum ListOfSolverOptions { InitialTimeStep, ActiveSystems, MaximumTimeStep, MinimumTimeStep, TimeStepGrowLimit, TimeStepShrinkLimit,
                           RelativeTolerance, AbsoluteTolerance, EventTolerance, EventDirection, DenseOutputSaveFrequency,
                           DenseOutputMinimumTimeStep, SharedGlobalVariables, SharedCouplingMatrices,
                           DeviceMessages, TensorCoreCoupling, DispatchOrder };

inline std::string VariablesToString(ListOfVariables v)
{
	static const char* const names[] = { "All", "TimeDomain", "ActualTime", "ActualState", "UnitParameters", "SystemParameters", "GlobalParameters",
	                                     "IntegerGlobalParameters", "UnitAccessories", "IntegerUnitAccessories", "SystemAccessories",
	                                     "IntegerSystemAccessories", "CouplingMatrix", "CouplingStrength", "CouplingIndex", "DenseOutput",
	                                     "DenseIndex", "DenseTime", "DenseState" };
	return ((int)v >= 0 && (int)v <= (int)DenseState) ? names[(int)v] : "Non-existent variable!";
}

inline std::string SolverOptionsToString(ListOfSolverOptions o)
{
	static const char* const names[] = { "InitialTimeStep", "ActiveSystems", "MaximumTimeStep", "MinimumTimeStep", "TimeStepGrowLimit",
	                                     "TimeStepShrinkLimit", "RelativeTolerance", "AbsoluteTolerance", "EventTolerance", "EventDirection",
	                                     "DenseOutputSaveFrequency", "DenseOutputMinimumTimeStep", "SharedGlobalVariables",
	                                     "SharedCouplingMatrices", "DeviceMessages", "TensorCoreCoupling", "DispatchOrder" };
	return ((int)o >= 0 && (int)o <= (int)DispatchOrder) ? names[(int)o] : "Non-existent solver option!";
}

// ---- device discovery helpers (CP/Interface.cuh:258-344) ---------------------------------------------------------------
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

// ---- binding of the user's callbacks ----------------------------------------------------------------------------------
// Forwards to the functions of the fixed names the user defined BEFORE including this header (found by ordinary lookup
// here, as the reference finds them from its kernels, CoupledSystmes_PerBlock_Solver.cuh:287-698).
namespace mpgos_b200 {
template <class P>
struct UserCoupledCallbacks
{
	static __device__ __forceinline__ void OdeFunction(int sid, int uid, P* F, P* X, P T, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC, P* CPT, P* CPF)
	{ CoupledSystems_PerBlock_OdeFunction(sid, uid, F, X, T, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC, CPT, CPF); }
	static __device__ __forceinline__ void EventFunction(int sid, int uid, P* EF, P T, P dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_EventFunction(sid, uid, EF, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static __device__ __forceinline__ void ActionAfterEventDetection(int sid, int uid, int IDX, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_ActionAfterEventDetection(sid, uid, IDX, UDT, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static __device__ __forceinline__ void ActionAfterSuccessfulTimeStep(int sid, int uid, int& UDT, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_ActionAfterSuccessfulTimeStep(sid, uid, UDT, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static __device__ __forceinline__ void Initialization(int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_Initialization(sid, uid, DOIDX, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
	static __device__ __forceinline__ void Finalization(int sid, int uid, int& DOIDX, P& T, P& dT, P* TD, P* X, P* uPAR, P* sPAR, P* gPAR, int* igPAR, P* uACC, int* iuACC, P* sACC, int* isACC)
	{ CoupledSystems_PerBlock_Finalization(sid, uid, DOIDX, T, dT, TD, X, uPAR, sPAR, gPAR, igPAR, uACC, iuACC, sACC, isACC); }
};
}

// ---- the host object ------------------------------------------------------------------------------------------------------
template <int NS, int UPS, int UD, int TPB, int SPB, int NC, int CBW, int CCI, int NUP, int NSP, int NGP, int NiGP, int NUA, int NiUA, int NSA, int NiSA,
          int NE, int NDO, Algorithms Algorithm, class Precision>
class ProblemSolver
{
	static_assert(std::is_same<Precision, double>::value || std::is_same<Precision, float>::value, "Precision is double or float");

	mpgos_coupled_solver* Handle;
	// pinned host mirrors owned by the runtime (reference: h_* members, CP/Interface.cuh:177-194); unit scope is stored
	// WITHOUT the reference's per-block padding: element (system s, unit u, serial i) at u + s*UPS + i*NS*UPS
	Precision *h_TimeDomain, *h_ActualState, *h_ActualTime, *h_UnitParameters, *h_SystemParameters, *h_GlobalParameters, *h_UnitAccessories,
	          *h_SystemAccessories, *h_CouplingStrength, *h_DenseOutputTimeInstances, *h_DenseOutputStates;
	int       *h_IntegerGlobalParameters, *h_IntegerUnitAccessories, *h_IntegerSystemAccessories, *h_CouplingIndex, *h_DenseOutputIndex;

	using Launcher = mpgos_b200::PerBlockLauncher<mpgos_b200::UserCoupledCallbacks<Precision>, Precision, UPS, UD, NC, CBW, CCI, NUP, NSP, NGP, NiGP,
	                                              NUA, NiUA, NSA, NiSA, NE, (NDO > 0), (int)Algorithm>;

	[[noreturn]] static void Fatal(const char* where)
	{
		std::cerr << "ERROR: In solver member function " << where << "!" << std::endl << "       " << mpgos_last_error() << std::endl;
		exit(EXIT_FAILURE);
	}
	static void Check(int rc, const char* where) { if (rc != MPGOS_OK) Fatal(where); }

	// reference: BoundCheck, CP/Interface.cuh:2067-2083 (message on std::cerr, then exit)
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

	template <class T> T* Mirror(ListOfVariables v) { size_t n = 0; return static_cast<T*>(mpgos_coupled_host_ptr(Handle, (int)v, &n)); }

	struct Slot { Precision* p; int* ip; int serials; };
	Slot UnitScope(ListOfVariables v)
	{
		switch (v)
		{
			case ActualState:            return { h_ActualState, nullptr, UD };
			case UnitParameters:         return { h_UnitParameters, nullptr, NUP };
			case UnitAccessories:        return { h_UnitAccessories, nullptr, NUA };
			case IntegerUnitAccessories: return { nullptr, h_IntegerUnitAccessories, NiUA };
			default:                     return { nullptr, nullptr, -1 };
		}
	}
	Slot SystemScope(ListOfVariables v)
	{
		switch (v)
		{
			case TimeDomain:               return { h_TimeDomain, nullptr, 2 };
			case SystemParameters:         return { h_SystemParameters, nullptr, NSP };
			case SystemAccessories:        return { h_SystemAccessories, nullptr, NSA };
			case IntegerSystemAccessories: return { nullptr, h_IntegerSystemAccessories, NiSA };
			case DenseTime:                return { h_DenseOutputTimeInstances, nullptr, NDO };
			case CouplingStrength:         return { h_CouplingStrength, nullptr, NC };
			default:                       return { nullptr, nullptr, -1 };
		}
	}
	Slot GlobalScope(ListOfVariables v)
	{
		switch (v)
		{
			case GlobalParameters:        return { h_GlobalParameters, nullptr, NGP };
			case IntegerGlobalParameters: return { nullptr, h_IntegerGlobalParameters, NiGP };
			case CouplingIndex:           return { nullptr, h_CouplingIndex, NC };
			default:                      return { nullptr, nullptr, -1 };
		}
	}
	static size_t UnitIndex(int s, int u, int serial) { return (size_t)u + (size_t)s * UPS + (size_t)serial * NS * UPS; }

	// the reference's thread bookkeeping (CP/Interface.cuh:386-391), kept only to reproduce the row order and the
	// PADDING rows of its unit-scope text dumps
	static constexpr int LogicalThreadsPerBlock = SPB * UPS;
	static constexpr int NumberOfBlockLaunches  = LogicalThreadsPerBlock / TPB + (LogicalThreadsPerBlock % TPB == 0 ? 0 : 1);
	static constexpr int ThreadPaddingPerBlock  = NumberOfBlockLaunches * TPB - LogicalThreadsPerBlock;
	static constexpr int GridSize               = NS / SPB + (NS % SPB == 0 ? 0 : 1);
	static constexpr int TotalLogicalThreads    = (LogicalThreadsPerBlock + ThreadPaddingPerBlock) * GridSize;

	static void OpenFile(std::ofstream& f, const std::string& name) { f.open(name.c_str()); f.precision(10); f.flags(std::ios::scientific); }

	template <class T> void WriteUnitScope(const std::string& FileName, int Columns, const T* Data)            // CP/Interface.cuh:1533-1571
	{
		std::ofstream f;  OpenFile(f, FileName);
		const int per_block = LogicalThreadsPerBlock + ThreadPaddingPerBlock;
		const int limit = (NS / SPB) * per_block + (NS % SPB) * UPS;
		for (int i = 0; i < TotalLogicalThreads; i++)
		{
			const int local = i % per_block, block = i / per_block;
			const bool real = local < LogicalThreadsPerBlock && i < limit;
			for (int j = 0; j < Columns; j++)
			{
				f.width(18);
				if (real) f << Data[UnitIndex(block * SPB + local / UPS, local % UPS, j)] << ',';
				else      f << "PADDING" << ',';
			}
			f << '\n';
		}
	}
	template <class T> void WriteColumns(const std::string& FileName, int Rows, int Columns, const T* Data)     // CP/Interface.cuh:1575-1597
	{
		std::ofstream f;  OpenFile(f, FileName);
		for (int i = 0; i < Rows; i++)
		{
			for (int j = 0; j < Columns; j++) { f.width(18); f << Data[i + j * Rows] << ','; }
			f << '\n';
		}
	}

	void Construct(const int* Devices, int DeviceCount, int Tile)
	{
		const int Device = Devices[0];
		mpgos_coupled_desc d;
		std::memset(&d, 0, sizeof(d));
		d.NS = NS; d.UPS = UPS; d.UD = UD; d.TPB = TPB; d.SPB = SPB; d.NC = NC; d.CBW = CBW; d.CCI = CCI; d.NUP = NUP; d.NSP = NSP; d.NGP = NGP; d.NiGP = NiGP;
		d.NUA = NUA; d.NiUA = NiUA; d.NSA = NSA; d.NiSA = NiSA; d.NE = NE; d.NDO = NDO;
		d.Algorithm = (int)Algorithm; d.PrecisionBytes = (int)sizeof(Precision); d.Device = Device;
		Handle = nullptr;
		if (DeviceCount == 1) Check(mpgos_coupled_create(&d, &Handle), "ProblemSolver (constructor)");
		else                  Check(mpgos_coupled_create_sharded(&d, Devices, DeviceCount, Tile, &Handle), "ProblemSolver (constructor)");
		h_TimeDomain = Mirror<Precision>(TimeDomain);  h_ActualState = Mirror<Precision>(ActualState);  h_ActualTime = Mirror<Precision>(ActualTime);
		h_UnitParameters = Mirror<Precision>(UnitParameters);  h_SystemParameters = Mirror<Precision>(SystemParameters);
		h_GlobalParameters = Mirror<Precision>(GlobalParameters);  h_IntegerGlobalParameters = Mirror<int>(IntegerGlobalParameters);
		h_UnitAccessories = Mirror<Precision>(UnitAccessories);  h_IntegerUnitAccessories = Mirror<int>(IntegerUnitAccessories);
		h_SystemAccessories = Mirror<Precision>(SystemAccessories);  h_IntegerSystemAccessories = Mirror<int>(IntegerSystemAccessories);
		h_CouplingStrength = Mirror<Precision>(CouplingStrength);  h_CouplingIndex = Mirror<int>(CouplingIndex);
		h_DenseOutputIndex = Mirror<int>(DenseIndex);  h_DenseOutputTimeInstances = Mirror<Precision>(DenseTime);  h_DenseOutputStates = Mirror<Precision>(DenseState);
		std::cout << "Object for Parameters scan is successfully created on device " << Device;
		for (int i = 1; i < DeviceCount; i++) std::cout << ", " << Devices[i];
		std::cout << " (B200-native coupled solver, " << NS << " systems x "
		          << UPS << " units, " << (Algorithm == RK4 ? "RK4" : "RKCK45") << (sizeof(Precision) == 8 ? ", double)" : ", float)") << std::endl << std::endl;
	}

public:
	explicit ProblemSolver(int Device) { Construct(&Device, 1, 0); }

	// Addition of this implementation: ONE object over several devices of the node; the NS systems are dealt to them in tiles of
	// `Tile` systems (0: 64), coupling matrices and global parameters are replicated, the host side is unchanged (one mirror over
	// all systems), callbacks receive the system's id in the whole sweep.  A single system is never split.
	explicit ProblemSolver(const std::vector<int>& Devices, int Tile = 0)
	{
		if (Devices.empty()) { std::cerr << "ERROR: In solver member function ProblemSolver (constructor)!" << std::endl << "       empty device list" << std::endl; exit(EXIT_FAILURE); }
		Construct(Devices.data(), (int)Devices.size(), Tile);
	}

	~ProblemSolver() { mpgos_coupled_destroy(Handle); std::cout << "Object for Parameters scan is deleted! Every memory have been deallocated!" << std::endl << std::endl; }

	ProblemSolver(const ProblemSolver&) = delete;
	ProblemSolver& operator=(const ProblemSolver&) = delete;

	// ---- options (CP/Interface.cuh:704-865) ----
	template <typename T> void SolverOption(ListOfSolverOptions Option, T Value)
	{
		if (Option == RelativeTolerance || Option == AbsoluteTolerance || Option == EventTolerance || Option == EventDirection)
		{
			std::cerr << "ERROR: In solver member function SolverOption!" << std::endl << "       Option: " << SolverOptionsToString(Option) << std::endl
			          << "       This option needs 2 input arguments or not applicable!" << std::endl;
			exit(EXIT_FAILURE);
		}
		Check(mpgos_coupled_set_option(Handle, (int)Option, (double)Value), "SolverOption");
	}

	template <typename T> void SolverOption(ListOfSolverOptions Option, int Index, T Value)
	{
		switch (Option)
		{
			case RelativeTolerance: case AbsoluteTolerance: BoundCheck("SolverOption", SolverOptionsToString(Option), Index, 0, UD - 1); break;
			case EventTolerance:    case EventDirection:    BoundCheck("SolverOption", SolverOptionsToString(Option), Index, 0, NE - 1); break;
			default:
				std::cerr << "ERROR: In solver member function SolverOption!" << std::endl << "       Option: " << SolverOptionsToString(Option) << std::endl
				          << "       This option needs 1 input arguments or not applicable!" << std::endl;
				exit(EXIT_FAILURE);
		}
		Check(mpgos_coupled_set_option_indexed(Handle, (int)Option, Index, (double)Value), "SolverOption");
	}

	// ---- SetHost, the six overloads of CP/Interface.cuh:868-1213 ----
	template <typename T> void SetHost(int SystemNumber, int UnitNumber, ListOfVariables Variable, int SerialNumber, T Value)            // unit scope
	{
		BoundCheck("SetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		BoundCheck("SetHost", "UnitNumber", UnitNumber, 0, UPS - 1);
		Slot s = UnitScope(Variable);
		if (s.serials < 0) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const size_t idx = UnitIndex(SystemNumber, UnitNumber, SerialNumber);
		if (s.ip) s.ip[idx] = (int)Value; else s.p[idx] = (Precision)Value;
	}

	template <typename T> void SetHost(int SystemNumber, ListOfVariables Variable, int SerialNumber, T Value)                            // system scope, dense time
	{
		BoundCheck("SetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		Slot s = SystemScope(Variable);
		if (s.serials < 0) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const size_t idx = (size_t)SystemNumber + (size_t)SerialNumber * NS;
		if (s.ip) s.ip[idx] = (int)Value; else s.p[idx] = (Precision)Value;
	}

	template <typename T> void SetHost(ListOfVariables Variable, int SerialNumber, T Value)                                              // global scope
	{
		Slot s = GlobalScope(Variable);
		if (s.serials < 0) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		if (s.ip) s.ip[SerialNumber] = (int)Value; else s.p[SerialNumber] = (Precision)Value;
	}

	template <typename T> void SetHost(int CouplingNumber, ListOfVariables Variable, int Row, int Col, T Value)                          // coupling matrix
	{
		BoundCheck("SetHost", "CouplingNumber", CouplingNumber, 0, NC - 1);
		BoundCheck("SetHost", "Coupling Matrix Row", Row, 0, UPS - 1);
		BoundCheck("SetHost", "Coupling Matrix Col", Col, 0, UPS - 1);
		if (Variable != CouplingMatrix) NotApplicable("SetHost", VariablesToString(Variable));
		if (CBW > 0)   // entries outside the band are an index error, CP/Interface.cuh:1040-1056
		{
			int ModCol = Col - Row + CBW;
			if (ModCol >= UPS) ModCol -= UPS;
			if (ModCol < 0) ModCol += UPS;
			BoundCheck("SetHost", "Coupling Matrix Shifted Diagonal", ModCol, 0, 2 * CBW);
		}
		Check(mpgos_coupled_set_matrix_entry(Handle, CouplingNumber, Row, Col, (double)Value), "SetHost");
	}

	template <typename T> void SetHost(int SystemNumber, ListOfVariables Variable, T Value)                                              // dense index, actual time
	{
		BoundCheck("SetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		if (Variable == DenseIndex)      { BoundCheck("SetHost", "DenseIndex-Value", (int)Value, 0, NDO - 1); h_DenseOutputIndex[SystemNumber] = (int)Value; }
		else if (Variable == ActualTime) h_ActualTime[SystemNumber] = (Precision)Value;
		else NotApplicable("SetHost", VariablesToString(Variable));
	}

	template <typename T> void SetHost(int SystemNumber, int UnitNumber, ListOfVariables Variable, int ComponentNumber, int SerialNumber, T Value)   // dense state
	{
		BoundCheck("SetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		BoundCheck("SetHost", "UnitNumber", UnitNumber, 0, UPS - 1);
		BoundCheck("SetHost", "ComponentNumber in dense state", ComponentNumber, 0, UD - 1);
		if (Variable != DenseState) NotApplicable("SetHost", VariablesToString(Variable));
		BoundCheck("SetHost", "DenseState", SerialNumber, 0, NDO - 1);
		h_DenseOutputStates[UnitIndex(SystemNumber, UnitNumber, ComponentNumber) + (size_t)SerialNumber * NS * UPS * UD] = (Precision)Value;
	}

	// ---- GetHost, the six overloads of CP/Interface.cuh:1403-1529 ----
	template <typename T> T GetHost(int SystemNumber, int UnitNumber, ListOfVariables Variable, int SerialNumber)
	{
		BoundCheck("GetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		BoundCheck("GetHost", "UnitNumber", UnitNumber, 0, UPS - 1);
		Slot s = UnitScope(Variable);
		if (s.serials < 0) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const size_t idx = UnitIndex(SystemNumber, UnitNumber, SerialNumber);
		return s.ip ? (T)s.ip[idx] : (T)s.p[idx];
	}

	template <typename T> T GetHost(int SystemNumber, ListOfVariables Variable, int SerialNumber)
	{
		BoundCheck("GetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		Slot s = SystemScope(Variable);
		if (s.serials < 0) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		const size_t idx = (size_t)SystemNumber + (size_t)SerialNumber * NS;
		return s.ip ? (T)s.ip[idx] : (T)s.p[idx];
	}

	template <typename T> T GetHost(ListOfVariables Variable, int SerialNumber)
	{
		Slot s = GlobalScope(Variable);
		if (s.serials < 0) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", VariablesToString(Variable), SerialNumber, 0, s.serials - 1);
		return s.ip ? (T)s.ip[SerialNumber] : (T)s.p[SerialNumber];
	}

	template <typename T> T GetHost(int CouplingNumber, ListOfVariables Variable, int Row, int Col)
	{
		BoundCheck("GetHost", "CouplingNumber", CouplingNumber, 0, NC - 1);
		BoundCheck("GetHost", "Coupling Matrix Row", Row, 0, UPS - 1);
		BoundCheck("GetHost", "Coupling Matrix Col", Col, 0, UPS - 1);
		if (Variable != CouplingMatrix) NotApplicable("GetHost", VariablesToString(Variable));
		double v = 0;
		Check(mpgos_coupled_get_matrix_entry(Handle, CouplingNumber, Row, Col, &v), "GetHost");
		return (T)v;
	}

	template <typename T> T GetHost(int SystemNumber, ListOfVariables Variable)
	{
		BoundCheck("GetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		if (Variable == DenseIndex) return (T)h_DenseOutputIndex[SystemNumber];
		if (Variable == ActualTime) return (T)h_ActualTime[SystemNumber];
		NotApplicable("GetHost", VariablesToString(Variable));
	}

	template <typename T> T GetHost(int SystemNumber, int UnitNumber, ListOfVariables Variable, int ComponentNumber, int SerialNumber)
	{
		BoundCheck("GetHost", "SystemNumber", SystemNumber, 0, NS - 1);
		BoundCheck("GetHost", "UnitNumber", UnitNumber, 0, UPS - 1);
		BoundCheck("GetHost", "ComponentNumber in dense state", ComponentNumber, 0, UD - 1);
		if (Variable != DenseState) NotApplicable("GetHost", VariablesToString(Variable));
		BoundCheck("GetHost", "DenseState", SerialNumber, 0, NDO - 1);
		return (T)h_DenseOutputStates[UnitIndex(SystemNumber, UnitNumber, ComponentNumber) + (size_t)SerialNumber * NS * UPS * UD];
	}

	// ---- copies (asynchronous on the object's stream, CP/Interface.cuh:1215-1400) ----
	void SynchroniseFromHostToDevice(ListOfVariables Variable) { Check(mpgos_coupled_sync_h2d(Handle, (int)Variable), "SynchroniseFromHostToDevice"); }
	void SynchroniseFromDeviceToHost(ListOfVariables Variable) { Check(mpgos_coupled_sync_d2h(Handle, (int)Variable), "SynchroniseFromDeviceToHost"); }

	// ---- text dumps: same file names, row order and format as CP/Interface.cuh:1533-1946 ----
	void Print(ListOfVariables Variable)
	{
		switch (Variable)
		{
			case TimeDomain:               WriteColumns("TimeDomainInSolverObject.txt", NS, 2, h_TimeDomain); break;
			case ActualTime:               WriteColumns("ActualTimeInSolverObject.txt", NS, 1, h_ActualTime); break;
			case ActualState:              WriteUnitScope("ActualStateInSolverObject.txt", UD, h_ActualState); break;
			case UnitParameters:           WriteUnitScope("UnitParametersInSolverObject.txt", NUP, h_UnitParameters); break;
			case SystemParameters:         WriteColumns("SystemParametersInSolverObject.txt", NS, NSP, h_SystemParameters); break;
			case GlobalParameters:         WriteColumns("GlobalParametersInSolverObject.txt", NGP, 1, h_GlobalParameters); break;
			case IntegerGlobalParameters:  WriteColumns("IntegerGlobalParametersInSolverObject.txt", NiGP, 1, h_IntegerGlobalParameters); break;
			case UnitAccessories:          WriteUnitScope("UnitAccessoriesInSolverObject.txt", NUA, h_UnitAccessories); break;
			case IntegerUnitAccessories:   WriteUnitScope("IntegerUnitAccessoriesInSolverObject.txt", NiUA, h_IntegerUnitAccessories); break;
			case SystemAccessories:        WriteColumns("SystemAccessoriesInSolverObject.txt", NS, NSA, h_SystemAccessories); break;
			case IntegerSystemAccessories: WriteColumns("IntegerSystemAccessoriesInSolverObject.txt", NS, NiSA, h_IntegerSystemAccessories); break;
			case CouplingStrength:         WriteColumns("CouplingStrengthInSolverObject.txt", NS, NC, h_CouplingStrength); break;
			case CouplingIndex:            WriteColumns("CouplingIndexInSolverObject.txt", NC, 1, h_CouplingIndex); break;
			default: NotApplicable("Print", VariablesToString(Variable));
		}
	}

	void Print(ListOfVariables Variable, int SerialNumber)
	{
		if (Variable == CouplingMatrix)     // the matrix as the user entered it (row, col), CP/Interface.cuh:1927-1933
		{
			BoundCheck("Print", "CouplingMatrix", SerialNumber, 0, NC - 1);
			std::vector<Precision> M((size_t)UPS * UPS);
			for (int r = 0; r < UPS; r++)
				for (int c = 0; c < UPS; c++) { double v = 0; Check(mpgos_coupled_get_matrix_entry(Handle, SerialNumber, r, c, &v), "Print"); M[r + (size_t)c * UPS] = (Precision)v; }
			WriteColumns("CouplingMatrixInSolverObject_" + std::to_string(SerialNumber) + ".txt", UPS, UPS, M.data());
			return;
		}
		if (Variable != DenseOutput) NotApplicable("Print", VariablesToString(Variable));
		BoundCheck("Print", "DenseOutput", SerialNumber, 0, NS - 1);
		const int sid = SerialNumber;
		std::ofstream f;  OpenFile(f, "DenseOutput_" + std::to_string(sid) + ".txt");
		auto unit_header = [&]() {
			f.width(18); f << "Serial No." << ',';
			for (int i = 0; i < UPS; i++) { f.width(18); f << "U" + std::to_string(i); if (i < UPS - 1) f << ','; }
			f << '\n';
		};
		auto unit_table = [&](const char* title, int n, auto value) {
			f << title << ":\n";
			unit_header();
			for (int i = 0; i < n; i++)
			{
				f.width(18); f << i << ',';
				for (int j = 0; j < UPS; j++) { f.width(18); f << value(j, i); if (j < UPS - 1) f << ','; }
				f << '\n';
			}
			f << '\n';
		};
		auto row = [&](const char* title, int n, auto value) {
			f << title << ":\n";
			for (int i = 0; i < n; i++) { f.width(18); f << value(i); if (i < n - 1) f << ','; }
			f << '\n' << '\n';
		};
		unit_table("UnitParameters", NUP, [&](int u, int i) { return h_UnitParameters[UnitIndex(sid, u, i)]; });
		unit_table("UnitAccessories", NUA, [&](int u, int i) { return h_UnitAccessories[UnitIndex(sid, u, i)]; });
		unit_table("IntegerUnitAccessories", NiUA, [&](int u, int i) { return h_IntegerUnitAccessories[UnitIndex(sid, u, i)]; });
		row("SystemParameters", NSP, [&](int i) { return h_SystemParameters[sid + (size_t)i * NS]; });
		row("SystemAccessories", NSA, [&](int i) { return h_SystemAccessories[sid + (size_t)i * NS]; });
		row("IntegerSystemAccessories", NiSA, [&](int i) { return h_IntegerSystemAccessories[sid + (size_t)i * NS]; });
		row("GlobalParameters", NGP, [&](int i) { return h_GlobalParameters[i]; });
		row("IntegerGlobalParameters", NiGP, [&](int i) { return h_IntegerGlobalParameters[i]; });
		f << "Time series:\n";
		f.width(18); f << "t" << ',';
		for (int j = 0; j < UPS; j++)
			for (int k = 0; k < UD; k++) { f.width(18); f << "U" + std::to_string(j) + "/C" + std::to_string(k); if (j < UPS - 1 || k < UD - 1) f << ','; }
		f << '\n';
		const int rows = NDO > 0 ? (h_DenseOutputIndex[sid] < NDO ? h_DenseOutputIndex[sid] : NDO) : 0;
		for (int i = 0; i < rows; i++)
		{
			f.width(18); f << h_DenseOutputTimeInstances[sid + (size_t)i * NS] << ',';
			for (int j = 0; j < UPS; j++)
				for (int k = 0; k < UD; k++)
				{
					f.width(18); f << h_DenseOutputStates[UnitIndex(sid, j, k) + (size_t)i * NS * UPS * UD];
					if (j < UPS - 1 || k < UD - 1) f << ',';
				}
			f << '\n';
		}
	}

	// ---- integrate one phase (CP/Interface.cuh:1949-1965) and synchronisation (:1969-1992) ----
	void Solve()                       { Check(mpgos_coupled_solve(Handle, &Launcher::launch), "Solve"); }
	void SynchroniseDevice()           { Check(mpgos_coupled_sync_device(Handle), "SynchroniseDevice"); }
	void InsertSynchronisationPoint()  { Check(mpgos_coupled_insert_sync_point(Handle), "InsertSynchronisationPoint"); }
	void SynchroniseSolver()           { Check(mpgos_coupled_sync_solver(Handle), "SynchroniseSolver"); }

	// ---- additions of this implementation ----
	void StepCounters(std::vector<unsigned long long>& Attempted, std::vector<unsigned long long>& Accepted)   // per system, accumulated over Solve() calls
	{
		Attempted.resize(NS);  Accepted.resize(NS);
		Check(mpgos_coupled_read_step_counters(Handle, Attempted.data(), Accepted.data()), "StepCounters");
	}
	mpgos_coupled_solver* NativeHandle() { return Handle; }
};

#endif
