enum Algorithms{ RK4, RKCK45 };
struct Struct_ThreadConfiguration
{
	int GridSize;
	int BlockSize;
	int NumberOfActiveThreads;
};

template <class Precision>
struct Struct_GlobalVariables
{
	Precision* d_TimeDomain;
	Precision* d_ActualState;
	Precision* d_ActualTime;
	Precision* d_ControlParameters;
	Precision* d_SharedParameters;
	int*       d_IntegerSharedParameters;
	Precision* d_Accessories;
	int*       d_IntegerAccessories;
	Precision* d_RelativeTolerance;
	Precision* d_AbsoluteTolerance;
	Precision* d_EventTolerance;
	int*       d_EventDirection;
	int*       d_DenseOutputIndex;
	Precision* d_DenseOutputTimeInstances;
	Precision* d_DenseOutputStates;
};

struct Struct_SharedMemoryUsage
{
	int PreferSharedMemory;  // Default: ON
	int IsAdaptive;
};

template <class Precision>
struct Struct_SolverOptions
{
	Precision InitialTimeStep;
	Precision MaximumTimeStep;
	Precision MinimumTimeStep;
	Precision TimeStepGrowLimit;
	Precision TimeStepShrinkLimit;
	int       DenseOutputSaveFrequency;
	Precision DenseOutputMinimumTimeStep;
};
