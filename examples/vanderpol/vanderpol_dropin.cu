// User program in the reference's style (system definition first, then the solver header, then a
// ProblemSolver object driven through SolverOption / SetHost / Synchronise* / Solve / GetHost),
// wrapped in one extern "C" function so that the tests can call it and compare with the checker.
#include "VanDerPol_SystemDefinition.cuh"
#include "SingleSystem_PerThread_Interface.cuh"

#define NT   1536
#define SD   2
#define NCP  1
#define NSP  2
#define NISP 1
#define NE   2
#define NA   2
#define NIA  2
#define NDO  64

template <Algorithms ALG>
static int run(int phases, const double* mu, double* state, double* time, double* accessories, int* integer_accessories,
               int* dense_index, double* dense_time, double* dense_state, unsigned long long* attempted, unsigned long long* accepted)
{
	ProblemSolver<NT, SD, NCP, NSP, NISP, NE, NA, NIA, NDO, ALG, double> Solver(0);

	Solver.SolverOption(ThreadsPerBlock, 64);
	Solver.SolverOption(InitialTimeStep, 1.0e-2);
	Solver.SolverOption(ActiveNumberOfThreads, NT);
	Solver.SolverOption(DenseOutputMinimumTimeStep, 0.0537);
	Solver.SolverOption(DenseOutputSaveFrequency, 2);
	for (int i = 0; i < SD; i++) { Solver.SolverOption(RelativeTolerance, i, 1.0e-9); Solver.SolverOption(AbsoluteTolerance, i, 1.0e-9); }
	Solver.SolverOption(EventTolerance, 0, 1.0e-7);
	Solver.SolverOption(EventTolerance, 1, 1.0e-7);
	Solver.SolverOption(EventDirection, 0, -1);
	Solver.SolverOption(EventDirection, 1, 0);

	for (int p = 0; p < NT; p++)
	{
		Solver.SetHost(p, TimeDomain, 0, 0.0);
		Solver.SetHost(p, TimeDomain, 1, 1.0e3);
		Solver.SetHost(p, ActualState, 0, 2.0);
		Solver.SetHost(p, ActualState, 1, 0.0);
		Solver.SetHost(p, ActualTime, 0.0);
		Solver.SetHost(p, ControlParameters, 0, mu[p]);
		Solver.SetHost(p, Accessories, 0, 0.0);
		Solver.SetHost(p, Accessories, 1, 0.0);
		Solver.SetHost(p, IntegerAccessories, 0, 0);
		Solver.SetHost(p, IntegerAccessories, 1, 0);
		Solver.SetHost(p, DenseIndex, 0);
	}
	Solver.SetHost(SharedParameters, 0, 0.5);
	Solver.SetHost(SharedParameters, 1, 1.1);
	Solver.SetHost(IntegerSharedParameters, 0, 3);

	Solver.SynchroniseFromHostToDevice(All);
	for (int i = 0; i < phases; i++) Solver.Solve();
	Solver.SynchroniseFromDeviceToHost(All);
	Solver.InsertSynchronisationPoint();
	Solver.SynchroniseSolver();

	for (int p = 0; p < NT; p++)
	{
		for (int i = 0; i < SD; i++)  state[p + i * NT] = Solver.template GetHost<double>(p, ActualState, i);
		time[p] = Solver.template GetHost<double>(p, ActualTime);
		for (int i = 0; i < NA; i++)  accessories[p + i * NT] = Solver.template GetHost<double>(p, Accessories, i);
		for (int i = 0; i < NIA; i++) integer_accessories[p + i * NT] = Solver.template GetHost<int>(p, IntegerAccessories, i);
		dense_index[p] = Solver.template GetHost<int>(p, DenseIndex);
		for (int k = 0; k < NDO; k++)
		{
			dense_time[p + k * NT] = Solver.template GetHost<double>(p, DenseTime, k);
			for (int i = 0; i < SD; i++) dense_state[p + i * NT + k * NT * SD] = Solver.template GetHost<double>(p, DenseState, i, k);
		}
	}
	std::vector<unsigned long long> a, b;
	Solver.StepCounters(a, b);
	for (int p = 0; p < NT; p++) { attempted[p] = a[p]; accepted[p] = b[p]; }
	Solver.Print(ActualState);
	Solver.Print(DenseOutput, 7);
	return 0;
}

extern "C" int vanderpol_instances() { return NT; }

extern "C" int vanderpol_run(int algorithm, int phases, const double* mu, double* state, double* time, double* accessories, int* integer_accessories,
                             int* dense_index, double* dense_time, double* dense_state, unsigned long long* attempted, unsigned long long* accepted)
{
	if (algorithm == 0) return run<RK4>(phases, mu, state, time, accessories, integer_accessories, dense_index, dense_time, dense_state, attempted, accepted);
	return run<RKCK45>(phases, mu, state, time, accessories, integer_accessories, dense_index, dense_time, dense_state, attempted, accepted);
}
