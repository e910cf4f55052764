// User program for the COUPLED module in the reference's style (system definition first, then the solver header, then a
// ProblemSolver object driven through SolverOption / SetHost / Synchronise* / Solve / GetHost / Print), wrapped in one
// extern "C" function so that the tests can call it and compare with the checker.
#include "SwingRing_SystemDefinition.cuh"
#include "CoupledSystems_PerBlock_Interface.cuh"

const int NS   = 37;    // systems
const int UPS  = 40;    // units per system (not a multiple of the warp size)
const int UD   = 2;
const int TPB  = 64;    // accepted for source compatibility
const int SPB  = 1;
const int NC   = 2;
const int CBW  = 2;     // banded matrix with wrap-around, one row per unit
const int CCI  = 0;
const int NUP  = 2;
const int NSP  = 1;
const int NGP  = 1;
const int NiGP = 1;
const int NUA  = 1;
const int NiUA = 1;
const int NSA  = 1;
const int NiSA = 1;
const int NE   = 1;
const int NDO  = 32;

extern "C" void swing_ring_shape(int* shape) { const int s[] = { NS, UPS, UD, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, NDO }; for (int i = 0; i < 16; i++) shape[i] = s[i]; }

// All arrays in the layouts of include/mpgos_b200_coupled.h: unit scope [serial][system][unit], system scope [serial][system],
// matrix [coupling][row][col] as the user thinks of it.
extern "C" int swing_ring_run(int phases, const double* omega, const double* damping, const double* gain, const double* theta0, const double* matrix,
                              double inertia, int harmonic, double t_end,
                              double* state, double* time, double* time_domain, double* unit_acc, int* int_unit_acc, double* sys_acc, int* int_sys_acc,
                              int* dense_index, double* dense_time, double* dense_state, unsigned long long* attempted, unsigned long long* accepted)
{
	ProblemSolver<NS, UPS, UD, TPB, SPB, NC, CBW, CCI, NUP, NSP, NGP, NiGP, NUA, NiUA, NSA, NiSA, NE, NDO, RKCK45, double> Solver(0);

	Solver.SolverOption(InitialTimeStep, 1.0e-3);
	Solver.SolverOption(ActiveSystems, NS);
	Solver.SolverOption(MinimumTimeStep, 1.0e-12);
	Solver.SolverOption(DenseOutputMinimumTimeStep, 0.05);
	Solver.SolverOption(DenseOutputSaveFrequency, 1);
	Solver.SolverOption(SharedGlobalVariables, 1);
	Solver.SolverOption(SharedCouplingMatrices, 1);
	for (int i = 0; i < UD; i++) { Solver.SolverOption(RelativeTolerance, i, 1.0e-9); Solver.SolverOption(AbsoluteTolerance, i, 1.0e-9); }
	Solver.SolverOption(EventTolerance, 0, 1.0e-7);
	Solver.SolverOption(EventDirection, 0, -1);

	for (int s = 0; s < NS; s++)
	{
		Solver.SetHost(s, TimeDomain, 0, 0.0);
		Solver.SetHost(s, TimeDomain, 1, t_end);
		Solver.SetHost(s, ActualTime, 0.0);
		Solver.SetHost(s, SystemParameters, 0, 0.0);
		Solver.SetHost(s, SystemAccessories, 0, 0.0);
		Solver.SetHost(s, IntegerSystemAccessories, 0, 0);
		Solver.SetHost(s, CouplingStrength, 0, gain[s]);
		Solver.SetHost(s, CouplingStrength, 1, gain[s]);
		Solver.SetHost(s, DenseIndex, 0);
		for (int u = 0; u < UPS; u++)
		{
			Solver.SetHost(s, u, ActualState, 0, theta0[s * UPS + u]);
			Solver.SetHost(s, u, ActualState, 1, 0.0);
			Solver.SetHost(s, u, UnitParameters, 0, omega[u]);
			Solver.SetHost(s, u, UnitParameters, 1, damping[u]);
			Solver.SetHost(s, u, UnitAccessories, 0, 0.0);
			Solver.SetHost(s, u, IntegerUnitAccessories, 0, 0);
		}
	}
	Solver.SetHost(GlobalParameters, 0, inertia);
	Solver.SetHost(IntegerGlobalParameters, 0, harmonic);
	Solver.SetHost(CouplingIndex, 0, 1);
	Solver.SetHost(CouplingIndex, 1, 1);
	for (int k = 0; k < NC; k++)
		for (int r = 0; r < UPS; r++)
			for (int d = -CBW; d <= CBW; d++)
			{
				const int c = ((r + d) % UPS + UPS) % UPS;
				Solver.SetHost(k, CouplingMatrix, r, c, matrix[(k * UPS + r) * UPS + c]);
			}

	Solver.SynchroniseFromHostToDevice(All);
	for (int i = 0; i < phases; i++) Solver.Solve();
	Solver.SynchroniseFromDeviceToHost(All);
	Solver.InsertSynchronisationPoint();
	Solver.SynchroniseSolver();

	const size_t NU = (size_t)NS * UPS;
	for (int s = 0; s < NS; s++)
	{
		time[s] = Solver.GetHost<double>(s, ActualTime);
		time_domain[s] = Solver.GetHost<double>(s, TimeDomain, 0);
		time_domain[s + NS] = Solver.GetHost<double>(s, TimeDomain, 1);
		sys_acc[s] = Solver.GetHost<double>(s, SystemAccessories, 0);
		int_sys_acc[s] = Solver.GetHost<int>(s, IntegerSystemAccessories, 0);
		dense_index[s] = Solver.GetHost<int>(s, DenseIndex);
		for (int p = 0; p < NDO; p++) dense_time[s + (size_t)p * NS] = Solver.GetHost<double>(s, DenseTime, p);
		for (int u = 0; u < UPS; u++)
		{
			for (int i = 0; i < UD; i++) state[u + (size_t)s * UPS + i * NU] = Solver.GetHost<double>(s, u, ActualState, i);
			unit_acc[u + (size_t)s * UPS] = Solver.GetHost<double>(s, u, UnitAccessories, 0);
			int_unit_acc[u + (size_t)s * UPS] = Solver.GetHost<int>(s, u, IntegerUnitAccessories, 0);
			for (int p = 0; p < NDO; p++)
				for (int i = 0; i < UD; i++) dense_state[u + (size_t)s * UPS + i * NU + (size_t)p * NU * UD] = Solver.GetHost<double>(s, u, DenseState, i, p);
		}
	}
	std::vector<unsigned long long> a, b;
	Solver.StepCounters(a, b);
	for (int s = 0; s < NS; s++) { attempted[s] = a[s]; accepted[s] = b[s]; }
	if (Solver.GetHost<double>(1, CouplingMatrix, 5, 7) != matrix[(1 * UPS + 5) * UPS + 7]) return 2;
	Solver.Print(ActualState);
	Solver.Print(CouplingStrength);
	Solver.Print(CouplingMatrix, 0);
	Solver.Print(DenseOutput, 5);
	return 0;
}
