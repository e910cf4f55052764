"""Development: pilot-slice emulation for the coupled module (C5): [0, f] in index order, then [f, 1] in the queue order the pilot's
step counts give, against the single phase [0, 1].  usage: python tools/probe_time_slices_coupled.py [fraction ...]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
WC = pkg.workloads_coupled
NS = 8192
wl = WC.bubble_ensemble(NS=NS, UPS=64)
V = pkg.CoupledVariables
fractions = [float(x) for x in sys.argv[1:]] or [0.05, 0.1]
with pkg.CoupledProblemSolver("bubble_ensemble", NS, 64) as ps:
    pkg.apply_coupled_workload(ps, wl)
    def run(frac, dispatch):
        ps.SolverOption(pkg.CoupledOptions.DispatchOrder, dispatch)
        best = None
        for rep in range(2):
            ps.host(V.TimeDomain)[0, :] = 0.0
            ps.host(V.TimeDomain)[1, :] = 1.0 if frac is None else frac
            ps.SynchroniseFromHostToDevice(V.All)
            ps.ResetStepCounters()
            ms = ps.SolveTimed(1)
            if frac is not None:
                ps.host(V.TimeDomain)[0, :] = frac
                ps.host(V.TimeDomain)[1, :] = 1.0
                ps.SynchroniseFromHostToDevice(V.TimeDomain)
                ms += ps.SolveTimed(1)
            best = ms if best is None else min(best, ms)
        return best, int(ps.StepCounters()[0].sum())
    print("single phase            : %.3f ms (%d steps)" % run(None, 1))
    for f in fractions:
        a = run(f, 0); b = run(f, 1)
        print("pilot %4.2f  index order: %.3f ms   cost order: %.3f ms (%d steps)" % (f, a[0], b[0], b[1]))
