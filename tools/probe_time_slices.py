"""Development: would a short pilot slice predict the cost of the rest of a phase?  Emulated with two phases: [0, f T] in index
order, then [f T, T] in the order the pilot's step counts give (DispatchOrder), against the single phase [0, T].
usage: python tools/probe_time_slices.py [fraction ...]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
wl = W.lorenz(NT=1 << 20)
V = pkg.Variables
fractions = [float(x) for x in sys.argv[1:]] or [0.05, 0.1, 0.2]
with pkg.ProblemSolver(wl.name, wl.NT, NDO=0, Algorithm=wl.shape.Algorithm, Precision=wl.shape.dtype) as ps:
    pkg.apply_workload(ps, wl)
    ps.SolverOption(pkg.Options.ThreadsPerBlock, 64)
    T1 = float(wl.buffers.TimeDomain[1, 0])
    def run(frac, dispatch):
        ps.SolverOption(pkg.Options.DispatchOrder, dispatch)
        best = None
        for rep in range(3):
            ps.host(V.TimeDomain)[1, :] = T1 if frac is None else frac * T1
            ps.SynchroniseFromHostToDevice(V.All)
            ps.ResetStepCounters()
            ms = ps.SolveTimed(1)
            if frac is not None:
                ps.host(V.TimeDomain)[0, :] = frac * T1
                ps.host(V.TimeDomain)[1, :] = T1
                ps.SynchroniseFromHostToDevice(V.TimeDomain)
                ms += ps.SolveTimed(1)
                ps.host(V.TimeDomain)[0, :] = 0.0
            best = ms if best is None else min(best, ms)
        return best, ps.TotalSteps()[0]
    print("single phase            : %.3f ms (%d steps)" % run(None, 1))
    for f in fractions:
        a = run(f, 0); b = run(f, 1)
        print("pilot %4.2f  index order: %.3f ms   cost order: %.3f ms (%d steps)" % (f, a[0], b[0], b[1]))
