"""Development: time the C5 coupled kernel (optionally of an A/B library via MPGOS_B200_LIBRARY)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
WC = pkg.workloads_coupled
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = WC.bubble_ensemble(NS=NS, UPS=64)
with pkg.CoupledProblemSolver("bubble_ensemble", NS, 64) as ps:
    pkg.apply_coupled_workload(ps, wl)
    ps.SolverOption(pkg.CoupledOptions.TensorCoreCoupling, int(os.environ.get('TC', '1')))
    best = None
    for rep in range(4):
        ps.SynchroniseFromHostToDevice(pkg.CoupledVariables.All)
        ps.ResetStepCounters()
        ms = ps.SolveTimed(1)
        best = ms if best is None else min(best, ms)
    att, acc = ps.StepCounters()
    out = pkg.read_back_coupled(ps, wl)
print(f"TC={os.environ.get('TC','1')} lib={os.environ.get('MPGOS_B200_LIBRARY','(default)'):34s} NS={NS}: {best:.2f} ms  {int(att.sum())/best/1e3:.3f} M system-steps/s  checksum {out.ActualState.sum():.12e}")
