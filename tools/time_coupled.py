"""Development: time the C5 coupled kernel (optionally of an A/B library via MPGOS_B200_LIBRARY)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
WC = pkg.workloads_coupled
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
UPS = int(os.environ.get('UPS', '64'))
DT = np.float32 if os.environ.get('F32') else np.float64
wl = WC.bubble_ensemble(NS=NS, UPS=UPS, dtype=DT, fast_math=bool(os.environ.get('FM')))
if os.environ.get('REVERSE'):      # the same sweep stored back to front
    wl.buffers = wl.buffers.take(np.arange(NS)[::-1])
with pkg.CoupledProblemSolver(wl.name, NS, UPS, Precision=DT) as ps:
    pkg.apply_coupled_workload(ps, wl)
    ps.SolverOption(pkg.CoupledOptions.TensorCoreCoupling, int(os.environ.get('TC', '1')))
    ps.SolverOption(pkg.CoupledOptions.DispatchOrder, int(os.environ.get('DISPATCH', '1')))
    PH = int(os.environ.get('PHASES', '1'))
    best = None
    for rep in range(4):
        ps.SynchroniseFromHostToDevice(pkg.CoupledVariables.All)
        ps.ResetStepCounters()
        ms = ps.SolveTimed(PH)
        best = ms if best is None else min(best, ms)
    att, acc = ps.StepCounters()
    out = pkg.read_back_coupled(ps, wl)
if os.environ.get('REFTAG'):
    sys.path[:0] = [os.path.join(ROOT, 'oracle')]
    import oracle_py
    for shared in (0, 1):
        ref, rms = oracle_py.run_coupled_ref_gpu(os.environ['REFTAG'], wl, phases=1, shared_matrices=shared, repeats=2)
        rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
        print(f'reference MPGOS {os.environ["REFTAG"]} SharedCouplingMatrices={shared}: {rms:.2f} ms; ours vs reference max rel {rel.max():.2e}')
print(f"phases={os.environ.get('PHASES','1')} dispatch={os.environ.get('DISPATCH','1')} reverse={os.environ.get('REVERSE','0')} UPS={UPS} TC={os.environ.get('TC','1')} lib={os.environ.get('MPGOS_B200_LIBRARY','(default)'):34s} NS={NS}: {best:.2f} ms  {int(att.sum())/best/1e3:.3f} M system-steps/s  checksum {out.ActualState.sum():.12e}")
