"""Development probe (GPU box): coupled module -- ours vs the reference MPGOS coupled kernels on the same GPU."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
import oracle_py
WC = pkg.workloads_coupled
from cases import make_coupled_cases
cc = make_coupled_cases(WC)
for name, tag in (("bubble_u16_rkck45", "bubble_ensemble_u16"), ("ring_b1c1_rkck45", "duffing_ring_b1c1")):
    _, wl, phases = cc[name]
    ref, ms = oracle_py.run_coupled_ref_gpu(tag, wl, phases=phases)
    out, att, acc = pkg.run_coupled_workload(wl, phases=phases)
    rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
    print(f"{name}: vs reference-on-GPU max rel dX {rel.max():.3e} median {np.median(rel):.3e} dT {np.abs(out.ActualTime-ref.ActualTime).max():.2e}")
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = WC.bubble_ensemble(NS=NS, UPS=64)
for tag, shared in (("bubble_ensemble_u64_s1", 0), ("bubble_ensemble_u64_s1", 1), ("bubble_ensemble_u64_s2", 0), ("bubble_ensemble_u64_s2", 1)):
    t0 = time.time()
    ref, ms = oracle_py.run_coupled_ref_gpu(tag, wl, phases=1, shared_matrices=shared, repeats=2)
    print(f"C5 bubble ensemble {NS} x 64: reference MPGOS {tag} SharedCouplingMatrices={shared}: {ms:.2f} ms  (harness wall {time.time()-t0:.1f} s)")
with pkg.CoupledProblemSolver("bubble_ensemble", NS, 64) as ps:
    pkg.apply_coupled_workload(ps, wl)
    best = None
    for rep in range(3):
        ps.SynchroniseFromHostToDevice(pkg.CoupledVariables.All)
        ps.ResetStepCounters()
        ms2 = ps.SolveTimed(1)
        best = ms2 if best is None else min(best, ms2)
    att, acc = ps.StepCounters()
    out = pkg.read_back_coupled(ps, wl)
rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
steps = int(att.sum())
print(f"C5 ours: {best:.2f} ms, {steps} system-steps ({steps/best/1e3:.3e} system-steps/s, {steps*64/best/1e3:.3e} unit-steps/s), speed-up x{ms/best:.2f}; "
      f"vs reference endpoints max rel {rel.max():.2e} median {np.median(rel):.2e}; flops/step {wl.flops_per_step} -> {steps*wl.flops_per_step/best/1e9:.2f} TFLOP/s")
print("steps per system: min", att.min(), "max", att.max(), "mean", att.mean())
