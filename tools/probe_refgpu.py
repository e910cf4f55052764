"""Development probe (GPU box): our CUDA path against the REFERENCE MPGOS compiled for sm_100a --
endpoint differences on the parity cases and kernel time on the BASELINE sizes."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
import oracle_py
from cases import make_cases
W = pkg.workloads
for name, (wl, phases) in make_cases(W).items():
    if not oracle_py.ref_gpu_available(wl):
        print(name, "no reference GPU library"); continue
    ref, ms = oracle_py.run_ref_gpu(wl, phases=phases)
    out, att, acc = pkg.run_workload(wl, phases=phases)
    rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
    print(f"{name:32s} vs reference-on-GPU: max rel dX {rel.max():.3e} median {np.median(rel):.3e} bit-identical states {(out.ActualState == ref.ActualState).mean():.3f} "
          f"dT {np.abs(out.ActualTime - ref.ActualTime).max():.2e} acci diff {int((out.IntegerAccessories != ref.IntegerAccessories).sum())} didx diff {int((out.DenseIndex != ref.DenseIndex).sum())}")
for gen, phases, label in ((lambda: W.lorenz(NT=1 << 20), 1, "C2 lorenz rkck45 f64 2^20"),
                           (lambda: W.lorenz(NT=1 << 20, dtype=np.float32), 1, "C2 lorenz rkck45 f32 2^20"),
                           (lambda: W.lorenz(NT=1 << 20, algorithm=W.RK4), 1, "lorenz rk4 f64 2^20"),
                           (lambda: W.impact_t6(NT=1 << 20), 8, "C3 impact 2^20, 8 phases"),
                           (lambda: W.keller_miksis_t5(NT=1 << 18, ndo=0), 2, "C4 kellermiksis 2^18 (no dense), 2 phases"),
                           (lambda: W.keller_miksis_t5(NT=1 << 18, ndo=200), 2, "C4 kellermiksis dense 2^18, 2 phases")):
    wl = gen()
    if not oracle_py.ref_gpu_available(wl):
        continue
    for tpb in (64, 128):
        ref, ms = oracle_py.run_ref_gpu(wl, phases=phases, threads_per_block=tpb, repeats=3)
        print(f"{label}: reference MPGOS on this GPU, {tpb} threads/block: {ms:.3f} ms")
    for refill in (0, 1):
        sh = wl.shape
        with pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype) as ps:
            pkg.apply_workload(ps, wl)
            ps.SolverOption(pkg.Options.LaneRefill, refill)
            best = None
            for rep in range(3):
                ps.SynchroniseFromHostToDevice(pkg.Variables.All)
                ps.ResetStepCounters()
                ms2 = ps.SolveTimed(phases)
                best = ms2 if best is None else min(best, ms2)
            att, acc = ps.TotalSteps()
            out = pkg.read_back(ps, wl)
        rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
        print(f"{label}: ours refill={refill}: {best:.3f} ms ({att / best / 1e6:.1f} G steps/s), speed-up x{ms / best:.2f}; vs reference endpoints: max rel {rel.max():.2e} median {np.median(rel):.2e} "
              f"frac>1e-6 {(rel.max(axis=0) > 1e-6).mean():.2e} acci diff {int((out.IntegerAccessories != ref.IntegerAccessories).sum())}")
