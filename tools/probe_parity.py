"""Development probe (GPU box): prints how far the CUDA path is from the CPU checker on the
parity cases, and a first timing.  Not a test; output goes to stdout."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
import oracle_py  # noqa: E402
from cases import make_cases  # noqa: E402

W = pkg.workloads
print("fma peak f64 %.2f TF  f32 %.2f TF" % (pkg.probe_fma_peak(0, 8), pkg.probe_fma_peak(0, 4)))
for name, (wl, phases) in make_cases(W).items():
    ref, ratt, racc = oracle_py.run_oracle(wl, phases=phases)
    for refill in (0, 1):
        out, att, acc = pkg.run_workload(wl, phases=phases, lane_refill=refill)
        scale = np.maximum(np.abs(ref.ActualState), 1e-300)
        rel = np.abs(out.ActualState - ref.ActualState) / np.maximum(np.abs(ref.ActualState), 1.0)
        dt = np.abs(out.ActualTime - ref.ActualTime).max()
        print(f"{name:32s} refill={refill} max rel dX {rel.max():.3e} median {np.median(rel):.3e} max dT {dt:.3e} "
              f"steps att diff: n={int((att != ratt).sum())} max={int(np.abs(att - ratt).max())} acc diff n={int((acc != racc).sum())} "
              f"acci diff {int((out.IntegerAccessories != ref.IntegerAccessories).sum())} didx diff {int((out.DenseIndex != ref.DenseIndex).sum())}")

for NT, alg, dt_, nm in ((1 << 20, W.RKCK45, np.float64, "lorenz rkck45 f64"), (1 << 20, W.RKCK45, np.float32, "lorenz rkck45 f32"), (1 << 18, W.RK4, np.float64, "lorenz rk4 f64")):
    wl = W.lorenz(NT=NT, algorithm=alg, dtype=dt_)
    for refill in (0, 1):
        for tpb in (64, 128, 256):
            with pkg.ProblemSolver("lorenz", NT, Algorithm=alg, Precision=dt_) as ps:
                pkg.apply_workload(ps, wl)
                ps.SolverOption(pkg.Options.LaneRefill, refill)
                ps.SolverOption(pkg.Options.ThreadsPerBlock, tpb)
                best = None
                for rep in range(3):
                    ps.SynchroniseFromHostToDevice(pkg.Variables.All)
                    ps.ResetStepCounters()
                    ms = ps.SolveTimed(1)
                    att, acc = ps.TotalSteps()
                    best = ms if best is None else min(best, ms)
                print(f"{nm} NT={NT} refill={refill} tpb={tpb}: {best:.3f} ms, attempted {att} accepted {acc} -> {att / best / 1e6:.1f} G system-steps/s")
