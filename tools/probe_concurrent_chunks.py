"""Development: C2 as ONE launch of 2^20 instances against K solver objects of 2^20 / K instances launched back to back on their own
streams (what the e2e loop does): wall clock around launch + synchronise, minimum of 30 repetitions, state restored before each."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
NT = 1 << 20
wl = W.lorenz(NT=NT)
def make(idx):
    sub = wl.buffers.take(idx); sh = wl.shape
    ps = pkg.ProblemSolver(wl.name, sub.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype)
    pkg.apply_workload(ps, wl, sub)
    ps.SynchroniseFromHostToDevice(pkg.Variables.All); ps.InsertSynchronisationPoint(); ps.SynchroniseSolver(); ps.SaveState()
    return ps
for K in (1, 2, 4, 8):
    for layout in ("slabs", "interleaved"):
        if K == 1 and layout == "interleaved": continue
        parts = [np.arange(NT)[k * (NT // K):(k + 1) * (NT // K)] if layout == "slabs" else np.arange(NT).reshape(-1, K, 4096 // K if False else 1)[:, k, 0] for k in range(K)] if layout == "slabs" else \
                [np.arange(NT).reshape(NT // (4096 * K), K, 4096)[:, k, :].reshape(-1) for k in range(K)]
        solvers = [make(p) for p in parts]
        best = 1e9
        for rep in range(30):
            for ps in solvers: ps.RestoreState(); ps.InsertSynchronisationPoint()
            for ps in solvers: ps.SynchroniseSolver()
            t0 = time.perf_counter()
            for ps in solvers: ps.Solve(); ps.InsertSynchronisationPoint()
            for ps in solvers: ps.SynchroniseSolver()
            best = min(best, time.perf_counter() - t0)
        single = min(solvers[0].SolveTimed(1) for _ in range(1)) if False else None
        print(f"K={K} {layout:12s}: {best * 1e3:.3f} ms wall (launch + synchronise)", flush=True)
        for ps in solvers: ps.close() if hasattr(ps, "close") else None
