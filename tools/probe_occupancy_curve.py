"""Development: cycles per warp-step as a function of resident warps per SM sub-partition (the grid is sized so that
exactly k warps land on each sub-partition): separates 'latency-bound' from 'pipe-bound'.
usage: python tools/probe_occupancy_curve.py [lorenz|impact|kellermiksis]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
name = sys.argv[1] if len(sys.argv) > 1 else "lorenz"
GEN = {"lorenz": lambda n: W.lorenz(NT=n), "impact": lambda n: W.impact_t6(NT=n), "kellermiksis": lambda n: W.keller_miksis_t5(NT=n, ndo=0)}
SMSP = 148 * 4
for k in (1, 2, 3, 4, 5, 6, 8, 12):
    NT = SMSP * 32 * k
    wl = GEN[name](NT); sh = wl.shape
    if os.environ.get("UNIFORM", "1") == "1":      # identical instances: every warp takes the same number of steps, no tail
        wl.buffers.ControlParameters[:, :] = wl.buffers.ControlParameters[:, NT // 2][:, None]
    with pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype) as ps:
        pkg.apply_workload(ps, wl)
        ps.SolverOption(pkg.Options.ThreadsPerBlock, 32)
        best = None
        for rep in range(3):
            ps.SynchroniseFromHostToDevice(pkg.Variables.All)
            ps.ResetStepCounters()
            ms = ps.SolveTimed(1)
            best = ms if best is None else min(best, ms)
        att, acc = ps.TotalSteps()
    warp_steps_per_smsp = att / 32 / SMSP
    print(f"{name} k={k:2d} warps/SMSP NT={NT:7d}: {best:7.3f} ms, {att/best/1e6:7.2f} G steps/s, {best*1e-3*1.965e9/warp_steps_per_smsp:7.1f} cycles per warp-step per SMSP "
          f"(x k = {best*1e-3*1.965e9/warp_steps_per_smsp*k:7.1f} cycles latency of one warp-step)", flush=True)
