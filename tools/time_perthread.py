"""Development: time thread-per-system workloads (optionally of an A/B library via MPGOS_B200_LIBRARY).
usage: python tools/time_perthread.py [impact:8 kellermiksis:2 lorenz:1 ...]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
GEN = {"lorenz": lambda: W.lorenz(NT=1 << 20), "lorenz_f32": lambda: W.lorenz(NT=1 << 20, dtype=np.float32), "impact": lambda: W.impact_t6(NT=1 << 20),
       "kellermiksis": lambda: W.keller_miksis_t5(NT=1 << 18, ndo=0), "kellermiksis_dense": lambda: W.keller_miksis_t5(NT=1 << 18, ndo=200),
       "kellermiksis_dense_fm": lambda: W.keller_miksis_t5(NT=1 << 18, ndo=200, fast_math=True),
       "duffing": lambda: W.duffing_t1(NT=10000, algorithm=W.RK4)}
lib = os.environ.get("MPGOS_B200_LIBRARY", "(default)").split("/")[-1]
for spec in (sys.argv[1:] or ["impact:8", "kellermiksis:2", "lorenz:1"]):
    name, phases = spec.split(":"); phases = int(phases)
    wl = GEN[name](); sh = wl.shape
    if os.environ.get('REVERSE'):      # the same sweep stored back to front (the expensive end of a sweep may sit on either side)
        wl.buffers = wl.buffers.take(np.arange(wl.NT)[::-1])
    for refill in [int(x) for x in os.environ.get('REFILLS', '0,1').split(',')]:
        for tpb in [int(x) for x in os.environ.get('TPBS', '64,128').split(',')]:
            with pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype) as ps:
                pkg.apply_workload(ps, wl)
                ps.SolverOption(pkg.Options.LaneRefill, refill)
                ps.SolverOption(pkg.Options.ThreadsPerBlock, tpb)
                ps.SolverOption(pkg.Options.DispatchOrder, int(os.environ.get('DISPATCH', '1')))
                best = None
                for rep in range(3):
                    ps.SynchroniseFromHostToDevice(pkg.Variables.All)
                    ps.ResetStepCounters()
                    ms = ps.SolveTimed(phases)
                    best = ms if best is None else min(best, ms)
                att, acc = ps.TotalSteps()
                chk = float(pkg.read_back(ps, wl).ActualState.sum())
            print(f"lib={lib:28s} {name}:{phases} refill={refill} tpb={tpb} dispatch={os.environ.get('DISPATCH', '1')} reverse={os.environ.get('REVERSE', '0')}: {best:8.3f} ms  {att / best / 1e6:7.2f} G steps/s  checksum {chk:.10e}", flush=True)
