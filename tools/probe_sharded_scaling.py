"""Development: ONE solver object over n devices of the node (mpgos_perthread_create_sharded): C2 Lorenz with 2^20 instances per device
(weak scaling) and 2^20 instances in total (strong scaling); kernel time = the slowest device (mpgos_perthread_solve_timed)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
ndev = torch.cuda.device_count()
base = None
for mode in ("weak", "strong"):
    for n in [d for d in (1, 2, 4, 8) if d <= ndev]:
        NT = (1 << 20) * (n if mode == "weak" else 1)
        wl = W.lorenz(NT=NT); sh = wl.shape
        with pkg.ProblemSolver(wl.name, NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype, Devices=list(range(n)), Tile=4096) as ps:
            pkg.apply_workload(ps, wl)
            best = None
            for rep in range(4):
                ps.SynchroniseFromHostToDevice(pkg.Variables.All); ps.ResetStepCounters()
                ms = ps.SolveTimed(1); best = ms if best is None else min(best, ms)
            att, acc = ps.TotalSteps()
            out = pkg.read_back(ps, wl)
        rate = att / best / 1e-3
        if base is None: base = rate
        print(f"one solver object over {n} device(s), {mode} scaling: {NT} instances, {best:.3f} ms, {rate:.4e} system-steps/s ({rate / base / (n if mode == 'weak' else 1):.3f} of n x one device)"
              f"  checksum {float(out.ActualState.sum()):.10e}", flush=True)
