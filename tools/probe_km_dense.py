"""Development: Keller-Miksis with dense output against the checker, phase by phase (optionally an A/B library via MPGOS_B200_LIBRARY)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
import oracle_py
from cases import make_cases
def rel(a, b): return np.abs(a - b) / np.maximum(np.abs(b), 1.0)
wl, _ = make_cases(pkg.workloads)["kellermiksis_dense_rkck45_f64"]
for phases in (1, 2):
    ref, ratt, racc = oracle_py.run_oracle(wl, phases=phases)
    out, att, acc = pkg.run_workload(wl, phases=phases)
    n = int(ref.DenseIndex.min())
    dt = rel(out.DenseTime[:n], ref.DenseTime[:n]); ds = rel(out.DenseState[:n], ref.DenseState[:n])
    print(f"lib={os.environ.get('MPGOS_B200_LIBRARY', '(default)').split('/')[-1]:32s} phases={phases}: endpoint {rel(out.ActualState, ref.ActualState).max():.2e}  accessories {rel(out.Accessories, ref.Accessories).max():.2e}"
          f"  same steps {(att == ratt).mean():.3f}  dense index equal {(out.DenseIndex == ref.DenseIndex).mean():.3f}  dense time {dt.max():.2e} (first point {dt[0].max():.2e})  dense state {ds.max():.2e}"
          f"  worst instance {int(dt.max(axis=0).argmax())}  worst point {int(dt.max(axis=1).argmax())} of {n}")
