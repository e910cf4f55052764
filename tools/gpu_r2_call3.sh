# round 2, GPU call 3: fast-math accuracy + parity, A/B of the "_fm" built-ins, full bench line
mkdir -p gpurun_out
( time python -m pytest tests/test_fast_math_gpu.py tests/test_sharded_solver.py tests/test_perthread_gpu.py tests/test_coupled_gpu.py -m gpu -x -q -s ) > gpurun_out/r2c3_pytest.txt 2>&1; tail -4 gpurun_out/r2c3_pytest.txt; grep "fast_math" gpurun_out/r2c3_pytest.txt
python - <<'PY'
import importlib, os, sys, time
import numpy as np
sys.path[:0] = [os.getcwd()]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
for fm in (False, True):
    wl = W.keller_miksis_t5(NT=1 << 18, ndo=200, fast_math=fm); sh = wl.shape
    for tpb in (64, 128):
        with pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype) as ps:
            pkg.apply_workload(ps, wl); ps.SolverOption(pkg.Options.ThreadsPerBlock, tpb)
            best = None
            for rep in range(3):
                ps.SynchroniseFromHostToDevice(pkg.Variables.All); ps.ResetStepCounters()
                ms = ps.SolveTimed(2); best = ms if best is None else min(best, ms)
            att, acc = ps.TotalSteps()
            print(f"kellermiksis_dense fast_math={fm} tpb={tpb}: {best:8.3f} ms for 2 phases, {att / best / 1e6:6.2f} G steps/s", flush=True)
WC = pkg.workloads_coupled
for fm in (False, True):
    wl = WC.bubble_ensemble(NS=8192, UPS=64, fast_math=fm); sh = wl.shape
    for tc in (1, 0):
        with pkg.CoupledProblemSolver(wl.name, wl.buffers.NS, sh.UPS, Algorithm=sh.Algorithm, TPB=sh.TPB, SPB=sh.SPB) as ps:
            pkg.apply_coupled_workload(ps, wl); ps.SolverOption(pkg.CoupledOptions.TensorCoreCoupling, tc)
            best = None
            for rep in range(3):
                ps.SynchroniseFromHostToDevice(pkg.CoupledVariables.All); ps.ResetStepCounters()
                ms = ps.SolveTimed(1); best = ms if best is None else min(best, ms)
            att, acc = ps.StepCounters()
            print(f"coupled bubbles fast_math={fm} tensor_core={tc}: {best:8.3f} ms, {att.sum() / best / 1e3:6.2f} M system-steps/s", flush=True)
PY
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; tail -4 gpurun_out/r2c3_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2c3_bench.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("HEADLINE value %.4e kernel_ms %.4f frac %.4f peak %.2f e2e %.4e refgpu x%s" % (d["value"], r["kernel_ms"], r["frac"], r["peak"], d["e2e"]["value"], d.get("reference_gpu",{}).get("speedup_kernel")))
    for k,c in d.get("configs",{}).items():
        if "error" in c: print(k, c); continue
        r=c["roofline"]; g=c.get("reference_gpu",{})
        print(k, "value %.4e kernel_ms %.4f frac %.4f e2e %.4e refgpu x%s (%s) cpu %s" % (c["value"], r["kernel_ms"], r["frac"], c["e2e"]["value"], g.get("speedup_kernel"), g.get("best_setting"), c.get("cpu_baseline",{}).get("value")))
except Exception as e: print("bench unreadable", e)
PY
