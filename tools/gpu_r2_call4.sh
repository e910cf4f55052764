# round 2, GPU call 4: whole GPU suite (tutorial programs on the device, large coupling matrices, pilot slice with parameter writes, fast math),
# register-cap variants of the fast-math Keller-Miksis kernel, dense-output stores with the sweep shuffled
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -s ) > gpurun_out/r2c4_pytest.txt 2>&1; tail -6 gpurun_out/r2c4_pytest.txt; grep -E "^\[|\[tutorial|\[fast_math\] (K|b)" gpurun_out/r2c4_pytest.txt | head -20
python - <<'PY'
import importlib, os, sys
sys.path[:0] = [os.getcwd()]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
for lib in ["", "build/variants/libmpgos_b200_km96.so", "build/variants/libmpgos_b200_km80.so"]:
    os.environ["MPGOS_B200_LIBRARY"] = lib
    import subprocess
    code = """
import importlib, os, sys
sys.path[:0] = [os.getcwd()]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
for fm in (False, True):
    wl = W.keller_miksis_t5(NT=1 << 18, ndo=200, fast_math=fm); sh = wl.shape
    with pkg.ProblemSolver(wl.name, wl.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype) as ps:
        pkg.apply_workload(ps, wl); ps.SolverOption(pkg.Options.ThreadsPerBlock, 128)
        best = None
        for rep in range(3):
            ps.SynchroniseFromHostToDevice(pkg.Variables.All); ps.ResetStepCounters()
            ms = ps.SolveTimed(2); best = ms if best is None else min(best, ms)
        att, acc = ps.TotalSteps()
        print("lib=%-40s kellermiksis_dense fast_math=%s: %8.3f ms for 2 phases, %6.2f G steps/s" % (os.environ.get("MPGOS_B200_LIBRARY", ""), fm, best, att / best / 1e6), flush=True)
"""
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MPGOS_B200_LIBRARY=lib))
PY
M="l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum"
for c in duffing_events_shuffled impact_dense_shuffled kellermiksis_dense_shuffled; do
  ncu --metrics $M --clock-control none -k regex:perthread_kernel --csv --log-file gpurun_out/r2c4_dense_$c.csv python tools/probe_dense_stores.py $c > gpurun_out/r2c4_dense_$c.txt 2>&1; tail -1 gpurun_out/r2c4_dense_$c.txt
done
