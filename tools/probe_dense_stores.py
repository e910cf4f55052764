"""Development: one launch of a dense-output workload, meant to run under
    ncu --metrics l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_op_write.sum,dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum
to see how well the dense-output stores coalesce when the lanes of a warp do NOT reach their storage points in the same iteration.
usage: python tools/probe_dense_stores.py duffing_events | impact_dense | kellermiksis_dense | coupled_ring   [stage]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
W = pkg.workloads
case = sys.argv[1]
if case == "coupled_ring":
    WC = pkg.workloads_coupled
    wl = WC.duffing_ring(NS=8192, CBW=1, CCI=1, ndo=100)
    out, att, acc = pkg.run_coupled_workload(wl, phases=1)
    pts = int(np.minimum(out.DenseIndex, 100).sum())
    print(f"{case}: {wl.buffers.NS} systems x 14 units, {int(att.sum())} system-steps, {pts} stored points x {14 * 2 + 1} words")
else:
    wl = {"duffing_events": lambda: W.duffing_t1(NT=1 << 18, algorithm=W.RKCK45, ndo=200),
          "impact_dense": lambda: W.impact_t6(NT=1 << 18, ndo=200),
          "kellermiksis_dense": lambda: W.keller_miksis_t5(NT=1 << 18, ndo=200),
          "kellermiksis_dense_shuffled": lambda: W.keller_miksis_t5(NT=1 << 18, ndo=200),
          "impact_dense_shuffled": lambda: W.impact_t6(NT=1 << 18, ndo=200),
          "duffing_events_shuffled": lambda: W.duffing_t1(NT=1 << 18, algorithm=W.RKCK45, ndo=200)}[case]()
    if case.endswith("_shuffled"):       # neighbouring lanes integrate unrelated instances: they reach their storage points in different iterations
        wl.buffers = wl.buffers.take(np.random.default_rng(0).permutation(wl.NT))
    out, att, acc = pkg.run_workload(wl, phases=1)
    pts = int(np.minimum(out.DenseIndex, 200).sum())
    same = float(np.mean(out.DenseIndex.reshape(-1, 32).max(1) == out.DenseIndex.reshape(-1, 32).min(1)))
    print(f"{case}: {wl.NT} instances, {int(att.sum())} system-steps, {pts} stored points x {wl.shape.SD + 1} words = {pts * (wl.shape.SD + 1) * 8} bytes; "
          f"warps whose lanes end with the same dense index: {same:.3f}")
