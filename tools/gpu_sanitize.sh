# compute-sanitizer over small parity cases: memcheck (out-of-bounds, misaligned) and racecheck (shared-memory hazards)
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import importlib, sys, os
import numpy as np
sys.path[:0] = [os.getcwd(), "oracle", "tests"]
pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
from cases import make_cases, make_coupled_cases
WC = pkg.workloads_coupled
which = sys.argv[1]
if which == "coupled":
    for ups, ns, tc in ((64, 13, 1), (32, 21, 1), (64, 5, 0)):
        wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 1.2)); wl.buffers.TimeDomain[1][:] = 0.05
        out, att, acc = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=tc); print("bubble", ups, ns, tc, int(att.sum()))
    cc = make_coupled_cases(WC)
    for name in ("ring_b1c1_rkck45", "ring_b0c0_rkck45", "bubble_u16_rkck45"):
        tag, wl, phases = cc[name]; out, att, acc = pkg.run_coupled_workload(wl, phases=1); print(name, int(att.sum()))
    tag, wl, phases = cc["ring_b1c1_rkck45_f32"]; out, att, acc = pkg.run_coupled_workload(wl, phases=1); print("ring f32", int(att.sum()))
    wl = WC.bubble_ensemble(NS=7, UPS=64, amplitude=(0.2, 1.2), dtype=np.float32); wl.buffers.TimeDomain[1][:] = 0.05
    out, att, acc = pkg.run_coupled_workload(wl, phases=1); print("bubble f32 64", int(att.sum()))
elif which == "dispatch":
    # launches large enough for the cost-ordered dispatch kernels to run (second phase), short time domains
    wl = pkg.workloads.impact_t6(NT=1 << 18)
    out, att, acc = pkg.run_workload(wl, phases=2); print("impact dispatch order", int(att.sum()))
    wl = WC.bubble_ensemble(NS=2600, UPS=16, amplitude=(0.1, 2.0)); wl.buffers.TimeDomain[1][:] = 0.02
    out, att, acc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=2); print("bubble pilot slice + queue order", int(att.sum()))
    cc = make_coupled_cases(WC)
    for name in ("ring_b1c1_rkck45", "ring_b0c0_rkck45"):      # multi-warp slots, events, dense output: pause / skip / resume paths
        tag, wl, phases = cc[name]; out, att, acc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=2); print(name, "pilot slice", int(att.sum()))
    wl = WC.bubble_ensemble(NS=21, UPS=64, amplitude=(0.2, 1.2)); wl.buffers.TimeDomain[1][:] = 0.05
    out, att, acc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=2); print("bubble 64 tensor-core pilot slice", int(att.sum()))
elif which == "round2":
    # round 2: fast-math forms (subroutines, constant table), one solver object over several parts (strided copies, global ids),
    # coupling matrices in global memory, the pilot slice with parameter writes
    W = pkg.workloads
    from cases import subsample
    wl = subsample(W.keller_miksis_t5(NT=4096, ndo=200, fast_math=True), 96)
    out, att, acc = pkg.run_workload(wl, phases=2); print("kellermiksis_dense_fm", int(att.sum()))
    for ups, ns, tc in ((64, 13, 1), (16, 9, 0)):
        wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 1.2), fast_math=True); wl.buffers.TimeDomain[1][:] = 0.05
        out, att, acc = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=tc); print("bubble fm", ups, ns, tc, int(att.sum()))
    wl = W.tidprobe(NT=5000)
    out, att, acc = pkg.run_workload(wl, phases=2, devices=[0, 0, 0], tile=256); print("tidprobe over 3 parts, ragged tiles", int(att.sum()))
    wl = subsample(W.impact_t6(NT=1 << 14), 3000)
    out, att, acc = pkg.run_workload(wl, phases=3, devices=[0, 0], tile=512); print("impact over 2 parts", int(att.sum()))
    wl = WC.duffing_ring(NS=23, CBW=1, CCI=1)
    out, att, acc = pkg.run_coupled_workload(wl, phases=2, devices=[0, 0, 0], tile=4); print("ring over 3 parts", int(att.sum()))
    wl = WC.duffing_ring(NS=3, CBW=0, CCI=0, UPS=192, ndo=40, TPB=192, SPB=1); wl.buffers.TimeDomain[1][:] = wl.buffers.TimeDomain[0][:] + 2.0
    out, att, acc = pkg.run_coupled_workload(wl, phases=1); print("ring 192 units, matrices in global memory", int(att.sum()))
    wl = WC.parameter_drift(NS=40)
    out, att, acc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=2); print("parameter_drift pilot slice", int(att.sum()))
else:
    c = make_cases(pkg.workloads)
    for name in ("lorenz_rkck45_f64", "duffing_rkck45_f64", "impact_rkck45_f64", "kellermiksis_dense_rkck45_f64", "lorenz_rkck45_f32"):
        wl, phases = c[name]
        for refill in (0, 1, 8):
            out, att, acc = pkg.run_workload(wl, phases=1, lane_refill=refill); print(name, refill, int(att.sum()))
PY
for tool in memcheck racecheck; do
  for which in ${SANITIZE_CASES:-coupled perthread dispatch round2}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py $which > gpurun_out/sanitize_${tool}_${which}.txt 2>&1
    echo "== $tool $which: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_${tool}_${which}.txt | sort | uniq -c | head -5; grep -cE "^(bubble|ring|lorenz|duffing|impact|keller|tidprobe|parameter)" gpurun_out/sanitize_${tool}_${which}.txt
  done
done
