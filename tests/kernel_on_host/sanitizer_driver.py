"""TEST INFRASTRUCTURE ONLY -- runs one host-compiled kernel library that was built with -fsanitize=thread (Makefile target
`tsan`) or -fsanitize=address,undefined (target `asan`) on a small case; started by tests/test_kernel_sanitizers_on_host.py in a subprocess with libtsan preloaded.  ThreadSanitizer
sees the emulated CUDA threads as what they are here -- std::threads -- and the emulated barriers as acquire/release
synchronisation, so an access to shared memory (function-local statics of the kernel) or to global memory that is not ordered
by a barrier is reported as a data race with the kernel's own file:line.
usage: sanitizer_driver.py <library> <case>"""
import copy
import ctypes as C
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import oracle_py as O  # noqa: E402
from cases import make_cases, make_coupled_cases  # noqa: E402

pkg = importlib.import_module("massively-parallel-gpu-ode-solver_b200")
lib, case = C.CDLL(sys.argv[1]), sys.argv[2]
if case.startswith("perthread:"):
    wl0, phases = make_cases(pkg.workloads)[case.split(":", 1)[1]]
    wl = copy.copy(wl0)
    wl.buffers = wl0.buffers.take(np.arange(0, wl0.NT, max(wl0.NT // 80, 1)))
    out = wl.buffers.copy()
    if os.environ.get("KOH_SHRINK_DENSE") == "1":      # negative control of the AddressSanitizer run: a dense-output buffer a tenth as long as declared
        out.DenseState = out.DenseState[: out.DenseState.shape[0] // 10].copy()
    cb, co, att, acc, _k = O._pack(wl, out, 1, 1)
    assert lib.ref_perthread_run(C.byref(cb), C.byref(co)) == 0
    ref, ratt, racc = O.run_oracle(wl, phases=1)
else:
    WC = pkg.workloads_coupled
    if case.startswith("bubble:"):
        ups, ns = map(int, case.split(":")[1:])
        wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 0.8))
    elif case.startswith("ring:"):      # ring:UPS:CBW:CCI -- several units per thread (tests/test_kernel_block_on_host.py)
        ups, cbw, cci = map(int, case.split(":")[1:])
        wl = WC.duffing_ring(NS=5, CBW=cbw, CCI=cci, UPS=ups, ndo=60)
        wl.buffers.TimeDomain[1, :] = 4.0
    else:
        tag, wl0, phases = make_coupled_cases(WC)[case]
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.take(np.arange(min(5, wl0.buffers.NS)))
    out = wl.buffers.copy()
    cb, co, att, acc, _k = O._cpack(wl, out, 1, 1)
    assert lib.ref_coupled_run(C.byref(cb), C.byref(co)) == 0
    ref, ratt, racc = O.run_coupled_oracle(wl, phases=1)
print("steps", int(att.sum()), "identical to the checker's:", bool(np.array_equal(att, ratt)))
