"""Golden vectors of the coupled (block-per-system) module: outputs of the REFERENCE's own coupled kernels run on the
host with one std::thread per CUDA thread (oracle/ref_coupled_harness.cpp -> oracle/_ref/libref_coupled_*.so).
Development container only:   make -C oracle refcoupled && python tests/golden/make_golden_coupled.py"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import oracle_py  # noqa: E402
from cases import make_coupled_cases, COUPLED_RESULT_FIELDS  # noqa: E402

WC = importlib.import_module("massively-parallel-gpu-ode-solver_b200.workloads_coupled")
for name, (tag, wl, phases) in make_coupled_cases(WC).items():
    out = oracle_py.run_coupled_ref(tag, wl, phases=phases)
    data = {f: getattr(out, f) for f in COUPLED_RESULT_FIELDS}
    data["unit_parameters_in"] = wl.buffers.UnitParameters
    data["coupling_matrix_in"] = wl.buffers.CouplingMatrix
    np.savez_compressed(os.path.join(HERE, "coupled_" + name + ".npz"), **data)
    print(name, "NS", wl.NS, "end times", out.ActualTime[:3])
