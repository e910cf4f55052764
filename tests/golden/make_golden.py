"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own kernel source
(compiled for the host by oracle/Makefile -> oracle/_ref/libref_*.so, from /root/reference).

Run in the development container only (the GPU box has no /root/reference):
    make -C oracle ref && python tests/golden/make_golden.py
The vectors pin the CPU restatement (tests/test_oracle_golden.py) wherever the reference sources
are not available."""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import oracle_py  # noqa: E402
from cases import make_cases, RESULT_FIELDS  # noqa: E402

W = importlib.import_module("massively-parallel-gpu-ode-solver_b200").workloads

for name, (wl, phases) in make_cases(W).items():
    out, att, acc = oracle_py.run_ref(wl, phases=phases)
    data = {f: getattr(out, f) for f in RESULT_FIELDS}
    data["attempted"], data["accepted"] = att, acc
    data["control_in"] = wl.buffers.ControlParameters   # guards against generator drift
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
    print(f"{name}: NT={wl.NT} phases={phases} attempted mean {att.mean():.1f} max {att.max()}")
