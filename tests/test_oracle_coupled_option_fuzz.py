"""The block-per-system CPU restatement against the reference's own coupled kernels run on the host (oracle/_ref, one
std::thread per CUDA thread) under RANDOM solver options: the coupled controller (CP/ExplicitRungeKutta_ErrorControllers.cuh:
504-539 -- safety 0.8, block-wide accept, clamps, forced acceptance at the minimum step), unit-scope events in all three
directions (CP/EventHandling.cuh), dense output with save frequency and minimum spacing (CP/DenseOutput.cuh), ActiveSystems.
Bit for bit.  The builds between them cover the four coupling-matrix layouts and both algorithms."""
import copy
import os

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(900)      # thread-per-CUDA-thread emulation: a protocol error would otherwise hang the suite

from cases import make_coupled_cases, random_coupled_options, COUPLED_RESULT_FIELDS

NAMES = ["ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c0_rkck45", "ring_b0c1_rk4", "bubble_u8_rkck45", "ring_b1c1_rkck45_f32"]
TRIALS = 3            # compared draws per build
MAX_STEPS = 1500      # the host emulation of the reference costs milliseconds per system-step (a barrier of std::threads per __syncthreads)


@pytest.mark.parametrize("name", NAMES)
def test_coupled_oracle_matches_reference_under_random_options(name, pkg, oracle):
    import importlib
    tag, wl0, phases = make_coupled_cases(importlib.import_module(pkg.__name__ + ".workloads_coupled"))[name]
    if not os.path.exists(oracle.coupled_ref_library_path(tag)):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(20261017 + sum(map(ord, name)))
    compared = 0
    for trial in range(3 * TRIALS):
        if compared == TRIALS:
            break
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.take(np.arange(min(3, wl0.buffers.NS)))
        wl.options = random_coupled_options(rng, wl, trial)
        out, att, acc = oracle.run_coupled_oracle(wl, phases=1)
        if att.sum() > MAX_STEPS:
            continue          # too long for the emulation: next draw
        compared += 1
        ref = oracle.run_coupled_ref(tag, wl, phases=1)
        for f in COUPLED_RESULT_FIELDS:
            assert np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True), f"{name} trial {trial}: {f} ({wl.options})"
        active = wl.options.ActiveSystems
        assert (att[active:] == 0).all() and (att[:active] > 0).all()
    assert compared == TRIALS
