"""The CPU restatement against the host-compiled reference (oracle/_ref) under RANDOM solver options.

The golden vectors pin the restatement at each tutorial's own options; this test walks the option space the
reference exposes through SolverOption (ST/Interface.cuh:932-1082) so that every branch of the controller
(ST/ExplicitRungeKutta_ErrorControllers.cuh:38-85 -- clamps, forced acceptance below 1.01*dtmin), of the event
step control (ST/EventHandling.cuh:21-44 -- all three directions, loose and tight bands) and of the dense
output (ST/DenseOutput.cuh:6-58 -- save frequency, minimum spacing, full buffer) is compared bit for bit.
Needs oracle/_ref (built from /root/reference by `make -C oracle ref`; it travels to the GPU box)."""
import copy

import numpy as np
import pytest

from cases import make_cases, random_options, RESULT_FIELDS

SYSTEMS = ["duffing_rkck45_f64", "duffing_rk4_f64", "impact_rkck45_f64", "lorenz_rkck45_f64", "lorenz_rkck45_f32",
           "kellermiksis_dense_rkck45_f64", "poincare_rkck45_f64"]
TRIALS = 15


@pytest.mark.parametrize("name", SYSTEMS)
def test_oracle_matches_reference_under_random_options(name, workloads, oracle):
    wl0, phases = make_cases(workloads)[name]
    if not oracle.ref_available(wl0):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(20261017 + sum(map(ord, name)))
    idx = np.unique(np.linspace(0, wl0.NT - 1, 24).round().astype(np.int64))
    forced = 0
    for trial in range(TRIALS):
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.take(idx)
        wl.options = random_options(rng, wl, trial)
        n_phases = min(phases, 2)
        ref, ratt, racc = oracle.run_ref(wl, phases=n_phases)
        out, att, acc = oracle.run_oracle(wl, phases=n_phases, double_literals=True)
        for f in RESULT_FIELDS:
            assert np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True), f"{name} trial {trial}: {f} ({wl.options})"
        assert np.array_equal(att, ratt) and np.array_equal(acc, racc), f"{name} trial {trial}: step counts ({wl.options})"
        active = wl.options.ActiveNumberOfThreads
        assert (att[active:] == 0).all() and (att[:active] > 0).all()
        forced += int(trial % 3 == 0)
    assert forced >= TRIALS // 3
