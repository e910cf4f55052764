"""The CUDA block-per-system path against the CPU checker under RANDOM solver options (cases.random_coupled_options -- the
draws that pin the checker bit for bit against the reference's own coupled kernels in tests/test_oracle_coupled_option_fuzz.py).
Runs the coupled controller's clamps and forced acceptance at the minimum step (CP/ExplicitRungeKutta_ErrorControllers.cuh:
504-539), unit-scope events in all three directions (CP/EventHandling.cuh), dense output with save frequency / minimum spacing
(CP/DenseOutput.cuh) and ActiveSystems < NS on the device, for the four coupling-matrix layouts, RK4 and RKCK45, FP64.

Gates as in tests/test_zz_options_perthread_gpu.py: GPU and CPU differ by rounding-level perturbations per step, so a system whose
accept/reject or event decision sits on a tie may take another step sequence; a semantic difference would move every system."""
import copy
import importlib

import numpy as np
import pytest

from cases import make_coupled_cases, random_coupled_options

pytestmark = pytest.mark.gpu

NAMES = ["ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c0_rkck45", "ring_b0c1_rk4", "bubble_u8_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4"]
TRIALS = 6
FRACTION = 0.85


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


@pytest.mark.parametrize("name", NAMES)
def test_random_solver_options_against_the_coupled_checker(name, pkg, oracle):
    tag, wl0, phases = make_coupled_cases(importlib.import_module(pkg.__name__ + ".workloads_coupled"))[name]
    rng = np.random.default_rng(20261017 + sum(map(ord, name)))
    state_ok, steps_ok, events_same, dense_same, state_err = [], [], [], [], []
    for trial in range(TRIALS):
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.copy()
        wl.options = random_coupled_options(rng, wl, trial)
        active = wl.options.ActiveSystems
        ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=4)
        if ratt.sum() > 200000:
            continue        # a fixed-step draw that would only cost time
        out, att, acc = pkg.run_coupled_workload(wl, phases=1)
        # inactive systems: untouched, bit for bit
        assert np.array_equal(out.ActualState[:, active:, :], wl.buffers.ActualState[:, active:, :]), (name, trial)
        assert (att[active:] == 0).all() and (att[:active] > 0).all() and (att >= acc).all(), (name, trial)
        err = np.maximum(rel_err(out.ActualState, ref.ActualState).max(axis=(0, 2)), rel_err(out.ActualTime, ref.ActualTime))[:active]
        state_err.append(err)
        state_ok.append(err < 1e-5)
        steps_ok.append((np.abs(att - ratt) <= 2)[:active])
        if wl.shape.NiUA:
            events_same.append((out.IntegerUnitAccessories == ref.IntegerUnitAccessories).all(axis=(0, 2))[:active])
        if wl.shape.NDO:
            dense_same.append((np.abs(out.DenseIndex - ref.DenseIndex) <= 2)[:active])
    cat = np.concatenate
    assert len(state_ok) >= TRIALS // 2
    assert cat(state_ok).mean() >= FRACTION, (name, "endpoints", cat(state_ok).mean())
    assert np.median(cat(state_err)) < 1e-8, (name, "median endpoint error", np.median(cat(state_err)))
    assert cat(steps_ok).mean() >= FRACTION, (name, "step counts", cat(steps_ok).mean())
    if events_same:
        assert cat(events_same).mean() >= FRACTION, (name, "event counters", cat(events_same).mean())
    if dense_same:
        assert cat(dense_same).mean() >= FRACTION, (name, "dense index", cat(dense_same).mean())
