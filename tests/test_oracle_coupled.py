"""The coupled (block-per-system) CPU restatement against golden vectors generated from the reference's own coupled
kernels (tests/golden/make_golden_coupled.py) -- bit for bit -- and, where the host build of the reference is present,
against the reference itself on one small case (the emulation runs one std::thread per CUDA thread: slow)."""
import os

import numpy as np
import pytest

from cases import make_coupled_cases, COUPLED_RESULT_FIELDS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["bubble_u8_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4", "ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c1_rk4", "ring_b0c0_rkck45", "ring_b1c1_rkck45_f32"]


@pytest.fixture(scope="module")
def ccases(pkg):
    import importlib
    return make_coupled_cases(importlib.import_module(pkg.__name__ + ".workloads_coupled"))


def test_case_list(ccases):
    assert sorted(ccases) == sorted(NAMES)


@pytest.mark.parametrize("name", NAMES)
def test_coupled_oracle_matches_golden_bit_for_bit(name, ccases, oracle):
    tag, wl, phases = ccases[name]
    gold = np.load(os.path.join(GOLDEN, "coupled_" + name + ".npz"))
    assert np.array_equal(gold["unit_parameters_in"], wl.buffers.UnitParameters), "workload generator drifted from the fixtures"
    assert np.array_equal(gold["coupling_matrix_in"], wl.buffers.CouplingMatrix)
    out, att, acc = oracle.run_coupled_oracle(wl, phases=phases)
    for f in COUPLED_RESULT_FIELDS:
        assert np.array_equal(getattr(out, f), gold[f], equal_nan=True), f"{name}: {f} differs from the reference"
    assert (att >= acc).all() and (acc > 0).all()


def test_coupled_oracle_matches_host_run_reference(ccases, oracle):
    tag, wl, phases = ccases["bubble_u8_rkck45"]
    if not os.path.exists(oracle.coupled_ref_library_path(tag)):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = oracle.run_coupled_ref(tag, wl, phases=phases)
    out, _, _ = oracle.run_coupled_oracle(wl, phases=phases)
    for f in COUPLED_RESULT_FIELDS:
        assert np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True), f


def test_ring_semantics(ccases, oracle):
    """The ring tutorial's observable behaviour: the phase ends when any unit reaches its second local maximum
    (user-defined termination from a unit-scope event), the step counter kept by unit 0 equals the accepted steps,
    min/max dt accessories bracket the initial step, dense output holds one point per accepted step plus the first."""
    tag, wl, phases = ccases["ring_b1c1_rkck45"]
    out, att, acc = oracle.run_coupled_oracle(wl, phases=1)
    assert (out.ActualTime < wl.buffers.TimeDomain[1]).all()
    assert (out.IntegerUnitAccessories[0].max(axis=1) == 2).all()
    assert np.array_equal(out.IntegerSystemAccessories[0], acc)
    assert (out.SystemAccessories[0] <= 1e-2).all() and (out.SystemAccessories[1] >= 1e-2).all()
    assert np.array_equal(out.DenseIndex, np.minimum(acc + 1, wl.shape.NDO))
