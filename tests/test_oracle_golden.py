"""The CPU restatement (oracle/) against the golden vectors generated from the reference's own
kernel source (tests/golden/make_golden.py), and -- where the host-compiled reference libraries
are present -- against the reference itself.  Bit-exact: both are compiled without FMA contraction."""
import os

import numpy as np
import pytest

from cases import make_cases, RESULT_FIELDS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASE_NAMES = ["lorenz_rk4_f64", "lorenz_rkck45_f64", "lorenz_rkck45_f32", "poincare_rk4_f64", "poincare_rkck45_f64",
              "duffing_rk4_f64", "duffing_rkck45_f64", "impact_rkck45_f64", "kellermiksis_rkck45_f64",
              "kellermiksis_dense_rkck45_f64"]


@pytest.fixture(scope="module")
def cases(workloads):
    return make_cases(workloads)


def test_case_list_matches_generator(cases):
    assert sorted(cases) == sorted(CASE_NAMES)
    for n in CASE_NAMES:
        assert os.path.exists(os.path.join(GOLDEN, n + ".npz")), n


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_golden_bit_for_bit(name, cases, oracle):
    wl, phases = cases[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(gold["control_in"], wl.buffers.ControlParameters), "workload generator drifted from the fixtures"
    # FP32: the reference's system definitions use double literals; the restatement reproduces that mode exactly
    out, att, acc = oracle.run_oracle(wl, phases=phases, double_literals=True)
    for f in RESULT_FIELDS:
        assert np.array_equal(getattr(out, f), gold[f], equal_nan=True), f"{name}: {f} differs from the reference"
    assert np.array_equal(att, gold["attempted"])
    assert np.array_equal(acc, gold["accepted"])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_matches_host_compiled_reference(name, cases, oracle):
    wl, phases = cases[name]
    if not oracle.ref_available(wl):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref, ratt, racc = oracle.run_ref(wl, phases=phases)
    out, att, acc = oracle.run_oracle(wl, phases=phases, double_literals=True)
    for f in RESULT_FIELDS:
        assert np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True), f"{name}: {f}"
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc)


def test_known_invariants(cases, oracle):
    # Lorenz RK4 dt=1e-3 on [0,10] takes 10 001 steps (10 000 + one clipped end step)
    wl, phases = cases["lorenz_rk4_f64"]
    out, att, acc = oracle.run_oracle(wl, phases=phases)
    assert (att == 10001).all() and (acc == 10001).all()
    assert np.array_equal(out.ActualTime, wl.buffers.TimeDomain[1])
    # adaptive runs end exactly on the end of the time domain, attempted >= accepted
    wl, phases = cases["lorenz_rkck45_f64"]
    out, att, acc = oracle.run_oracle(wl, phases=phases)
    assert np.array_equal(out.ActualTime, wl.buffers.TimeDomain[1]) and (att >= acc).all()


def test_multithreaded_oracle_is_deterministic(cases, oracle):
    wl, phases = cases["impact_rkck45_f64"]
    a, att1, _ = oracle.run_oracle(wl, phases=phases, host_threads=1)
    b, att4, _ = oracle.run_oracle(wl, phases=phases, host_threads=4)
    assert np.array_equal(a.ActualState, b.ActualState) and np.array_equal(att1, att4)


def test_inactive_instances_untouched(cases, oracle, workloads):
    import copy
    wl, phases = cases["lorenz_rkck45_f64"]
    wl = copy.copy(wl)
    wl.options = copy.copy(wl.options)
    wl.options.ActiveNumberOfThreads = 50
    out, att, _ = oracle.run_oracle(wl, phases=1)
    assert np.array_equal(out.ActualState[:, 50:], wl.buffers.ActualState[:, 50:]) and (att[50:] == 0).all() and (att[:50] > 0).all()
