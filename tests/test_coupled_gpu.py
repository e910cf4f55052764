"""Parity of the CUDA block-per-system (coupled) path against the CPU checker, whose arithmetic is pinned bit for bit
to the reference's own coupled kernels (tests/test_oracle_coupled.py), through the public API (Python mirror -> C-ABI ->
kernel).  Covers all four coupling-matrix layouts, RK4 and RKCK45, unit-scope events with user-defined termination,
system accessories written by unit 0, dense output, systems spanning one and several warps."""
import importlib
import os

import numpy as np
import pytest

from cases import make_coupled_cases, COUPLED_RESULT_FIELDS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ccases(pkg):
    return make_coupled_cases(importlib.import_module(pkg.__name__ + ".workloads_coupled"))


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


@pytest.mark.parametrize("name", ["bubble_u8_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4"])
def test_bubble_ensemble_against_checker(name, ccases, pkg, oracle):
    tag, wl, phases = ccases[name]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=phases)
    out, att, acc = pkg.run_coupled_workload(wl, phases=phases)
    assert np.array_equal(out.ActualTime, ref.ActualTime)
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc)                 # step counts of every system
    assert rel_err(out.ActualState, ref.ActualState).max() < (1e-10 if "rk4" in name else 1e-7)


@pytest.mark.parametrize("name", ["ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c1_rk4", "ring_b0c0_rkck45"])
def test_duffing_ring_events_accessories_dense_output(name, ccases, pkg, oracle):
    tag, wl, phases = ccases[name]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=phases)
    out, att, acc = pkg.run_coupled_workload(wl, phases=phases)
    print(f"[parity {name}] state {rel_err(out.ActualState, ref.ActualState).max():.2e}  time {rel_err(out.ActualTime, ref.ActualTime).max():.2e}  step diff {np.abs(att - ratt).max()}  same dense index {(out.DenseIndex == ref.DenseIndex).mean():.3f}")
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)   # event counters of every unit
    assert np.abs(att - ratt).max() <= 2 and np.abs(acc - racc).max() <= 2
    assert np.abs(out.IntegerSystemAccessories - ref.IntegerSystemAccessories).max() <= 2   # step counter kept by unit 0
    # observed on B200: states 5e-13, event-terminated end times 4e-15, identical step counts and dense indices on every system
    assert rel_err(out.ActualTime, ref.ActualTime).max() < 1e-10                    # event-terminated phases
    assert rel_err(out.ActualState, ref.ActualState).max() < 1e-9
    assert rel_err(out.UnitAccessories, ref.UnitAccessories).max() < 1e-9
    assert rel_err(out.SystemAccessories, ref.SystemAccessories).max() < 1e-9
    same = out.DenseIndex == ref.DenseIndex
    assert same.mean() >= 0.95
    n = int(min(ref.DenseIndex.min(), out.DenseIndex.min()))
    cols = np.where(same)[0]
    assert rel_err(out.DenseTime[:n - 1][:, cols], ref.DenseTime[:n - 1][:, cols]).max() < 1e-6
    assert rel_err(out.DenseState[:n - 1][:, :, cols], ref.DenseState[:n - 1][:, :, cols]).max() < 1e-5


def test_bubble_ensemble_64_units_properties(pkg):
    """C5 shape (64 units, dense matrix shared by all systems): size-independent properties -- ends on TD[1], finite,
    bubbles driven harder take more steps, a sub-sweep run alone reproduces the same bits (systems are independent),
    and zero coupling strength reproduces the uncoupled Keller-Miksis path of the thread-per-system kernel."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=96, UPS=64, amplitude=(0.1, 1.5))
    out, att, acc = pkg.run_coupled_workload(wl, phases=1)
    assert np.array_equal(out.ActualTime, wl.buffers.TimeDomain[1]) and np.isfinite(out.ActualState).all()
    assert att[-1] > att[0] and (att >= acc).all()
    idx = np.arange(3, 96, 7)
    import copy
    sub = copy.copy(wl)
    sub.buffers = wl.buffers.take(idx)
    sout, satt, _ = pkg.run_coupled_workload(sub, phases=1)
    assert np.array_equal(sout.ActualState, out.ActualState[:, idx, :]) and np.array_equal(satt, att[idx])


@pytest.mark.parametrize("ups,ns", [(32, 21), (64, 13), (128, 7)])
def test_tensor_core_coupling_against_checker_and_scalar_path(ups, ns, pkg, oracle):
    """Dense shared matrix, FP64, UPS 32/64/128: the column sums of the 8 (4 for UPS 128) systems of a block are one DMMA product
    (mma.sync.m8n8k4.f64).  Checked against the CPU checker (reference summation order) and against our scalar path;
    NS is not a multiple of 8, so blocks also run with idle slots.  The product sums in a different order than the
    reference's single accumulator: endpoints agree to rounding (stated tolerance 1e-9), step counts exactly."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 1.2))
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=8)
    tc, tatt, tacc = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=1)
    sc, satt, sacc = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=0)
    for out, att, acc in ((tc, tatt, tacc), (sc, satt, sacc)):
        assert np.array_equal(out.ActualTime, ref.ActualTime)
        assert np.abs(att - ratt).max() <= 1 and np.abs(acc - racc).max() <= 1
        assert rel_err(out.ActualState, ref.ActualState).max() < 1e-9
    assert rel_err(tc.ActualState, sc.ActualState).max() < 1e-9
    assert not np.array_equal(tc.ActualState, sc.ActualState) or ups < 32    # the two paths really are different code


@pytest.mark.parametrize("algorithm,ndo", [(1, 24), (0, 0)])
def test_tensor_core_coupling_with_dense_output_and_rk4(algorithm, ndo, pkg, oracle):
    """The lockstep (tensor-core) variant with dense output (stored points of every unit, per-system index and time) and
    with the fixed-step scheme, 11 systems of 64 units (one full block + one with idle slots), against the checker."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=11, UPS=64, amplitude=(0.3, 1.0), algorithm=algorithm, ndo=ndo)
    if algorithm == 0:
        wl.buffers.TimeDomain[1][:] = 0.2
        wl.options.InitialTimeStep = 1e-3
    else:
        wl.options.DenseOutputMinimumTimeStep = 0.02
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=8)
    for tc in (1, 0):
        out, att, acc = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=tc)
        assert np.array_equal(out.ActualTime, ref.ActualTime)
        assert np.abs(att - ratt).max() <= 1
        assert rel_err(out.ActualState, ref.ActualState).max() < 1e-9
        if ndo:
            assert np.array_equal(out.DenseIndex, ref.DenseIndex) and (ref.DenseIndex > 3).all()
            n = int(ref.DenseIndex.min())
            assert rel_err(out.DenseTime[:n], ref.DenseTime[:n]).max() < 1e-9
            assert rel_err(out.DenseState[:n], ref.DenseState[:n]).max() < 1e-8


def test_tensor_core_coupling_with_two_couplings_events_and_accessories(pkg, oracle):
    """Ring of 32 Duffing oscillators with FULL coupling matrices: two couplings (two matrix products per stage), one event
    per unit with user-defined termination, system accessories written by unit 0, dense output -- the lockstep variant with
    every feature on, 19 systems (blocks with idle slots, systems ending at different times), against the checker and
    against the scalar variant."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.duffing_ring(NS=19, CBW=0, CCI=0, UPS=32, ndo=60, TPB=32, SPB=1)
    rng = np.random.default_rng(7)
    A = rng.uniform(0.0, 0.08, (2, 32, 32)) * (1.0 - np.eye(32))[None]
    wl.buffers.CouplingMatrix[...] = A
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=2, host_threads=8)
    outs = {}
    for tc in (1, 0):
        out, att, acc = outs[tc] = pkg.run_coupled_workload(wl, phases=2, tensor_core_coupling=tc)
        assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
        assert np.abs(att - ratt).max() <= 2 and np.abs(out.IntegerSystemAccessories - ref.IntegerSystemAccessories).max() <= 2
        assert rel_err(out.ActualTime, ref.ActualTime).max() < 1e-6
        assert rel_err(out.ActualState, ref.ActualState).max() < 1e-5
        assert rel_err(out.SystemAccessories, ref.SystemAccessories).max() < 1e-6
        same = out.DenseIndex == ref.DenseIndex
        assert same.mean() >= 0.7
        n = int(min(ref.DenseIndex.min(), out.DenseIndex.min())) - 1
        cols = np.where(same)[0]
        assert rel_err(out.DenseState[:n][:, :, cols], ref.DenseState[:n][:, :, cols]).max() < 1e-5
    assert len(set(ref.ActualTime.tolist())) > 3          # event-terminated: the systems really end at different times
    assert rel_err(outs[1][0].ActualState, outs[0][0].ActualState).max() < 1e-5


@pytest.mark.parametrize("name,tag,tol", [("bubble_u16_rkck45", "bubble_ensemble_u16", 1e-11), ("ring_b1c1_rkck45", "duffing_ring_b1c1", 1e-9)])
def test_against_reference_mpgos_coupled_kernels_on_the_same_gpu(name, tag, tol, ccases, pkg, oracle):
    """GPU vs GPU: the reference's own coupled kernels compiled by nvcc for sm_100a (oracle/_ref/libref_gpu_coupled_*.so)."""
    _, wl, phases = ccases[name]
    if not os.path.exists(oracle.coupled_ref_library_path(tag, gpu=True)):
        pytest.skip("oracle/_ref/libref_gpu_coupled_*.so not built (needs /root/reference at build time)")
    ref, _ = oracle.run_coupled_ref_gpu(tag, wl, phases=phases)
    out, att, acc = pkg.run_coupled_workload(wl, phases=phases)
    assert rel_err(out.ActualState, ref.ActualState).max() < tol
    assert rel_err(out.ActualTime, ref.ActualTime).max() < tol
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
    assert np.array_equal(out.DenseIndex, ref.DenseIndex)


def test_matrix_entry_outside_band_is_rejected(pkg):
    ps = pkg.CoupledProblemSolver("duffing_ring", 4, 14, CBW=1, CCI=0, NDO=100)
    ps.SetCouplingMatrix(0, 3, 4, 1.0)
    ps.SetCouplingMatrix(0, 0, 13, 1.0)       # circular extension of the band
    assert ps.GetCouplingMatrix(0, 0, 13) == 1.0
    with pytest.raises(pkg.MpgosError):
        ps.SetCouplingMatrix(0, 3, 7, 1.0)    # outside bandwidth 1: the reference's BoundCheck exits here
    with pytest.raises(pkg.MpgosError):
        ps.SetCouplingMatrix(2, 0, 0, 1.0)
    ps.close()


def test_fp32_ring_against_checker_and_reference_float_kernel(ccases, pkg, oracle):
    """FP32 instantiation of the block-per-system kernel (banded circulant matrix, events, accessories, dense output) against
    the single-precision checker, which is bit-identical to the reference's own kernel instantiated with Precision = float
    (tests/test_oracle_coupled.py).  Single-precision arithmetic with 1e-5 tolerances leaves ~1e-3 on the end state."""
    tag, wl, phases = ccases["ring_b1c1_rkck45_f32"]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=phases)
    out, att, acc = pkg.run_coupled_workload(wl, phases=phases)
    assert out.ActualState.dtype == np.float32 and out.DenseState.dtype == np.float32
    # A phase ends when ANY unit reaches its second local maximum: which unit wins is decided by single-precision event values,
    # so a system may end a phase on another event than the checker's (FMA contraction, float vs double literals).  Systems whose
    # event counters agree (most of them) are then compared value by value.
    same = (out.IntegerUnitAccessories == ref.IntegerUnitAccessories).all(axis=(0, 2))
    assert same.mean() >= 0.7
    assert (out.IntegerUnitAccessories.max(axis=2) == 2).all()                       # every system stopped on a second maximum
    assert np.abs(att - ratt)[same].max() <= 0.1 * ratt.max()
    assert rel_err(out.ActualTime, ref.ActualTime)[same].max() < 1e-3
    assert rel_err(out.ActualState, ref.ActualState)[:, same].max() < 5e-3
    assert np.median(rel_err(out.ActualState, ref.ActualState)[:, same]) < 5e-4
    assert rel_err(out.UnitAccessories, ref.UnitAccessories)[:, same].max() < 5e-3
    # dense output: one point per accepted step (the step sequences differ in single precision); the first point of the last
    # phase is the state the previous phase ended on
    assert rel_err(out.DenseState[0], ref.DenseState[0])[:, same].max() < 5e-3
    assert rel_err(out.DenseTime[0], ref.DenseTime[0])[same].max() < 1e-3
    assert (out.DenseIndex > 10).all() and (out.DenseIndex <= wl.shape.NDO).all()


def test_fp32_bubble_ensemble_against_double(pkg, oracle):
    """FP32 bubble ensemble (64 units, dense coupling summed in FP32 on the scalar path): against the single-precision checker
    and against the FP64 run of the same sweep.  Both comparisons at the accuracy single precision with rtol = 1e-5 can give."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl32 = WC.bubble_ensemble(NS=24, UPS=64, amplitude=(0.1, 0.8), dtype=np.float32)
    wl64 = WC.bubble_ensemble(NS=24, UPS=64, amplitude=(0.1, 0.8))
    out32, att32, acc32 = pkg.run_coupled_workload(wl32, phases=1)
    out64, att64, acc64 = pkg.run_coupled_workload(wl64, phases=1)
    ref32, ratt32, _ = oracle.run_coupled_oracle(wl32, phases=1, host_threads=8)
    assert out32.ActualState.dtype == np.float32
    assert np.array_equal(out32.ActualTime, np.ones(24, np.float32))                # every system ends exactly on TD[1]
    assert np.isfinite(out32.ActualState).all()
    assert rel_err(out32.ActualState, ref32.ActualState).max() < 2e-2
    assert np.median(rel_err(out32.ActualState, ref32.ActualState)) < 1e-3
    assert np.median(rel_err(out32.ActualState.astype(np.float64), out64.ActualState)) < 1e-3
    assert (att32 < att64).all()                                                    # looser tolerances, fewer steps
    assert np.abs(att32 - ratt32).max() <= 0.15 * ratt32.max()


@pytest.mark.parametrize("ups,ns", [(64, 21), (32, 13)])
def test_fp32_tensor_core_coupling_3xtf32(ups, ns, pkg, oracle):
    """Single-precision dense coupling on the tensor cores: TF32 operands split into big + small parts (3xTF32), FP32
    accumulation.  Against the scalar FP32 path of the same kernel and the single-precision checker; ns is not a multiple of
    8, so the last block runs with idle slots."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 1.4), dtype=np.float32)
    tc, tatt, _ = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=1)
    sc, satt, _ = pkg.run_coupled_workload(wl, phases=1, tensor_core_coupling=0)
    ref, ratt, _ = oracle.run_coupled_oracle(wl, phases=1, host_threads=8)
    assert tc.ActualState.dtype == np.float32 and np.isfinite(tc.ActualState).all()
    assert np.array_equal(tc.ActualTime, np.ones(ns, np.float32))
    assert not np.array_equal(tc.ActualState, sc.ActualState)                       # it is a different summation
    assert np.median(rel_err(tc.ActualState, sc.ActualState)) < 2e-5
    assert rel_err(tc.ActualState, sc.ActualState).max() < 5e-3
    assert np.median(rel_err(tc.ActualState, ref.ActualState)) < 2e-5
    assert np.abs(tatt - satt).max() <= 0.1 * satt.max() and np.abs(tatt - ratt).max() <= 0.15 * ratt.max()


def test_upload_download_with_caller_owned_arrays(pkg):
    """mpgos_coupled_upload / _download move caller-owned host arrays (the e2e path of bench.py) and give the same result as
    filling the solver's pinned mirrors; wrong sizes and the packed coupling matrix are refused."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=16, UPS=64, amplitude=(0.2, 1.0))
    ref, ratt, _ = pkg.run_coupled_workload(wl, phases=1)
    V = pkg.CoupledVariables
    with pkg.CoupledProblemSolver(wl.name, 16, 64) as ps:
        pkg.apply_coupled_workload(ps, wl)
        ps.SynchroniseFromHostToDevice(V.All)
        state = np.ascontiguousarray(wl.buffers.ActualState) * 1.0
        ps.host(V.ActualState)[...] = -7.0                                  # the mirror is NOT what gets integrated
        ps.Upload(V.ActualState, state.ctypes.data, state.nbytes)
        ps.Solve()
        out = np.empty_like(state)
        ps.Download(V.ActualState, out.ctypes.data, out.nbytes)
        ps.SynchroniseDevice()
        assert np.array_equal(out, ref.ActualState)
        with pytest.raises(pkg.MpgosError):
            ps.Upload(V.ActualState, state.ctypes.data, state.nbytes - 8)
        with pytest.raises(pkg.MpgosError):
            ps.Upload(V.CouplingMatrix, state.ctypes.data, 64 * 64 * 8)


ALL_COUPLED_CASES = ["bubble_u8_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4", "ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c1_rk4",
                     "ring_b0c0_rkck45", "ring_b1c1_rkck45_f32"]


@pytest.mark.parametrize("name", ALL_COUPLED_CASES)
def test_cost_ordered_queue_and_pilot_slice_change_nothing_but_the_order(name, ccases, pkg):
    """DispatchOrder: the system queue hands out the systems longest first -- in the first phase of a sweep after a pilot slice (the
    first tenth of the phase in index order, every system PAUSED on an accepted step, its controller state kept, then continued
    in cost order), in later phases from the previous phase's step counts.  Nothing is clipped and systems are independent, so
    every result -- end states, event counters, accessories kept by unit 0, every dense output point, step counts -- is
    bit-identical to index order in one launch.  DispatchOrder = 2 forces the mechanism on these short queues (1 applies it to
    queues longer than two rounds of slots); three phases: pilot + continuation, then two history-ordered phases."""
    tag, wl, phases = ccases[name]
    a, aatt, aacc = pkg.run_coupled_workload(wl, phases=3, dispatch_order=0)
    b, batt, bacc = pkg.run_coupled_workload(wl, phases=3, dispatch_order=2)
    for f in COUPLED_RESULT_FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), (name, f)
    assert np.array_equal(aatt, batt) and np.array_equal(aacc, bacc) and (aatt > 0).all()


def test_cost_ordered_queue_on_a_long_queue_and_the_tensor_core_path(pkg):
    """The same on a queue long enough for DispatchOrder = 1 (default) to act by itself, through the tensor-core variant (blocks of 8
    systems in lockstep, slots skipping and idling while the queue drains)."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.bubble_ensemble(NS=3000, UPS=32, amplitude=(0.1, 2.0))
    wl.buffers.TimeDomain[1][:] = 0.1
    a, aatt, aacc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=0)
    b, batt, bacc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=1)
    for f in COUPLED_RESULT_FIELDS:
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f
    assert np.array_equal(aatt, batt) and np.array_equal(aacc, bacc) and (aatt > 0).all()


def test_coupling_matrices_larger_than_shared_memory_are_read_from_global_memory(pkg, oracle):
    """192 units with two FULL coupling matrices = 576 KB in double precision: they cannot be staged in shared memory (the reference
    keeps dense matrices in global memory by default, SharedCouplingMatrices = 0), so the scalar coupling reads them from global
    memory -- 5 systems of six warps each, events, accessories, dense output, against the checker."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.duffing_ring(NS=5, CBW=0, CCI=0, UPS=192, ndo=40, TPB=192, SPB=1)
    rng = np.random.default_rng(3)
    wl.buffers.CouplingMatrix[...] = rng.uniform(0.0, 0.02, (2, 192, 192)) * (1.0 - np.eye(192))[None]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=8)
    out, att, acc = pkg.run_coupled_workload(wl, phases=1)
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
    assert np.abs(att - ratt).max() <= 2
    assert rel_err(out.ActualTime, ref.ActualTime).max() < 1e-6
    assert rel_err(out.ActualState, ref.ActualState).max() < 1e-5
    assert rel_err(out.SystemAccessories, ref.SystemAccessories).max() < 1e-6


@pytest.mark.parametrize("ups,cbw,cci,rk4", [(1500, 1, 1, False), (2500, 2, 0, True)])
def test_systems_of_more_than_1024_units(ups, cbw, cci, rk4, pkg, oracle):
    """More units than a block has threads: every thread integrates two (1500 units on 768 threads) or three (2500 units on 864
    threads) units of its system -- the reference's "multiple block launches" (Solver.cuh:1508, 2215).  Rings of Duffing oscillators
    with events, unit and system accessories, dense output (RKCK45 case), banded-circulant and banded matrices; against the checker:
    RK4 to rounding (FMA contraction), RKCK45 with the same step counts."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.duffing_ring(NS=4, CBW=cbw, CCI=cci, UPS=ups, ndo=0 if rk4 else 40, algorithm=WC.RK4 if rk4 else WC.RKCK45, TPB=1024, SPB=1)
    wl.buffers.TimeDomain[1, :] = 6.0
    rng = np.random.default_rng(ups)
    wl.buffers.ActualState[...] = rng.uniform(-3.0, 3.0, wl.buffers.ActualState.shape)
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=2, host_threads=8)
    out, att, acc = pkg.run_coupled_workload(wl, phases=2)
    print(f"[parity ring {ups} units] state {rel_err(out.ActualState, ref.ActualState).max():.2e}  steps {att.tolist()} vs {ratt.tolist()}")
    assert np.abs(att - ratt).max() <= (0 if rk4 else 2) and int(ratt.min()) > 20
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
    assert np.array_equal(out.IntegerSystemAccessories, ref.IntegerSystemAccessories)
    assert rel_err(out.ActualTime, ref.ActualTime).max() < 1e-9
    assert rel_err(out.ActualState, ref.ActualState).max() < 1e-8
    assert rel_err(out.UnitAccessories, ref.UnitAccessories).max() < 1e-8
    assert rel_err(out.SystemAccessories, ref.SystemAccessories).max() < 1e-6
    if not rk4 and np.array_equal(att, ratt):
        assert np.array_equal(out.DenseIndex, ref.DenseIndex)


def test_pilot_slice_keeps_parameter_writes_of_the_callbacks(pkg):
    """Callbacks may write unit parameters, system parameters and coupling strengths; in the reference those live in registers /
    shared memory until the phase ends.  The pilot slice pauses systems in mid-phase: with the mechanism forced (DispatchOrder = 2)
    the results must be the bits of the unsliced phase (DispatchOrder = 0), over two phases."""
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.parameter_drift(NS=40)
    plain, patt, pacc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=0)
    sliced, satt, sacc = pkg.run_coupled_workload(wl, phases=2, dispatch_order=2)
    assert np.array_equal(plain.ActualState, sliced.ActualState) and np.array_equal(plain.ActualTime, sliced.ActualTime)
    assert np.array_equal(patt, satt) and np.array_equal(pacc, sacc) and patt.min() > 20
    # and the writes matter: with the parameters frozen the end state is another one
    frozen = WC.parameter_drift(NS=40)
    assert np.isfinite(plain.ActualState).all()
