"""The product's block-per-system kernel SOURCE on the CPU: include/mpgos_b200/perblock_kernel.cuh compiled by g++ and run with one
std::thread per CUDA thread (tests/kernel_on_host/cuda_block_on_host.h: barriers for __syncthreads / __syncwarp / the named
barriers of the slots, warp shuffles and votes through a per-warp exchange buffer, the FP64 tensor-core instruction rebuilt from
the 32 lanes' fragments), against the coupled checker, whose arithmetic is pinned bit for bit to the reference's own coupled
kernels (tests/test_oracle_coupled.py, tests/test_oracle_coupled_option_fuzz.py).

What this proves without a GPU: the kernel's cooperative logic -- system queue, slots of one and of two warps, shared-memory
coupling terms for all four matrix layouts, the lockstep tensor-core variant with idle slots, the per-system reductions of the
controller and of the event step control, unit-0 bookkeeping of dense output and system accessories, the pilot slice with its
continuation in cost order -- is the reference's algorithm.  TEST INFRASTRUCTURE ONLY: the product has no CPU path, and the
emulation cannot speak for the device (memory model, real barriers, the PTX primitives): that is tests/test_coupled_gpu.py."""
import copy
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(900)      # thread-per-CUDA-thread emulation: a protocol error would otherwise hang the suite

from cases import make_coupled_cases, random_coupled_options, COUPLED_RESULT_FIELDS

HERE = os.path.dirname(os.path.abspath(__file__))
KOH = os.path.join(HERE, "kernel_on_host")
GOLDEN_CASES = ["ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c1_rk4", "ring_b0c0_rkck45", "bubble_u8_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4"]


@pytest.fixture(scope="module")
def koh():
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA toolkit headers not found")
    r = subprocess.run(["make", "-j8", "-C", KOH, "all", f"CUDAINC={cuda_inc}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return KOH


@pytest.fixture(scope="module")
def WC(pkg):
    return importlib.import_module(pkg.__name__ + ".workloads_coupled")


def run_block_kernel_on_host(koh, oracle, build, wl, phases, pilot_slice=False, user=False):
    lib = C.CDLL(os.path.join(koh, "_build", f"libkohb_{'user_' if user else ''}{build}.so"))
    s = oracle._CShape()
    lib.ref_coupled_shape(C.byref(s))
    oracle._ccheck(s, wl)
    out = wl.buffers.copy()
    cb, co, att, acc, _keep = oracle._cpack(wl, out, phases, 1)
    os.environ["KOH_SLICE"] = "1" if pilot_slice else "0"
    try:
        assert lib.ref_coupled_run(C.byref(cb), C.byref(co)) == 0
    finally:
        os.environ.pop("KOH_SLICE", None)
    return out, att, acc


def rel_err(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return (np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max() if a.size else 0.0


def identical(out, ref):
    return all(np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True) for f in COUPLED_RESULT_FIELDS)


def check_against_checker(name, out, att, acc, ref, ratt, racc, tol=1e-9):
    if "rk4" in name:
        assert identical(out, ref), name                       # fixed step: bit for bit
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc), (name, att, ratt)
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
    assert np.array_equal(out.IntegerSystemAccessories, ref.IntegerSystemAccessories)
    assert np.array_equal(out.DenseIndex, ref.DenseIndex)
    assert rel_err(out.ActualState, ref.ActualState) < tol and rel_err(out.ActualTime, ref.ActualTime) < tol
    assert rel_err(out.UnitAccessories, ref.UnitAccessories) < tol and rel_err(out.SystemAccessories, ref.SystemAccessories) < 1e-6


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_block_kernel_source_against_the_checker(name, koh, pkg, WC, oracle):
    """All four coupling-matrix layouts, RK4 and RKCK45, unit-scope events with user-defined termination, system accessories,
    dense output; two blocks of two or three slots pulling systems from the queue."""
    tag, wl, phases = make_coupled_cases(WC)[name]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=phases)
    out, att, acc = run_block_kernel_on_host(koh, oracle, name, wl, phases)
    check_against_checker(name, out, att, acc, ref, ratt, racc)


def test_block_kernel_source_single_precision(koh, WC, oracle):
    tag, wl, phases = make_coupled_cases(WC)["ring_b1c1_rkck45_f32"]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=phases)
    out, att, acc = run_block_kernel_on_host(koh, oracle, "ring_b1c1_rkck45_f32", wl, phases)
    assert out.ActualState.dtype == np.float32
    assert np.abs(att - ratt).max() <= 8 and np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories)
    assert rel_err(out.ActualState, ref.ActualState) < 5e-3 and rel_err(out.ActualTime, ref.ActualTime) < 1e-4


@pytest.mark.parametrize("build,ups,ns", [("bubble_u64_rkck45", 64, 5), ("bubble_u32_rkck45_mma", 32, 11), ("bubble_u64_rkck45_mma", 64, 11)])
def test_block_kernel_source_wide_systems_and_tensor_core_variant(build, ups, ns, koh, WC, oracle):
    """Systems of two warps (the slot synchronises on its named barrier) and the lockstep tensor-core variant: 8 slots per block,
    11 systems (blocks run with idle slots), the DMMA product emulated from the lanes' fragments.  The product sums in another
    order than the reference's single accumulator: endpoints to 1e-12, step counts exactly."""
    wl = WC.bubble_ensemble(NS=ns, UPS=ups, amplitude=(0.2, 1.2))
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=4)
    out, att, acc = run_block_kernel_on_host(koh, oracle, build, wl, 1)
    assert np.array_equal(out.ActualTime, ref.ActualTime)
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc)
    assert rel_err(out.ActualState, ref.ActualState) < 1e-12


@pytest.mark.parametrize("build,ups,cbw,cci,alg", [("ring_u50_upt2", 50, 1, 1, "rkck45"), ("ring_u70_upt3", 70, 2, 0, "rk4")])
def test_several_units_per_thread(build, ups, cbw, cci, alg, koh, WC, oracle):
    """The kernel's UPT > 1 form -- thread t of a slot integrates units t, t + TPS, ... (on the device: systems of more than 1024
    units, the reference's "multiple block launches") -- with 50 units on 32 threads and 70 units on 32 threads: events, unit and
    system accessories, dense output, banded and banded-circulant matrices; against the checker, RK4 bit for bit."""
    wl = WC.duffing_ring(NS=5, CBW=cbw, CCI=cci, UPS=ups, ndo=60, algorithm=WC.RK4 if alg == "rk4" else WC.RKCK45)
    wl.buffers.TimeDomain[1, :] = 8.0
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=2)
    out, att, acc = run_block_kernel_on_host(koh, oracle, build, wl, 2)
    check_against_checker(build + "_" + alg, out, att, acc, ref, ratt, racc)
    assert int(ratt.min()) > 20


@pytest.mark.parametrize("name", ["ring_b1c1_rkck45", "ring_b0c1_rk4", "bubble_u16_rkck45"])
def test_pilot_slice_and_cost_ordered_continuation_change_nothing(name, koh, WC, oracle):
    """SliceMode 1 (every system paused on its first accepted step past a tenth of the phase) followed by SliceMode 2 (the rest, in
    order of decreasing pilot cost) gives the bits of the unsliced phase: controller state, dense-output bookkeeping and event
    values survive the pause."""
    tag, wl, phases = make_coupled_cases(WC)[name]
    whole, watt, wacc = run_block_kernel_on_host(koh, oracle, name, wl, phases)
    sliced, satt, sacc = run_block_kernel_on_host(koh, oracle, name, wl, phases, pilot_slice=True)
    assert identical(whole, sliced)
    assert np.array_equal(watt, satt) and np.array_equal(wacc, sacc)


@pytest.mark.parametrize("name", ["ring_b1c1_rkck45", "ring_b1c0_rkck45", "ring_b0c0_rkck45", "ring_b0c1_rk4", "bubble_u8_rkck45"])
def test_block_kernel_source_under_random_solver_options(name, koh, WC, oracle):
    """cases.random_coupled_options: clamps and forced acceptance of the coupled controller, all three event directions, dense
    output with save frequency / minimum spacing, fewer active systems than allocated.  RK4: bit for bit in every draw; RKCK45: the
    reference's step sequence on at least 90 % of the systems (a draw may put a decision on a rounding tie)."""
    tag, wl0, phases = make_coupled_cases(WC)[name]
    rng = np.random.default_rng(20261017 + sum(map(ord, name)))
    same_steps, close, events = [], [], []
    done = 0
    for trial in range(12):
        if done == 5:
            break
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.copy()
        wl.options = random_coupled_options(rng, wl, trial)
        active = wl.options.ActiveSystems
        ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1)
        if ratt.sum() > 6000:
            continue          # the thread-per-thread emulation costs ~0.3 ms per system-step
        done += 1
        out, att, acc = run_block_kernel_on_host(koh, oracle, name, wl, 1)
        assert np.array_equal(out.ActualState[:, active:, :], wl.buffers.ActualState[:, active:, :]) and (att[active:] == 0).all()
        if "rk4" in name:
            assert identical(out, ref), (name, trial, wl.options)
            assert np.array_equal(att, ratt) and np.array_equal(acc, racc)
            continue
        same = ((att == ratt) & (acc == racc))[:active]
        same_steps.append(same)
        err = np.abs(out.ActualState - ref.ActualState).max(axis=(0, 2)) / np.maximum(np.abs(ref.ActualState).max(axis=(0, 2)), 1.0)
        close.append(err[:active][same] < 1e-6)
        if wl.shape.NiUA:
            events.append((out.IntegerUnitAccessories == ref.IntegerUnitAccessories).all(axis=(0, 2))[:active][same])
        assert np.array_equal(out.DenseIndex[:active][same], ref.DenseIndex[:active][same])
    assert done == 5
    if "rk4" in name:
        return
    cat = np.concatenate
    assert cat(same_steps).mean() >= 0.9, (name, cat(same_steps).mean())
    assert cat(close).all(), name
    if events:
        assert cat(events).all(), name


@pytest.mark.parametrize("name", ["bubble_u16_rk4", "ring_b0c1_rk4", "ring_b1c1_rkck45", "bubble_u8_rkck45"])
def test_non_finite_states_follow_the_reference(name, koh, WC, oracle):
    """One unit of system 0 starts at 1e200.  Coupled RK4: message, end of the phase, the step is NOT committed (unlike the
    thread-per-system module; CP/ExplicitRungeKutta_ErrorControllers.cuh:421-441); RKCK45: rejected and shrunk down to 1.01*dtmin,
    then the phase ends (:504-539).  The reference's own kernels on the host == the checker bit for bit; the kernel source == the
    checker in step counts and in what is finite, the other systems unaffected."""
    tag, wl0, _ = make_coupled_cases(WC)[name]
    wl = copy.copy(wl0)
    wl.buffers = wl0.buffers.copy()
    wl.options = copy.copy(wl0.options)
    wl.buffers.ActualState[0, 0, 1] = 1e200
    wl.options.MinimumTimeStep = 1e-5
    chk, catt, cacc = oracle.run_coupled_oracle(wl, phases=1)
    if os.path.exists(oracle.coupled_ref_library_path(tag)):
        ref = oracle.run_coupled_ref(tag, wl, phases=1)
        assert identical(chk, ref)
    assert cacc[0] == 0 and (cacc[1:] > 0).all()
    out, att, acc = run_block_kernel_on_host(koh, oracle, name, wl, 1)
    assert np.array_equal(att, catt) and np.array_equal(acc, cacc)
    assert np.array_equal(np.isfinite(out.ActualState), np.isfinite(chk.ActualState))
    assert np.array_equal(out.ActualState[:, 0, :], chk.ActualState[:, 0, :], equal_nan=True)       # system 0: never advanced
    assert rel_err(out.ActualState[:, 1:, :], chk.ActualState[:, 1:, :]) < 1e-9
    if "rk4" in name:
        assert identical(out, chk)


def test_tensor_core_variant_with_two_couplings_events_and_dense_output(koh, WC, oracle):
    """Ring of 32 Duffing oscillators with two FULL coupling matrices (two matrix products per stage), one event per unit with
    user-defined termination, system accessories written by unit 0, dense output: the lockstep tensor-core variant (8 slots, 5
    systems: idle slots, systems leaving the lockstep at different times) and the scalar variant (3 slots) against the checker."""
    wl = WC.duffing_ring(NS=5, CBW=0, CCI=0, UPS=32, ndo=60, TPB=32, SPB=1)
    rng = np.random.default_rng(7)
    wl.buffers.CouplingMatrix[...] = rng.uniform(0.0, 0.08, (2, 32, 32)) * (1.0 - np.eye(32))[None]
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=4)
    assert len(set(ref.ActualTime.tolist())) == 5
    for build in ("ring_u32_rkck45_mma", "ring_u32_rkck45"):
        out, att, acc = run_block_kernel_on_host(koh, oracle, build, wl, 1)
        check_against_checker(build, out, att, acc, ref, ratt, racc)
        n = int(ref.DenseIndex.min())
        assert rel_err(out.DenseTime[:n], ref.DenseTime[:n]) < 1e-9 and rel_err(out.DenseState[:n], ref.DenseState[:n]) < 1e-7


def test_single_precision_tensor_core_variant_3xtf32(koh, WC, oracle):
    """FP32 kernel, dense coupling as 3xTF32 (big + small TF32 parts, mma.m16n8k8 with FP32 accumulation, rebuilt here from the
    lanes' fragments with cvt.rna rounding) and the scalar FP32 variant, against the single-precision checker: same step counts,
    endpoints at single-precision accuracy -- the split keeps FP32 accuracy, a plain TF32 product (10 mantissa bits) would not."""
    wl = WC.bubble_ensemble(NS=11, UPS=32, amplitude=(0.2, 1.2), dtype=np.float32)
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=1, host_threads=4)
    for build in ("bubble_u32_rkck45_mma_f32", "bubble_u32_rkck45_f32"):
        out, att, acc = run_block_kernel_on_host(koh, oracle, build, wl, 1)
        assert out.ActualState.dtype == np.float32 and np.array_equal(out.ActualTime, ref.ActualTime)
        assert np.abs(att - ratt).max() <= 1 and np.abs(acc - racc).max() <= 1
        assert rel_err(out.ActualState, ref.ActualState) < 1e-5


@pytest.mark.parametrize("name", ["ring_b1c1_rkck45", "ring_b0c1_rk4", "ring_b0c0_rkck45", "bubble_u16_rkck45", "bubble_u16_rk4"])
def test_reference_coupled_tutorial_definitions_unmodified_through_our_kernel(name, koh, WC, oracle):
    """The reference's own coupled *_SystemDefinition.cuh files (ring of Duffing oscillators, bubble ensemble; included from
    /root/reference at build time, not copied; bound as the drop-in header's UserCoupledCallbacks binds them) through OUR block
    kernel, against the same files through the REFERENCE's coupled kernels run on the host (oracle/_ref): fixed step bit for bit,
    adaptive with the same step sequence of every system."""
    tag, wl0, phases = make_coupled_cases(WC)[name]
    if not os.path.isdir("/root/reference/Tutorials_CoupledSystems_PerBlock") or not os.path.exists(oracle.coupled_ref_library_path(tag)):
        pytest.skip("needs /root/reference (development container only)")
    r = subprocess.run(["make", "-j8", "-C", koh, "user"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    wl = copy.copy(wl0)
    wl.buffers = wl0.buffers.take(np.arange(min(3, wl0.buffers.NS)))
    ref = oracle.run_coupled_ref(tag, wl, phases=1)
    chk, catt, cacc = oracle.run_coupled_oracle(wl, phases=1)        # step counts (the reference keeps none); bit-identical to `ref`
    assert identical(chk, ref)
    out, att, acc = run_block_kernel_on_host(koh, oracle, name, wl, 1, user=True)
    assert np.array_equal(att, catt) and np.array_equal(acc, cacc)
    assert np.array_equal(out.IntegerUnitAccessories, ref.IntegerUnitAccessories) and np.array_equal(out.DenseIndex, ref.DenseIndex)
    if "rk4" in name:
        assert identical(out, ref)
    else:
        assert rel_err(out.ActualState, ref.ActualState) < 1e-9 and rel_err(out.ActualTime, ref.ActualTime) < 1e-12
