"""The product's thread-per-system kernel SOURCE on the CPU: include/mpgos_b200/perthread_kernel.cuh is compiled by g++ as an
ordinary function (tests/kernel_on_host/: a shim for the CUDA built-ins, libm stand-ins for the few inline-PTX primitives of
ptx_primitives.cuh; one "thread" per "block") and run against the checker, whose arithmetic is pinned bit for bit to the
reference's own kernel (tests/test_oracle_golden.py, tests/test_oracle_option_fuzz.py).

What this proves without a GPU: the kernel's control flow and decision logic -- stepper, pow-free controller with its
single-precision screening and exact cold path, event step control, accept/commit order, dense output, phase end -- are the
reference's, at the tutorials' options and under randomly drawn ones.  What it cannot prove: anything about the device
(occupancy, lane mappings, the PTX primitives themselves) -- that is tests/test_perthread_gpu.py.  TEST INFRASTRUCTURE ONLY:
the product has no CPU path and nothing under include/ or the package refers to tests/kernel_on_host.

Two arithmetic modes, both without FMA contraction: the shipped compile-time choices (increment-form stages, one FP64 Newton
step for the step-size factor) and those of the strict library (the reference's expression shapes)."""
import copy
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.timeout(900)      # thread-per-CUDA-thread emulation: a protocol error would otherwise hang the suite

from cases import make_cases, random_options, RESULT_FIELDS

HERE = os.path.dirname(os.path.abspath(__file__))
KOH = os.path.join(HERE, "kernel_on_host")
RK4 = ["lorenz_rk4_f64", "poincare_rk4_f64", "duffing_rk4_f64"]
RKCK45 = ["lorenz_rkck45_f64", "poincare_rkck45_f64", "duffing_rkck45_f64", "impact_rkck45_f64", "kellermiksis_rkck45_f64",
          "kellermiksis_dense_rkck45_f64"]
TRIALS = 12


@pytest.fixture(scope="module")
def koh():
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA toolkit headers not found")
    r = subprocess.run(["make", "-j8", "-C", KOH, "all", f"CUDAINC={cuda_inc}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return KOH


def run_kernel_on_host(koh, oracle, name, wl, phases, strict, refill=False):
    lib = C.CDLL(os.path.join(koh, "_build", f"libkoh_{name}{'_refill' if refill else '_strict' if strict else ''}.so"))
    s = oracle._Shape()
    lib.ref_perthread_shape(C.byref(s))
    oracle._check_shape(s, wl)
    out = wl.buffers.copy()
    cb, co, att, acc, _keep = oracle._pack(wl, out, phases, 1)
    lib.ref_perthread_run.restype = C.c_int
    assert lib.ref_perthread_run(C.byref(cb), C.byref(co)) == 0
    return out, att, acc


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def identical(out, ref):
    return all(np.array_equal(getattr(out, f), getattr(ref, f), equal_nan=True) for f in RESULT_FIELDS)


@pytest.fixture(scope="module")
def cases(workloads):
    return make_cases(workloads)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", RK4)
def test_rk4_kernel_source_is_bit_identical_to_the_reference_arithmetic(name, strict, koh, cases, oracle):
    """Fixed step: no controller arithmetic is involved, so the kernel body reproduces the checker (= the reference's own
    kernel, test_oracle_golden.py) bit for bit -- states, times, accessories, event counters, every dense-output sample."""
    wl, phases = cases[name]
    ref, ratt, racc = oracle.run_oracle(wl, phases=phases)
    out, att, acc = run_kernel_on_host(koh, oracle, name, wl, phases, strict)
    assert identical(out, ref)
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", RKCK45)
def test_rkck45_kernel_source_takes_the_reference_step_sequence(name, strict, koh, cases, oracle):
    """Adaptive: the pow-free step-size factor (and, in the shipped mode, the increment-form stages) differ from the reference's
    arithmetic by rounding; the step sequence of every instance is still the reference's at the tutorials' options."""
    wl, phases = cases[name]
    ref, ratt, racc = oracle.run_oracle(wl, phases=phases)
    out, att, acc = run_kernel_on_host(koh, oracle, name, wl, phases, strict)
    assert np.array_equal(out.ActualTime, ref.ActualTime) or name.startswith("impact")
    assert ((att == ratt) & (acc == racc)).mean() >= (1.0 if strict else 0.97)
    assert rel_err(out.ActualState, ref.ActualState).max() < 1e-8
    assert np.array_equal(out.IntegerAccessories, ref.IntegerAccessories)
    assert np.array_equal(out.DenseIndex, ref.DenseIndex)
    if wl.shape.NA:
        assert rel_err(out.Accessories, ref.Accessories).max() < 1e-8


def test_fp32_kernel_source(koh, cases, oracle):
    wl, phases = cases["lorenz_rkck45_f32"]
    ref, ratt, _ = oracle.run_oracle(wl, phases=phases, double_literals=False)
    for strict in (True, False):
        out, att, _ = run_kernel_on_host(koh, oracle, "lorenz_rkck45_f32", wl, phases, strict)
        assert out.ActualState.dtype == np.float32 and np.array_equal(out.ActualTime, ref.ActualTime)
        assert np.median(rel_err(out.ActualState.astype(np.float64), ref.ActualState.astype(np.float64))) < 1e-4
        assert (att == ratt).mean() >= 0.8 and np.abs(att - ratt).max() <= 0.1 * ratt.max()


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("name", ["duffing_rk4_f64", "duffing_rkck45_f64", "poincare_rkck45_f64", "lorenz_rkck45_f64", "impact_rkck45_f64",
                                  "kellermiksis_dense_rkck45_f64"])
def test_kernel_source_under_random_solver_options(name, strict, koh, cases, oracle):
    """cases.random_options: forced acceptance at the minimum step, the clamps on the step-size factor and on the new step (the
    exact code behind the single-precision screening), all three event directions with loose and tight bands, dense output with
    save frequency > 1 / minimum spacing / full buffer, fewer active instances than allocated.  RK4: bit for bit in every draw.
    RKCK45: the same step sequence on >= 97 % of the instances (a draw may put an accept/reject or event decision on a rounding
    tie, and the impact law then takes another trajectory), endpoints to 1e-6 on those."""
    wl0, phases = cases[name]
    rng = np.random.default_rng(20261017 + sum(map(ord, name)))
    idx = np.unique(np.linspace(0, wl0.NT - 1, 24).round().astype(np.int64))
    n_phases = min(phases, 2)
    same_steps, close, dense_same, events_same = [], [], [], []
    for trial in range(TRIALS):
        wl = copy.copy(wl0)
        wl.buffers = wl0.buffers.take(idx)
        wl.options = random_options(rng, wl, trial)
        active = wl.options.ActiveNumberOfThreads
        ref, ratt, racc = oracle.run_oracle(wl, phases=n_phases)
        out, att, acc = run_kernel_on_host(koh, oracle, name, wl, n_phases, strict)
        assert np.array_equal(out.ActualState[:, active:], wl.buffers.ActualState[:, active:]) and (att[active:] == 0).all()
        if "rk4" in name:
            assert identical(out, ref), (name, trial, wl.options)
            assert np.array_equal(att, ratt) and np.array_equal(acc, racc)
            continue
        same = ((att == ratt) & (acc == racc))[:active]
        same_steps.append(same)
        err = np.maximum(rel_err(out.ActualState, ref.ActualState).max(axis=0), rel_err(out.ActualTime, ref.ActualTime))[:active]
        close.append(err[same] < 1e-6)
        if wl.shape.NDO:
            dense_same.append((out.DenseIndex == ref.DenseIndex)[:active][same])
        if wl.shape.NIA:
            events_same.append((out.IntegerAccessories == ref.IntegerAccessories).all(axis=0)[:active][same])
    if "rk4" in name:
        return
    cat = np.concatenate
    assert cat(same_steps).mean() >= 0.97, (name, cat(same_steps).mean())
    assert cat(close).mean() >= 0.995, (name, cat(close).mean())
    for flags in (dense_same, events_same):
        if flags:
            assert cat(flags).all(), name


@pytest.mark.parametrize("name", ["lorenz_rkck45_f64", "impact_rkck45_f64", "duffing_rk4_f64", "kellermiksis_dense_rkck45_f64"])
def test_lane_refill_mapping_is_bit_identical_on_host(name, koh, cases, oracle):
    """The persistent-warp mapping (LaneRefill: idle lanes counted by a warp ballot, instance ids taken from the queue with one
    atomic per refill; thresholds 1, 5, 12 and 32 idle lanes) run with real 32-thread warps -- one std::thread per CUDA thread,
    two blocks of two warps sharing the queue: every instance is integrated exactly once, the kernel terminates, and each result
    is bit-identical to the direct mapping, also with fewer active instances than allocated."""
    wl0, phases = cases[name]
    for active in (None, 50):
        wl = copy.copy(wl0)
        wl.options = copy.copy(wl0.options)
        wl.options.ActiveNumberOfThreads = active
        direct, datt, dacc = run_kernel_on_host(koh, oracle, name, wl, phases, False)
        refill, ratt, racc = run_kernel_on_host(koh, oracle, name, wl, phases, False, refill=True)
        assert identical(direct, refill), (name, active)
        assert np.array_equal(datt, ratt) and np.array_equal(dacc, racc)
        n = wl.NT if active is None else active
        assert (ratt[:n] > 0).all() and (ratt[n:] == 0).all()


@pytest.mark.parametrize("name", ["lorenz_rk4_f64", "lorenz_rkck45_f64", "duffing_rk4_f64", "duffing_rkck45_f64", "impact_rkck45_f64"])
def test_non_finite_states_follow_the_reference(name, koh, cases, oracle):
    """an initial condition that overflows in the first right-hand side (1e200) or a little later (1e120).
    RK4: message, end of the phase, the non-finite state IS committed (ST/ExplicitRungeKutta_ErrorControllers.cuh:13-19).
    RKCK45: the step is rejected and shrunk by the shrink limit until 1.01*dtmin, then the phase ends with the last finite
    state (:60-70).  Checked three ways: host-compiled reference == checker bit for bit; kernel source == checker in step
    counts, times and NaN pattern, healthy neighbours untouched."""
    wl0, _ = cases[name]
    wl = copy.copy(wl0)
    wl.buffers = wl0.buffers.take(np.arange(16))
    wl.options = copy.copy(wl0.options)
    wl.buffers.ActualState[0, ::4] = 1e200
    wl.buffers.ActualState[0, 1::4] = 1e120
    wl.options.MinimumTimeStep = 1e-6
    chk, catt, cacc = oracle.run_oracle(wl, phases=1, double_literals=True)
    if oracle.ref_available(wl):
        ref, ratt, racc = oracle.run_ref(wl, phases=1)
        assert identical(chk, ref) and np.array_equal(catt, ratt) and np.array_equal(cacc, racc)
    assert not np.isfinite(chk.ActualState[:, ::4]).all() or "rkck45" in name
    for strict in (True, False):
        out, att, acc = run_kernel_on_host(koh, oracle, name, wl, 1, strict)
        assert np.array_equal(att, catt) and np.array_equal(acc, cacc), (name, strict)
        assert np.array_equal(np.isfinite(out.ActualState), np.isfinite(chk.ActualState))
        assert np.array_equal(out.ActualTime[::4], chk.ActualTime[::4])
        assert np.array_equal(out.IntegerAccessories, chk.IntegerAccessories)
        healthy = np.isfinite(chk.ActualState).all(axis=0)
        assert rel_err(out.ActualState[:, healthy], chk.ActualState[:, healthy]).max() < 1e-6
        if "rk4" in name:
            assert identical(out, chk)


FP32_CASES = {
    # name: (workload factory, sub-sample, phases, tolerance on the median endpoint error)
    "poincare_rk4": (lambda W: W.poincare_t3(NT=4096, algorithm=W.RK4), 64, 2, 0.0),
    "duffing_rkck45": (lambda W: W.duffing_t1(NT=4096, algorithm=W.RKCK45), 64, 2, 1e-5),
    "impact_rkck45": (lambda W: W.impact_t6(NT=1 << 14), 64, 4, 1e-4),
    "kellermiksis_rkck45": (lambda W: W.keller_miksis_t5(NT=4096, ndo=0), 64, 1, 1e-5),
}


@pytest.mark.parametrize("name", sorted(FP32_CASES))
def test_fp32_kernel_source_of_the_other_systems(name, koh, workloads, oracle):
    """The single-precision instantiations (packed FFMA2 stages from 3 components on, exact limit compares instead of the
    screening) against the true-FP32 checker with the settings of workloads.single_precision: fixed step bit for bit, adaptive
    at single-precision accuracy (single-precision step sequences decorrelate: step counts within 10 %)."""
    from cases import subsample
    factory, n, phases, tol = FP32_CASES[name]
    wl = workloads.single_precision(subsample(factory(workloads), n))
    ref, ratt, racc = oracle.run_oracle(wl, phases=phases, double_literals=False)
    for strict in (True, False):
        out, att, acc = run_kernel_on_host(koh, oracle, name + "_f32", wl, phases, strict)
        assert out.ActualState.dtype == np.float32 and np.isfinite(out.ActualState).all()
        if "rk4" in name:
            assert identical(out, ref) and np.array_equal(att, ratt)
            continue
        err = rel_err(out.ActualState.astype(np.float64), ref.ActualState.astype(np.float64)).max(axis=0)
        assert np.median(err) < tol, (name, strict, np.median(err))
        assert np.abs(att - ratt).max() <= max(4, 0.1 * ratt.max())


USER_CASES = ["lorenz_rk4_f64", "lorenz_rkck45_f64", "duffing_rk4_f64", "duffing_rkck45_f64", "poincare_rk4_f64", "poincare_rkck45_f64",
              "impact_rkck45_f64", "kellermiksis_rkck45_f64"]


@pytest.mark.parametrize("name", USER_CASES)
def test_reference_tutorial_system_definitions_unmodified_through_our_kernel(name, koh, cases, oracle):
    """The drop-in claim at its most literal, on the CPU: the reference's own *_SystemDefinition.cuh files (included from
    /root/reference at build time, not copied; bound to the kernel as the drop-in header's UserCallbacks binds them) run through
    OUR kernel body, against the same files run through the REFERENCE's kernel (oracle/_ref).  Fixed step: bit for bit.  Adaptive:
    the same step sequence on every instance, endpoints to 1e-9."""
    tut = "/root/reference/Tutorials_SingleSystem_PerThread"
    wl, phases = cases[name]
    if not os.path.isdir(tut) or not oracle.ref_available(wl):
        pytest.skip("needs /root/reference (development container only)")
    r = subprocess.run(["make", "-j8", "-C", koh, "user"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ref, ratt, racc = oracle.run_ref(wl, phases=phases)
    lib = C.CDLL(os.path.join(koh, "_build", f"libkoh_user_{name}.so"))
    shape = oracle._Shape()
    lib.ref_perthread_shape(C.byref(shape))
    oracle._check_shape(shape, wl)
    out = wl.buffers.copy()
    cb, co, att, acc, _keep = oracle._pack(wl, out, phases, 1)
    lib.ref_perthread_run.restype = C.c_int
    assert lib.ref_perthread_run(C.byref(cb), C.byref(co)) == 0
    assert np.array_equal(att, ratt) and np.array_equal(acc, racc)
    assert np.array_equal(out.IntegerAccessories, ref.IntegerAccessories) and np.array_equal(out.DenseIndex, ref.DenseIndex)
    if "rk4" in name:
        assert identical(out, ref)
    else:
        assert rel_err(out.ActualState, ref.ActualState).max() < 1e-9
