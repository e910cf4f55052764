"""Race check (ThreadSanitizer) and memory check (AddressSanitizer + UBSan, second half of the file) of the kernel sources on the
CPU.  Race check of the kernels' shared-memory and barrier protocol: the host-compiled kernel sources
(tests/kernel_on_host, one std::thread per CUDA thread) built with -fsanitize=thread.  ThreadSanitizer sees the emulated
barriers as acquire/release synchronisation, so any access to the kernel's shared memory (function-local statics here) or to
global memory that no barrier orders is reported with the kernel's own file:line.  Covered: slots of one warp (__syncwarp) and of
two warps (named barriers), the lockstep tensor-core variant (block barriers around the matrix product), the system queue, and
the persistent-warp lane mapping of the thread-per-system kernel (ballot + one atomic per refill).  A build whose slot barrier
does not synchronise is the negative control: it must be reported.

Complements compute-sanitizer racecheck on the device (profiles/r1_sanitizer.txt); like it, it checks the runs it sees."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.timeout(900)      # thread-per-CUDA-thread emulation: a protocol error would otherwise hang the suite

HERE = os.path.dirname(os.path.abspath(__file__))
KOH = os.path.join(HERE, "kernel_on_host")
TSANCXX = os.environ.get("TSANCXX", "/usr/bin/g++")


@pytest.fixture(scope="module")
def tsan():
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")) or not os.path.exists(TSANCXX):
        pytest.skip("CUDA toolkit headers or the system g++ not found")
    runtime = subprocess.run([TSANCXX, "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(runtime) or not os.path.exists(runtime):
        pytest.skip("libtsan not installed")
    r = subprocess.run(["make", "-j8", "-C", KOH, "tsan", f"CUDAINC={cuda_inc}", f"TSANCXX={TSANCXX}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return runtime


def run_under_tsan(runtime, build, case, halt=False, pilot_slice=False):
    env = dict(os.environ)
    env["KOH_SLICE"] = "1" if pilot_slice else "0"
    env["LD_PRELOAD"] = runtime
    env["TSAN_OPTIONS"] = (f"halt_on_error={int(halt)} report_signal_unsafe=0 exitcode=66 "
                           f"suppressions={os.path.join(KOH, 'tsan.supp')}")
    r = subprocess.run([sys.executable, os.path.join(KOH, "sanitizer_driver.py"), os.path.join(KOH, "_build", build + ".so"), case],
                       capture_output=True, text=True, env=env, timeout=600)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("build,case", [("tsan_ring", "ring_b1c1_rkck45"), ("tsan_bubble_u64", "bubble:64:3"),
                                        ("tsan_bubble_u32_mma", "bubble:32:5"), ("tsan_impact_refill", "perthread:impact_rkck45_f64"),
                                        ("tsan_ring_u50_upt2", "ring:50:1:1")])      # two units per thread, coupling terms in dynamic shared memory
def test_no_data_race_in_the_kernel_protocols(build, case, tsan):
    code, text = run_under_tsan(tsan, build, case)
    assert "identical to the checker's: True" in text, text[-3000:]
    assert "ThreadSanitizer: data race" not in text, text[:6000]
    assert code == 0, text[-3000:]


def test_no_data_race_with_pilot_slice_and_continuation(tsan):
    """SliceMode 1 then 2: systems paused and resumed through the Slice* arrays in global memory, queue in cost order."""
    code, text = run_under_tsan(tsan, "tsan_ring", "ring_b1c1_rkck45", pilot_slice=True)
    assert "identical to the checker's: True" in text and "ThreadSanitizer: data race" not in text and code == 0, text[:6000]


def test_a_barrier_that_does_not_synchronise_is_reported(tsan):
    code, text = run_under_tsan(tsan, "tsan_ring_broken", "ring_b1c1_rkck45", halt=True)
    assert "ThreadSanitizer: data race" in text and "perblock_kernel.cuh" in text
    assert code == 66


# ---- memory check: the same harnesses under AddressSanitizer + UBSan -------------------------------------------------------------

@pytest.fixture(scope="module")
def asan():
    cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")) or not os.path.exists(TSANCXX):
        pytest.skip("CUDA toolkit headers or the system g++ not found")
    runtime = subprocess.run([TSANCXX, "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(runtime) or not os.path.exists(runtime):
        pytest.skip("libasan not installed")
    r = subprocess.run(["make", "-j8", "-C", KOH, "asan", f"CUDAINC={cuda_inc}", f"TSANCXX={TSANCXX}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return runtime


def run_under_asan(runtime, build, case, shrink_dense=False):
    env = dict(os.environ)
    env["LD_PRELOAD"] = runtime
    env["ASAN_OPTIONS"] = "detect_leaks=0"
    env["KOH_SHRINK_DENSE"] = "1" if shrink_dense else "0"
    r = subprocess.run([sys.executable, os.path.join(KOH, "sanitizer_driver.py"), os.path.join(KOH, "_build", build + ".so"), case],
                       capture_output=True, text=True, env=env, timeout=600)
    return r.returncode, r.stdout + r.stderr


@pytest.mark.parametrize("build,case", [("asan_ring", "ring_b1c1_rkck45"), ("asan_bubble_u32_mma", "bubble:32:5"),
                                        ("asan_duffing", "perthread:duffing_rkck45_f64"), ("asan_impact_refill", "perthread:impact_rkck45_f64")])
def test_no_out_of_bounds_access_or_undefined_behaviour(build, case, asan):
    """Every global array is a numpy allocation made under the preloaded AddressSanitizer runtime, i.e. it has red zones: an index
    past an array (padding lanes, the unguarded first dense-output point, packed coupling matrices, the queue tail) would abort."""
    code, text = run_under_asan(asan, build, case)
    assert code == 0 and "identical to the checker's: True" in text, text[-4000:]
    assert "AddressSanitizer" not in text and "runtime error" not in text


def test_a_dense_output_buffer_that_is_too_short_is_reported(asan):
    code, text = run_under_asan(asan, "asan_duffing", "perthread:duffing_rkck45_f64", shrink_dense=True)
    assert code != 0 and "AddressSanitizer: heap-buffer-overflow" in text and "perthread_kernel.cuh" in text
