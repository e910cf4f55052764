"""The drop-in claim as a test: every tutorial program of the reference compiles UNMODIFIED against include/ (same include
order, same template arguments, same unscoped enums and method overloads).  Needs the reference checkout and nvcc, i.e. the
development container; skipped elsewhere.  Compile only (-c): running them is the GPU tests' job."""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _tutorials():
    out = []
    for group in ("Tutorials_SingleSystem_PerThread", "Tutorials_CoupledSystems_PerBlock"):
        for d in sorted(glob.glob(os.path.join(REF, group, "T*"))):
            mains = [f for f in glob.glob(os.path.join(d, "*.cu"))]
            out += [(os.path.basename(d), f) for f in mains]
    return out


FAST_MATH = ["-DMPGOS_B200_FAST_MATH", "-include", "mpgos_b200/fast_math.cuh"]


def _compile(item, extra=()):
    name, src = item
    with tempfile.TemporaryDirectory() as tmp:
        cmd = ["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", os.path.join(ROOT, "include"), *extra,
               "-I", os.path.dirname(src), "-c", src, "-o", os.path.join(tmp, "t.o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r.returncode, r.stderr[-1500:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "SourceCodes")) or shutil.which("nvcc") is None,
                    reason="needs /root/reference and nvcc (development container)")
def test_every_reference_tutorial_compiles_unmodified_against_the_dropin_headers():
    items = _tutorials()
    assert len(items) >= 10          # T1..T8 per-thread, T1..T2 coupled
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(_compile, items))
    failed = [(n, err) for n, rc, err in results if rc != 0]
    assert not failed, failed


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "SourceCodes")) or shutil.which("nvcc") is None,
                    reason="needs /root/reference and nvcc (development container)")
def test_every_reference_tutorial_compiles_unmodified_with_the_fast_math_redirect():
    """The opt-in transcendentals for UNMODIFIED programs: -DMPGOS_B200_FAST_MATH -include mpgos_b200/fast_math.cuh turns the calls of
    pow / sin / cos / sincos / sincospi into the fast functions in device code and leaves host code (the tutorials' main() programs
    compute parameters with pow, the bubble tutorial includes <random>) on the C library."""
    items = _tutorials()
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        results = list(pool.map(lambda it: _compile(it, FAST_MATH), items))
    failed = [(n, err) for n, rc, err in results if rc != 0]
    assert not failed, failed


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "SourceCodes")) or shutil.which("nvcc") is None,
                    reason="needs /root/reference and nvcc (development container)")
def test_ring_tutorial_compiles_in_single_precision():
    """The coupled module's FP32 instantiation through the drop-in header: the reference's ring tutorial with its own
    `#define PRECISION` switched to float (the one edit its comment invites: "// float, double")."""
    src = os.path.join(REF, "Tutorials_CoupledSystems_PerBlock", "T1_RingOfDuffingOscillators", "RingOfDuffingOscillators.cu")
    text = open(src).read()
    assert "#define PRECISION double" in text
    with tempfile.TemporaryDirectory() as tmp:
        f32 = os.path.join(tmp, "RingOfDuffingOscillators_f32.cu")
        open(f32, "w").write(text.replace("#define PRECISION double", "#define PRECISION float"))
        cmd = ["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-w", "-I", os.path.join(ROOT, "include"),
               "-I", os.path.dirname(src), "-c", f32, "-o", os.path.join(tmp, "t.o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
