"""The C-ABI library: it loads, exports every symbol include/mpgos_b200.h declares, knows the
built-in systems, validates arguments, and fails loudly (no fallback) without a CUDA device.
No compute calls here -- those are the -m gpu tests."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(name="mpgos_b200.h"):
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpgos_[a-z0-9_]+)\s*\(", text)) - {"mpgos_perthread_launcher", "mpgos_coupled_launcher"})


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mpgos_b200.h but not exported"
    assert sorted(pkg.ABI_SYMBOLS) == declared
    assert lib.mpgos_abi_version() == 3


def test_library_exports_every_declared_coupled_symbol(pkg):
    lib = pkg.load_library()
    declared = header_functions("mpgos_b200_coupled.h")
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mpgos_b200_coupled.h but not exported"
    assert sorted(pkg.COUPLED_ABI_SYMBOLS) == declared
    strict = pkg.load_library(strict=True)
    for name in declared + header_functions():
        assert hasattr(strict, name)


def test_coupled_builtins_and_descriptor_checks(pkg):
    from importlib import import_module
    sc = import_module(pkg.__name__ + ".solver_coupled")
    lib = pkg.load_library()
    sc._bind(lib)
    d = sc._CDesc()
    assert lib.mpgos_coupled_builtin_desc(b"bubble_ensemble", 64, 0, 0, C.byref(d)) == 0
    assert (d.UD, d.NC, d.NUP, d.NE) == (2, 1, 21, 0)
    assert lib.mpgos_coupled_builtin_desc(b"duffing_ring", 14, 1, 1, C.byref(d)) == 0
    assert (d.UD, d.NC, d.NUP, d.NGP, d.NUA, d.NiUA, d.NSA, d.NiSA, d.NE) == (2, 2, 3, 1, 3, 1, 2, 1, 1)
    assert lib.mpgos_coupled_builtin_desc(b"bubble_ensemble", 63, 0, 0, C.byref(d)) == -5
    assert lib.mpgos_coupled_builtin_launcher(b"bubble_ensemble", 64, 0, 0, 1, 0)
    assert lib.mpgos_coupled_builtin_launcher(b"duffing_ring", 14, 1, 1, 1, 1)
    assert not lib.mpgos_coupled_builtin_launcher(b"duffing_ring", 14, 1, 1, 1, 0)
    h = C.c_void_p()
    bad = sc._CDesc(NS=4, UPS=14, UD=2, NC=2, CBW=9, CCI=0, PrecisionBytes=8, Algorithm=1)
    assert lib.mpgos_coupled_create(C.byref(bad), C.byref(h)) == -1            # band wider than the ring
    bad = sc._CDesc(NS=4, UPS=14, UD=2, NC=2, PrecisionBytes=2, Algorithm=1)
    assert lib.mpgos_coupled_create(C.byref(bad), C.byref(h)) == -1            # double or float only
    # FP32 instantiations of the coupled module are looked up with the precision stated
    assert lib.mpgos_coupled_builtin_launcher_p(b"bubble_ensemble", 64, 0, 0, 1, 0, 4)
    assert lib.mpgos_coupled_builtin_launcher_p(b"duffing_ring", 14, 1, 1, 1, 1, 4)
    assert lib.mpgos_coupled_builtin_launcher_p(b"bubble_ensemble", 64, 0, 0, 1, 0, 8) == lib.mpgos_coupled_builtin_launcher(b"bubble_ensemble", 64, 0, 0, 1, 0)
    assert not lib.mpgos_coupled_builtin_launcher_p(b"bubble_ensemble", 128, 0, 0, 1, 0, 4)


def test_builtin_systems_and_shapes(pkg):
    lib = pkg.load_library()
    names = [lib.mpgos_perthread_builtin_name(i).decode() for i in range(lib.mpgos_perthread_builtin_count())]
    assert names == ["lorenz", "poincare", "duffing", "impact", "kellermiksis", "kellermiksis_dense", "tidprobe", "kellermiksis_fm", "kellermiksis_dense_fm"]
    from importlib import import_module
    solver = import_module(pkg.__name__ + ".solver")
    expect = {"lorenz": (3, 1, 0, 0, 0, 0, 0), "poincare": (2, 1, 1, 0, 0, 0, 0), "duffing": (2, 1, 1, 0, 2, 3, 1),
              "impact": (3, 1, 4, 0, 2, 3, 2), "kellermiksis": (2, 13, 0, 0, 0, 1, 0), "kellermiksis_dense": (2, 13, 0, 0, 0, 1, 0), "tidprobe": (1, 1, 0, 0, 0, 1, 2),
              "kellermiksis_fm": (2, 13, 0, 0, 0, 1, 0), "kellermiksis_dense_fm": (2, 13, 0, 0, 0, 1, 0)}
    for n, shape in expect.items():
        d = solver._Desc()
        assert lib.mpgos_perthread_builtin_desc(n.encode(), C.byref(d)) == 0
        assert (d.SD, d.NCP, d.NSP, d.NISP, d.NE, d.NA, d.NIA) == shape
    d = solver._Desc()
    assert lib.mpgos_perthread_builtin_desc(b"nonesuch", C.byref(d)) == -5
    assert b"nonesuch" in lib.mpgos_last_error()
    # launchers exist for the BASELINE configurations, not for unknown combinations
    assert lib.mpgos_perthread_builtin_launcher(b"lorenz", 1, 8, 0)
    assert lib.mpgos_perthread_builtin_launcher(b"lorenz", 1, 4, 0)
    assert lib.mpgos_perthread_builtin_launcher(b"duffing", 0, 8, 1)
    assert lib.mpgos_perthread_builtin_launcher(b"impact", 1, 8, 0)
    assert lib.mpgos_perthread_builtin_launcher(b"kellermiksis_dense", 1, 8, 1)
    assert not lib.mpgos_perthread_builtin_launcher(b"lorenz", 7, 8, 0)


def test_create_rejects_bad_descriptors(pkg):
    lib = pkg.load_library()
    from importlib import import_module
    solver = import_module(pkg.__name__ + ".solver")
    h = C.c_void_p()
    d = solver._Desc(NT=0, SD=3, PrecisionBytes=8, Algorithm=1)
    assert lib.mpgos_perthread_create(C.byref(d), C.byref(h)) == -1 and not h.value
    d = solver._Desc(NT=8, SD=3, PrecisionBytes=2, Algorithm=1)
    assert lib.mpgos_perthread_create(C.byref(d), C.byref(h)) == -1
    d = solver._Desc(NT=8, SD=3, PrecisionBytes=8, Algorithm=5)
    assert lib.mpgos_perthread_create(C.byref(d), C.byref(h)) == -1
    assert lib.mpgos_perthread_create(None, C.byref(h)) == -1


def test_no_cpu_fallback(pkg):
    """Without a CUDA device, creating a solver must raise -- never silently compute elsewhere."""
    lib = pkg.load_library()
    if lib.mpgos_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(pkg.MpgosError):
        pkg.ProblemSolver("lorenz", 64)
    with pytest.raises(pkg.MpgosError):
        pkg.run_workload(pkg.workloads.lorenz(NT=64))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or include/ may reference it."""
    pkgdir = os.path.join(ROOT, "massively-parallel-gpu-ode-solver_b200")
    for base in (pkgdir, os.path.join(ROOT, "include")):
        for dp, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dp, f)).read()
                    assert "oracle_py" not in text and "liboracle" not in text and "libref_" not in text, os.path.join(dp, f)


def test_workload_generators(workloads):
    wl = workloads.lorenz(NT=1 << 10)
    assert wl.buffers.ControlParameters[0, 0] == 0.0 and wl.buffers.ControlParameters[0, -1] == 21.0
    km = workloads.keller_miksis_t5(NT=128)
    assert km.shape.NCP == 13 and km.buffers.ControlParameters.shape == (13, 128) and np.isfinite(km.buffers.ControlParameters).all()
    sub = km.buffers.take([0, 5, 127])
    assert sub.NT == 3 and np.array_equal(sub.ControlParameters[:, 1], km.buffers.ControlParameters[:, 5])


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs on the host cores only (no CUDA) and prints one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in ([], ["--workload", "coupled"]):
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "64"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        line = json.loads(r.stdout.strip().splitlines()[-1])
        assert line["impl"] == "reference" and line["unit"] == "system-steps/s" and line["value"] > 0
        assert line["e2e"] == {"value": line["value"], "unit": "system-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
        assert "workload" in line["config"] and line["higher_is_better"] is True


def test_single_precision_workloads(workloads, pkg):
    """workloads.single_precision / the coupled generators' dtype argument: same sweep, float32 buffers, tolerances a float resolves."""
    wl = workloads.impact_t6(NT=256)
    w32 = workloads.single_precision(wl)
    assert w32.shape.dtype == np.float32 and w32.buffers.ActualState.dtype == np.float32 and w32.buffers.IntegerAccessories.dtype == np.int32
    assert np.array_equal(w32.buffers.ControlParameters, wl.buffers.ControlParameters.astype(np.float32))
    o = w32.options.resolved(w32.shape, w32.NT)
    assert o.RelativeTolerance.dtype == np.float32 and (o.RelativeTolerance == np.float32(1e-5)).all() and (o.EventTolerance == np.float32(1e-4)).all()
    assert o.MinimumTimeStep >= 1e-6 and wl.shape.dtype == np.float64                 # the original is untouched
    WC = pkg.workloads_coupled
    c32 = WC.bubble_ensemble(NS=3, UPS=16, dtype=np.float32)
    c64 = WC.bubble_ensemble(NS=3, UPS=16)
    assert c32.buffers.UnitParameters.dtype == np.float32 and c32.buffers.CouplingMatrix.dtype == np.float32
    assert np.array_equal(c32.buffers.UnitParameters, c64.buffers.UnitParameters.astype(np.float32))
    assert (c32.options.RelativeTolerance == 1e-5).all() and (c64.options.RelativeTolerance == np.array([1e-10, 1e-8])).all()


def test_option_enums_match_the_headers(pkg):
    """The Python mirrors' option numbers are the C enums' (both modules), including the options added by this implementation."""
    def enum(header, name):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        body = re.search(r"typedef enum " + name + r"\s*\{(.*?)\}", text, flags=re.S).group(1)
        return [t.split("=")[0].strip() for t in body.split(",") if t.strip()]
    per = enum("mpgos_b200.h", "mpgos_option")
    for i, n in enumerate(per):
        assert int(getattr(pkg.Options, n.replace("MPGOS_OPT_", ""))) == i, n
    cpl = enum("mpgos_b200_coupled.h", "mpgos_coupled_option")
    for i, n in enumerate(cpl):
        assert int(getattr(pkg.CoupledOptions, n.replace("MPGOS_COPT_", ""))) == i, n
    assert per[-1] == "MPGOS_OPT_DispatchOrder" and cpl[-1] == "MPGOS_COPT_DispatchOrder"
