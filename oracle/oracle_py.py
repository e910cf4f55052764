"""TEST INFRASTRUCTURE ONLY -- ctypes loader for the checker libraries.

* ``run_ref``    : oracle/_ref/libref_<system>_<algo>_<prec>.so -- the reference's own kernel source
                   compiled for the host (built by oracle/Makefile from /root/reference).
* ``run_oracle`` : oracle/liboracle_perthread.so -- our CPU restatement of the same algorithm.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this
module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class _Options(C.Structure):
    _fields_ = [("InitialTimeStep", C.c_double), ("MaximumTimeStep", C.c_double), ("MinimumTimeStep", C.c_double),
                ("TimeStepGrowLimit", C.c_double), ("TimeStepShrinkLimit", C.c_double),
                ("DenseOutputMinimumTimeStep", C.c_double), ("DenseOutputSaveFrequency", C.c_int),
                ("NumberOfActiveThreads", C.c_int), ("Phases", C.c_int), ("HostThreads", C.c_int)]


class _Buffers(C.Structure):
    _fields_ = [("NT", C.c_int)] + [(n, C.c_void_p) for n in (
        "TimeDomain", "ActualState", "ActualTime", "ControlParameters", "SharedParameters",
        "IntegerSharedParameters", "Accessories", "IntegerAccessories", "RelativeTolerance",
        "AbsoluteTolerance", "EventTolerance", "EventDirection", "DenseOutputIndex",
        "DenseOutputTimeInstances", "DenseOutputStates", "steps_attempted", "steps_accepted")]


class _Shape(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("SD", "NCP", "NSP", "NISP", "NE", "NA", "NIA", "NDO", "Algorithm", "Precision")]


def _ptr(a: Optional[np.ndarray]):
    if a is None or a.size == 0:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _pack(wl, buffers, phases, host_threads):
    """Build the C structs for a run that mutates ``buffers`` in place."""
    shape, NT = wl.shape, buffers.NT
    o = wl.options.resolved(shape, NT)
    att = np.zeros(NT, np.int64)
    acc = np.zeros(NT, np.int64)
    keep = [o.RelativeTolerance, o.AbsoluteTolerance, o.EventTolerance, o.EventDirection, att, acc]
    cb = _Buffers(NT, _ptr(buffers.TimeDomain), _ptr(buffers.ActualState), _ptr(buffers.ActualTime),
                  _ptr(buffers.ControlParameters), _ptr(buffers.SharedParameters),
                  _ptr(buffers.IntegerSharedParameters), _ptr(buffers.Accessories),
                  _ptr(buffers.IntegerAccessories), _ptr(o.RelativeTolerance), _ptr(o.AbsoluteTolerance),
                  _ptr(o.EventTolerance), _ptr(o.EventDirection), _ptr(buffers.DenseIndex),
                  _ptr(buffers.DenseTime), _ptr(buffers.DenseState), _ptr(att), _ptr(acc))
    co = _Options(o.InitialTimeStep, o.MaximumTimeStep, o.MinimumTimeStep, o.TimeStepGrowLimit,
                  o.TimeStepShrinkLimit, o.DenseOutputMinimumTimeStep, o.DenseOutputSaveFrequency,
                  o.ActiveNumberOfThreads, phases, host_threads)
    return cb, co, att, acc, keep


def _prec_tag(dtype) -> str:
    return "f64" if np.dtype(dtype) == np.float64 else "f32"


def ref_library_path(wl, fast: bool = False) -> str:
    from_algo = {0: "rk4", 1: "rkck45"}[wl.shape.Algorithm]
    name = f"libref_{wl.name}_{from_algo}_{_prec_tag(wl.shape.dtype)}{'_fast' if fast else ''}.so"
    return os.path.join(HERE, "_ref", name)


def ref_available(wl, fast: bool = False) -> bool:
    return os.path.exists(ref_library_path(wl, fast))


def _check_shape(s: _Shape, wl):
    sh = wl.shape
    got = (s.SD, s.NCP, s.NSP, s.NISP, s.NE, s.NA, s.NIA, s.NDO, s.Algorithm, s.Precision)
    want = (sh.SD, sh.NCP, sh.NSP, sh.NISP, sh.NE, sh.NA, sh.NIA, sh.NDO, sh.Algorithm, np.dtype(sh.dtype).itemsize)
    if got != want:
        raise ValueError(f"checker library shape {got} does not match workload shape {want}")


def run_ref(wl, buffers=None, phases: Optional[int] = None, host_threads: int = 1, fast: bool = False
            ) -> Tuple[object, np.ndarray, np.ndarray]:
    """Run the host-compiled REFERENCE kernel; returns (buffers_after, attempted, accepted)."""
    lib = C.CDLL(ref_library_path(wl, fast))
    s = _Shape()
    lib.ref_perthread_shape(C.byref(s))
    _check_shape(s, wl)
    out = (buffers if buffers is not None else wl.buffers).copy()
    cb, co, att, acc, _keep = _pack(wl, out, wl.phases if phases is None else phases, host_threads)
    lib.ref_perthread_run.restype = C.c_int
    rc = lib.ref_perthread_run(C.byref(cb), C.byref(co))
    if rc != 0:
        raise RuntimeError(f"ref_perthread_run failed: {rc}")
    return out, att, acc


_ORACLE = {}


def oracle_library_path(fast: bool = False) -> str:
    return os.path.join(HERE, "liboracle_perthread_fast.so" if fast else "liboracle_perthread.so")


def run_oracle(wl, buffers=None, phases: Optional[int] = None, host_threads: int = 1, fast: bool = False,
               double_literals: bool = False) -> Tuple[object, np.ndarray, np.ndarray]:
    """Run our CPU restatement; returns (buffers_after, attempted, accepted).

    ``double_literals`` only matters for float32 runs: True evaluates the system definitions'
    floating literals in double exactly as the reference's FP32 instantiation does (bit-comparable
    with oracle/_ref); False is the true-FP32 arithmetic of the CUDA built-in systems."""
    path = oracle_library_path(fast)
    lib = _ORACLE.get(path)
    if lib is None:
        lib = _ORACLE[path] = C.CDLL(path)
        lib.oracle_perthread_run.restype = C.c_int
        lib.oracle_perthread_query_shape.restype = C.c_int
    sh = wl.shape
    prec = np.dtype(sh.dtype).itemsize
    if prec == 4 and double_literals:
        prec = -4
    s = _Shape()
    rc = lib.oracle_perthread_query_shape(wl.name.encode(), sh.Algorithm, prec, sh.NDO, C.byref(s))
    if rc != 0:
        raise KeyError(f"oracle has no system {wl.name!r} for algorithm {sh.Algorithm} precision {prec}")
    _check_shape(s, wl)
    out = (buffers if buffers is not None else wl.buffers).copy()
    cb, co, att, acc, _keep = _pack(wl, out, wl.phases if phases is None else phases, host_threads)
    rc = lib.oracle_perthread_run(wl.name.encode(), sh.Algorithm, prec, sh.NDO, C.byref(cb), C.byref(co))
    if rc != 0:
        raise RuntimeError(f"oracle_perthread_run failed: {rc}")
    return out, att, acc


# ---- the reference compiled by nvcc for sm_100a (oracle/_ref/libref_gpu_*.so) -------------------------

def ref_gpu_library_path(wl, variant: str = "") -> str:
    """variant "": compiled without a register cap; "stock": with the -maxrregcount of the reference's own tutorial makefile."""
    algo = {0: "rk4", 1: "rkck45"}[wl.shape.Algorithm]
    return os.path.join(HERE, "_ref", f"libref_gpu_{wl.name}_{algo}_{_prec_tag(wl.shape.dtype)}{'_' + variant if variant else ''}.so")


def ref_gpu_available(wl, variant: str = "") -> bool:
    return os.path.exists(ref_gpu_library_path(wl, variant))


_REFGPU = {}


def run_ref_gpu(wl, buffers=None, phases: Optional[int] = None, threads_per_block: int = 64, repeats: int = 1, variant: str = ""):
    """Run the REFERENCE MPGOS kernel on the GPU; returns (buffers_after, kernel_ms) -- kernel_ms is the best
    device time of the Solve() calls of one run over `repeats` runs (CUDA events, copies excluded)."""
    path = ref_gpu_library_path(wl, variant)
    lib = _REFGPU.get(path)
    if lib is None:
        lib = _REFGPU[path] = C.CDLL(path)
    s = _Shape()
    lib.ref_perthread_shape(C.byref(s))
    _check_shape(s, wl)
    out = (buffers if buffers is not None else wl.buffers).copy()
    if out.NT > lib.ref_perthread_gpu_capacity():
        raise ValueError("sweep larger than the capacity the reference library was compiled for")
    cb, co, att, acc, _keep = _pack(wl, out, wl.phases if phases is None else phases, 1)
    ms = C.c_float()
    rc = lib.ref_perthread_gpu_run(C.byref(cb), C.byref(co), threads_per_block, repeats, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"ref_perthread_gpu_run failed: {rc}")
    return out, ms.value


# ---- coupled (block-per-system) module ---------------------------------------------------------------------------

class _CShape(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("UPS", "UD", "NC", "CBW", "CCI", "NUP", "NSP", "NGP", "NiGP", "NUA", "NiUA", "NSA", "NiSA", "NE", "NDO",
                                        "Algorithm", "Precision")]


class _COptions(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("InitialTimeStep", "MaximumTimeStep", "MinimumTimeStep", "TimeStepGrowLimit", "TimeStepShrinkLimit",
                                           "DenseOutputMinimumTimeStep")] + [(n, C.c_int) for n in ("DenseOutputSaveFrequency", "ActiveSystems", "Phases", "HostThreads")]


class _CBuffers(C.Structure):
    _fields_ = [("NS", C.c_int)] + [(n, C.c_void_p) for n in (
        "TimeDomain", "ActualState", "ActualTime", "UnitParameters", "SystemParameters", "GlobalParameters", "IntegerGlobalParameters",
        "UnitAccessories", "IntegerUnitAccessories", "SystemAccessories", "IntegerSystemAccessories", "CouplingMatrix", "CouplingStrength",
        "CouplingIndex", "RelativeTolerance", "AbsoluteTolerance", "EventTolerance", "EventDirection", "DenseOutputIndex",
        "DenseOutputTimeInstances", "DenseOutputStates", "steps_attempted", "steps_accepted")]


def _cpack(wl, b, phases, host_threads):
    o = wl.options.resolved(wl.shape, b.NS)
    dt = np.dtype(getattr(wl.shape, "dtype", np.float64))
    o.RelativeTolerance, o.AbsoluteTolerance, o.EventTolerance = (np.ascontiguousarray(a, dt) for a in (o.RelativeTolerance, o.AbsoluteTolerance, o.EventTolerance))
    att, acc = np.zeros(b.NS, np.int64), np.zeros(b.NS, np.int64)
    keep = [o.RelativeTolerance, o.AbsoluteTolerance, o.EventTolerance, o.EventDirection]
    cb = _CBuffers(b.NS, _ptr(b.TimeDomain), _ptr(b.ActualState), _ptr(b.ActualTime), _ptr(b.UnitParameters), _ptr(b.SystemParameters),
                   _ptr(b.GlobalParameters), _ptr(b.IntegerGlobalParameters), _ptr(b.UnitAccessories), _ptr(b.IntegerUnitAccessories),
                   _ptr(b.SystemAccessories), _ptr(b.IntegerSystemAccessories), _ptr(b.CouplingMatrix), _ptr(b.CouplingStrength),
                   _ptr(b.CouplingIndex), _ptr(o.RelativeTolerance), _ptr(o.AbsoluteTolerance), _ptr(o.EventTolerance), _ptr(o.EventDirection),
                   _ptr(b.DenseIndex), _ptr(b.DenseTime), _ptr(b.DenseState), _ptr(att), _ptr(acc))
    co = _COptions(o.InitialTimeStep, o.MaximumTimeStep, o.MinimumTimeStep, o.TimeStepGrowLimit, o.TimeStepShrinkLimit,
                   o.DenseOutputMinimumTimeStep, o.DenseOutputSaveFrequency, o.ActiveSystems, phases, host_threads)
    return cb, co, att, acc, keep


def _ccheck(s, wl, check_precision=True):
    sh = wl.shape
    got = (s.UPS, s.UD, s.NC, s.CBW, s.CCI, s.NUP, s.NSP, s.NGP, s.NiGP, s.NUA, s.NiUA, s.NSA, s.NiSA, s.NE, s.NDO, s.Algorithm)
    want = (sh.UPS, sh.UD, sh.NC, sh.CBW, sh.CCI, sh.NUP, sh.NSP, sh.NGP, sh.NiGP, sh.NUA, sh.NiUA, sh.NSA, sh.NiSA, sh.NE, sh.NDO, sh.Algorithm)
    if check_precision and s.Precision != np.dtype(getattr(sh, "dtype", np.float64)).itemsize:
        raise ValueError(f"coupled checker library precision {s.Precision} bytes does not match the workload's")
    if got != want:
        raise ValueError(f"coupled checker library shape {got} does not match workload shape {want}")


_CLIBS = {}


def run_coupled_oracle(wl, buffers=None, phases=None, host_threads: int = 1, fast: bool = False):
    """Our CPU restatement of the block-per-system phase; returns (buffers_after, attempted, accepted) per system."""
    path = os.path.join(HERE, "liboracle_coupled_fast.so" if fast else "liboracle_coupled.so")
    lib = _CLIBS.get(path)
    if lib is None:
        lib = _CLIBS[path] = C.CDLL(path)
    sh = wl.shape
    s = _CShape()
    rc = lib.oracle_coupled_query_shape(wl.name.encode(), sh.UPS, sh.CBW, sh.CCI, sh.Algorithm, sh.NDO, C.byref(s))
    if rc != 0:
        raise KeyError(f"coupled oracle has no system {wl.name!r} with UPS={sh.UPS} CBW={sh.CBW} CCI={sh.CCI}")
    _ccheck(s, wl, check_precision=False)         # the shape query answers for double; the run call states the precision
    out = (buffers if buffers is not None else wl.buffers).copy()
    cb, co, att, acc, _keep = _cpack(wl, out, wl.phases if phases is None else phases, host_threads)
    rc = lib.oracle_coupled_run_p(wl.name.encode(), sh.UPS, sh.CBW, sh.CCI, sh.Algorithm, sh.NDO, np.dtype(getattr(sh, "dtype", np.float64)).itemsize,
                                  C.byref(cb), C.byref(co))
    if rc != 0:
        raise RuntimeError(f"oracle_coupled_run failed: {rc}")
    return out, att, acc


def coupled_ref_library_path(tag: str, gpu: bool = False) -> str:
    return os.path.join(HERE, "_ref", f"libref_{'gpu_' if gpu else ''}coupled_{tag}.so")


def run_coupled_ref(tag: str, wl, buffers=None, phases=None):
    """The reference's own coupled kernel source on the host (one std::thread per CUDA thread); `tag` names the build
    (oracle/Makefile REFCOUPLED list).  Returns buffers_after."""
    lib = C.CDLL(coupled_ref_library_path(tag))
    s = _CShape()
    lib.ref_coupled_shape(C.byref(s))
    _ccheck(s, wl)
    out = (buffers if buffers is not None else wl.buffers).copy()
    if out.NS > lib.ref_coupled_capacity():
        raise ValueError("more systems than the reference library was compiled for")
    cb, co, att, acc, _keep = _cpack(wl, out, wl.phases if phases is None else phases, 1)
    rc = lib.ref_coupled_run(C.byref(cb), C.byref(co))
    if rc != 0:
        raise RuntimeError(f"ref_coupled_run failed: {rc}")
    return out


def run_coupled_ref_gpu(tag: str, wl, buffers=None, phases=None, shared_matrices: int = 0, repeats: int = 1):
    """The REFERENCE MPGOS coupled module on the GPU (oracle/_ref/libref_gpu_coupled_<tag>.so); returns (buffers_after, kernel_ms)."""
    path = coupled_ref_library_path(tag, gpu=True)
    lib = _REFGPU.get(path)
    if lib is None:
        lib = _REFGPU[path] = C.CDLL(path)
    s = _CShape()
    lib.ref_coupled_shape(C.byref(s))
    _ccheck(s, wl)
    out = (buffers if buffers is not None else wl.buffers).copy()
    if out.NS > lib.ref_coupled_capacity():
        raise ValueError("more systems than the reference library was compiled for")
    cb, co, att, acc, _keep = _cpack(wl, out, wl.phases if phases is None else phases, 1)
    ms = C.c_float()
    rc = lib.ref_coupled_gpu_run(C.byref(cb), C.byref(co), shared_matrices, repeats, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"ref_coupled_gpu_run failed: {rc}")
    return out, ms.value
