"""ctypes binding of ``libmpgos_b200.so`` and a Python mirror of the reference's ``ProblemSolver``.

Method names, argument order and meaning follow the reference's host object
(``SourceCodes/SingleSystem_PerThread_Interface.cuh:180-207``): ``SolverOption``, ``SetHost``,
``GetHost``, ``SynchroniseFromHostToDevice``, ``SynchroniseFromDeviceToHost``, ``Solve``,
``InsertSynchronisationPoint``, ``SynchroniseSolver``, ``SynchroniseDevice``.  Errors that make the
reference print and ``exit()`` raise :class:`MpgosError` here.

Nothing in this module computes: every call forwards to the CUDA library.  If the library is
missing or no CUDA device is usable, the failure is loud (exception), never a fallback.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class MpgosError(RuntimeError):
    pass


class Algorithms(enum.IntEnum):          # Interface.cuh:33
    RK4 = 0
    RKCK45 = 1


class Variables(enum.IntEnum):           # Interface.cuh:35-47
    All = 0
    TimeDomain = 1
    ActualState = 2
    ActualTime = 3
    ControlParameters = 4
    SharedParameters = 5
    IntegerSharedParameters = 6
    Accessories = 7
    IntegerAccessories = 8
    DenseOutput = 9
    DenseIndex = 10
    DenseTime = 11
    DenseState = 12


class Options(enum.IntEnum):             # Interface.cuh:49-62 (+ two of ours)
    ThreadsPerBlock = 0
    InitialTimeStep = 1
    ActiveNumberOfThreads = 2
    MaximumTimeStep = 3
    MinimumTimeStep = 4
    TimeStepGrowLimit = 5
    TimeStepShrinkLimit = 6
    RelativeTolerance = 7
    AbsoluteTolerance = 8
    EventTolerance = 9
    EventDirection = 10
    DenseOutputMinimumTimeStep = 11
    DenseOutputSaveFrequency = 12
    PreferSharedMemory = 13
    LaneRefill = 14
    DeviceMessages = 15
    DispatchOrder = 16


_INDEXED = (Options.RelativeTolerance, Options.AbsoluteTolerance, Options.EventTolerance, Options.EventDirection)

# every symbol include/mpgos_b200.h declares
ABI_SYMBOLS = (
    "mpgos_last_error", "mpgos_abi_version", "mpgos_device_count",
    "mpgos_perthread_create", "mpgos_perthread_destroy", "mpgos_perthread_host_ptr",
    "mpgos_perthread_set_option", "mpgos_perthread_set_option_indexed", "mpgos_perthread_get_option",
    "mpgos_perthread_sync_h2d", "mpgos_perthread_sync_d2h", "mpgos_perthread_solve",
    "mpgos_perthread_insert_sync_point", "mpgos_perthread_sync_solver", "mpgos_perthread_sync_device",
    "mpgos_perthread_read_step_counters", "mpgos_perthread_reset_step_counters", "mpgos_perthread_total_steps",
    "mpgos_perthread_stream", "mpgos_perthread_device", "mpgos_perthread_device_ptr", "mpgos_perthread_solve_timed",
    "mpgos_perthread_builtin_count", "mpgos_perthread_builtin_name", "mpgos_perthread_builtin_desc",
    "mpgos_perthread_builtin_launcher", "mpgos_probe_fma_peak", "mpgos_perthread_save_state", "mpgos_perthread_restore_state",
    "mpgos_perthread_upload", "mpgos_perthread_download", "mpgos_probe_dmma_peak",
    "mpgos_perthread_create_sharded", "mpgos_perthread_shard_count", "mpgos_perthread_shard_info", "mpgos_perthread_shard_global_id",
    "mpgos_shard_instances", "mpgos_shard_global_id", "mpgos_fast_math_probe",
)


class _Desc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("NT", "SD", "NCP", "NSP", "NISP", "NE", "NA", "NIA", "NDO", "Algorithm",
                                        "PrecisionBytes", "Device", "InstanceOffset", "GlobalInstances",
                                        "ShardIndex", "ShardCount", "ShardTile")]


def library_path(strict: bool = False) -> str:
    """The product library, or (strict) the same sources built with -fmad=false for the parity tests."""
    override = os.environ.get("MPGOS_B200_LIBRARY")     # development: A/B builds of the same library
    if override and not strict:
        return override
    return os.path.join(HERE, "libmpgos_b200_strict.so" if strict else "libmpgos_b200.so")


def build_library(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", os.path.join(HERE, "csrc")], capture_output=True, text=True)
    if r.returncode != 0:
        raise MpgosError("building libmpgos_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout)
    return library_path()


_LIBS = {}


def load_library(strict: bool = False):
    """Load the CUDA library; raises if it has not been built (there is no other implementation)."""
    if strict in _LIBS:
        return _LIBS[strict]
    path = library_path(strict)
    if not os.path.exists(path):
        raise MpgosError(f"{path} is missing: run __graft_entry__.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.mpgos_last_error.restype = C.c_char_p
    lib.mpgos_perthread_create.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_void_p)]
    lib.mpgos_perthread_destroy.argtypes = [C.c_void_p]
    lib.mpgos_perthread_host_ptr.restype = C.c_void_p
    lib.mpgos_perthread_host_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    lib.mpgos_perthread_device_ptr.restype = C.c_void_p
    lib.mpgos_perthread_device_ptr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    lib.mpgos_perthread_set_option.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.mpgos_perthread_set_option_indexed.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    lib.mpgos_perthread_get_option.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.mpgos_perthread_sync_h2d.argtypes = [C.c_void_p, C.c_int]
    lib.mpgos_perthread_sync_d2h.argtypes = [C.c_void_p, C.c_int]
    lib.mpgos_perthread_solve.argtypes = [C.c_void_p, C.c_void_p]
    lib.mpgos_perthread_solve_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    for f in ("insert_sync_point", "sync_solver", "sync_device", "reset_step_counters", "save_state", "restore_state"):
        getattr(lib, "mpgos_perthread_" + f).argtypes = [C.c_void_p]
    lib.mpgos_perthread_read_step_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mpgos_perthread_total_steps.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.mpgos_perthread_stream.restype = C.c_void_p
    lib.mpgos_perthread_stream.argtypes = [C.c_void_p]
    lib.mpgos_perthread_device.argtypes = [C.c_void_p]
    lib.mpgos_perthread_builtin_name.restype = C.c_char_p
    lib.mpgos_perthread_builtin_desc.argtypes = [C.c_char_p, C.POINTER(_Desc)]
    lib.mpgos_perthread_builtin_launcher.restype = C.c_void_p
    lib.mpgos_perthread_builtin_launcher.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
    lib.mpgos_perthread_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.mpgos_perthread_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    lib.mpgos_probe_fma_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    if hasattr(lib, "mpgos_probe_dmma_peak"):      # absent from A/B builds of older revisions (MPGOS_B200_LIBRARY)
        lib.mpgos_probe_dmma_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    if hasattr(lib, "mpgos_perthread_create_sharded"):
        lib.mpgos_perthread_create_sharded.argtypes = [C.POINTER(_Desc), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.mpgos_perthread_shard_count.argtypes = [C.c_void_p]
        lib.mpgos_perthread_shard_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.mpgos_perthread_shard_global_id.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.mpgos_shard_instances.argtypes = [C.c_int] * 4
        lib.mpgos_shard_global_id.argtypes = [C.c_int] * 4
    _LIBS[strict] = lib
    return lib


def probe_fma_peak(device: int = 0, precision_bytes: int = 8) -> float:
    """Measured FMA throughput of the device in TFLOP/s (roofline denominator)."""
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.mpgos_probe_fma_peak(device, precision_bytes, C.byref(v)))
    return v.value


def fast_math_probe(function: int, x, y=None, device: int = 0):
    """(fast0, fast1, cuda0, cuda1): include/mpgos_b200/fast_math.cuh against the CUDA math library on the device; function 0 sincos,
    1 sincospi, 2 pow(x, y); the bundled subroutine pow_sincospi_sincos: 3 its pow(x, y), 4 its sincospi(x), 5 its sincos(y); 6 rcp2: 1/x and 1/y; 7: CUDA's sin(x), cos(x) against CUDA's sincos(x)."""
    lib = load_library()
    x = np.ascontiguousarray(x, np.float64)
    y = np.zeros_like(x) if y is None else np.ascontiguousarray(y, np.float64)
    out = [np.empty_like(x) for _ in range(4)]
    lib.mpgos_fast_math_probe.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    _check(lib, lib.mpgos_fast_math_probe(device, function, x.size, x.ctypes.data, y.ctypes.data, *[o.ctypes.data for o in out]))
    return tuple(out)


def probe_dmma_peak(device: int = 0) -> float:
    """Measured throughput of the FP64 tensor-core instruction (mma.sync.m8n8k4.f64) in TFLOP/s."""
    lib = load_library()
    v = C.c_double()
    _check(lib, lib.mpgos_probe_dmma_peak(device, C.byref(v)))
    return v.value


def _check(lib, rc: int):
    if rc != 0:
        raise MpgosError(f"mpgos_b200 error {rc}: {lib.mpgos_last_error().decode()}")


class ProblemSolver:
    """Python mirror of ``ProblemSolver<NT,SD,NCP,NSP,NISP,NE,NA,NIA,NDO,Algorithm,Precision>`` for a
    system compiled into the library (``system`` names it) -- template arguments become constructor
    arguments, everything else keeps the reference's names."""

    def __init__(self, system: str, NT: int, NDO: int = 0, Algorithm: int = Algorithms.RKCK45, Precision=np.float64,
                 Device: int = 0, InstanceOffset: int = 0, GlobalInstances: int = 0, strict: bool = False,
                 Devices=None, Tile: int = 0):
        """``Devices`` (a list of device ordinals): ONE solver object over several devices of the node -- the sweep is dealt to them
        in tiles of ``Tile`` instances, everything else (host mirrors in the sweep's layout, SetHost/GetHost, synchronisation,
        Solve) stays as for one device (mpgos_perthread_create_sharded)."""
        self._lib = lib = load_library(strict)
        self._h = None
        self.system = system
        self.dtype = np.dtype(Precision)
        d = _Desc()
        _check(lib, lib.mpgos_perthread_builtin_desc(system.encode(), C.byref(d)))
        d.NT, d.NDO, d.Algorithm, d.PrecisionBytes, d.Device = NT, NDO, int(Algorithm), self.dtype.itemsize, Device
        d.InstanceOffset, d.GlobalInstances = InstanceOffset, GlobalInstances
        self.desc = d
        self._launcher = lib.mpgos_perthread_builtin_launcher(system.encode(), int(Algorithm), self.dtype.itemsize, int(NDO > 0))
        if not self._launcher:
            raise MpgosError(lib.mpgos_last_error().decode())
        h = C.c_void_p()
        if Devices is not None:
            devs = (C.c_int * len(Devices))(*[int(x) for x in Devices])
            _check(lib, lib.mpgos_perthread_create_sharded(C.byref(d), devs, len(Devices), int(Tile), C.byref(h)))
        else:
            _check(lib, lib.mpgos_perthread_create(C.byref(d), C.byref(h)))
        self._h = h
        self._views: Dict[int, np.ndarray] = {}

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if self._h is not None:
            self._lib.mpgos_perthread_destroy(self._h)
            self._h = None
            self._views = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- host mirrors -----------------------------------------------------------------------------
    def host(self, variable: int) -> np.ndarray:
        """numpy view of the pinned host mirror, shaped [serial, instance] (flat index = instance + serial*NT)."""
        variable = int(variable)
        v = self._views.get(variable)
        if v is not None:
            return v
        n = C.c_size_t()
        p = self._lib.mpgos_perthread_host_ptr(self._h, variable, C.byref(n))
        ints = variable in (Variables.IntegerSharedParameters, Variables.IntegerAccessories, Variables.DenseIndex)
        dt = np.dtype(np.int32) if ints else self.dtype
        if n.value == 0:
            arr = np.zeros(0, dt)
        else:
            if not p:
                raise MpgosError(self._lib.mpgos_last_error().decode())
            buf = (C.c_char * (n.value * dt.itemsize)).from_address(p)
            arr = np.frombuffer(buf, dtype=dt)
        d = self.desc
        NT = d.NT
        shape = {Variables.TimeDomain: (2, NT), Variables.ActualState: (d.SD, NT), Variables.ActualTime: (NT,),
                 Variables.ControlParameters: (d.NCP, NT), Variables.SharedParameters: (d.NSP,),
                 Variables.IntegerSharedParameters: (d.NISP,), Variables.Accessories: (d.NA, NT),
                 Variables.IntegerAccessories: (d.NIA, NT), Variables.DenseIndex: (NT,),
                 Variables.DenseTime: (d.NDO, NT), Variables.DenseState: (d.NDO, d.SD, NT)}[Variables(variable)]
        arr = arr.reshape(shape)
        self._views[variable] = arr
        return arr

    def _bound(self, what: str, value: int, hi: int):
        if value < 0 or value > hi:
            raise MpgosError(f"{what} index {value} outside 0-{hi}")

    def SetHost(self, *args):
        """SetHost(problem, var, serial, value) | (problem, var, value) | (var, serial, value) |
        (problem, var, serial, time_step, value) -- Interface.cuh:708-876."""
        d = self.desc
        if len(args) == 5:
            problem, var, comp, step, value = args
            if Variables(var) != Variables.DenseState:
                raise MpgosError("SetHost with 5 arguments is only applicable to DenseState")
            self._bound("ProblemNumber", problem, d.NT - 1); self._bound("DenseState", comp, d.SD - 1); self._bound("DenseState/TimeStep", step, d.NDO - 1)
            self.host(var)[step, comp, problem] = value
        elif len(args) == 4:
            problem, var, serial, value = args
            self._bound("ProblemNumber", problem, d.NT - 1)
            a = self.host(var)
            if a.ndim != 2:
                raise MpgosError(f"SetHost with 4 arguments is not applicable to {Variables(var).name}")
            self._bound(Variables(var).name, serial, a.shape[0] - 1)
            a[serial, problem] = value
        elif len(args) == 3 and Variables(args[0]) in (Variables.SharedParameters, Variables.IntegerSharedParameters):
            var, serial, value = args
            a = self.host(var)
            self._bound(Variables(var).name, serial, a.shape[0] - 1)
            a[serial] = value
        elif len(args) == 3:
            problem, var, value = args
            self._bound("ProblemNumber", problem, d.NT - 1)
            a = self.host(var)
            if a.ndim != 1 or Variables(var) not in (Variables.ActualTime, Variables.DenseIndex):
                raise MpgosError(f"SetHost with 3 arguments is not applicable to {Variables(var).name}")
            if Variables(var) == Variables.DenseIndex:
                self._bound("DenseIndex-Value", int(value), d.NDO - 1)      # value check as Interface.cuh:770
            a[problem] = value
        else:
            raise TypeError("SetHost takes 3, 4 or 5 arguments")

    def GetHost(self, *args):
        """GetHost(problem, var, serial) | (problem, var) | (var, serial) | (problem, var, serial, time_step)."""
        d = self.desc
        if len(args) == 4:
            problem, var, comp, step = args
            self._bound("ProblemNumber", problem, d.NT - 1); self._bound("DenseState", comp, d.SD - 1); self._bound("DenseState/TimeStep", step, d.NDO - 1)
            return self.host(var)[step, comp, problem]
        if len(args) == 3:
            problem, var, serial = args
            self._bound("ProblemNumber", problem, d.NT - 1)
            a = self.host(var)
            self._bound(Variables(var).name, serial, a.shape[0] - 1)
            return a[serial, problem]
        if len(args) == 2 and Variables(args[0]) in (Variables.SharedParameters, Variables.IntegerSharedParameters):
            var, serial = args
            a = self.host(var)
            self._bound(Variables(var).name, serial, a.shape[0] - 1)
            return a[serial]
        if len(args) == 2:
            problem, var = args
            self._bound("ProblemNumber", problem, d.NT - 1)
            return self.host(var)[problem]
        raise TypeError("GetHost takes 2, 3 or 4 arguments")

    # -- options ------------------------------------------------------------------------------------
    def SolverOption(self, option: int, *args):
        """SolverOption(option, value) | SolverOption(option, index, value) -- Interface.cuh:613-703."""
        if len(args) == 1:
            _check(self._lib, self._lib.mpgos_perthread_set_option(self._h, int(option), float(args[0])))
        elif len(args) == 2:
            _check(self._lib, self._lib.mpgos_perthread_set_option_indexed(self._h, int(option), int(args[0]), float(args[1])))
        else:
            raise TypeError("SolverOption takes (option, value) or (option, index, value)")

    def GetOption(self, option: int, index: int = 0) -> float:
        v = C.c_double()
        _check(self._lib, self._lib.mpgos_perthread_get_option(self._h, int(option), index, C.byref(v)))
        return v.value

    # -- copies, solve, synchronisation (all asynchronous on the solver's stream) ------------------
    def SynchroniseFromHostToDevice(self, variable: int = Variables.All):
        _check(self._lib, self._lib.mpgos_perthread_sync_h2d(self._h, int(variable)))

    def SynchroniseFromDeviceToHost(self, variable: int = Variables.All):
        _check(self._lib, self._lib.mpgos_perthread_sync_d2h(self._h, int(variable)))

    def Solve(self):
        _check(self._lib, self._lib.mpgos_perthread_solve(self._h, self._launcher))

    def InsertSynchronisationPoint(self):
        _check(self._lib, self._lib.mpgos_perthread_insert_sync_point(self._h))

    def SynchroniseSolver(self):
        _check(self._lib, self._lib.mpgos_perthread_sync_solver(self._h))

    def SynchroniseDevice(self):
        _check(self._lib, self._lib.mpgos_perthread_sync_device(self._h))

    # -- additions: step counters, timed solve -------------------------------------------------------
    def StepCounters(self) -> Tuple[np.ndarray, np.ndarray]:
        att = np.zeros(self.desc.NT, np.uint64)
        acc = np.zeros(self.desc.NT, np.uint64)
        _check(self._lib, self._lib.mpgos_perthread_read_step_counters(self._h, att.ctypes.data, acc.ctypes.data))
        return att, acc

    def ResetStepCounters(self):
        _check(self._lib, self._lib.mpgos_perthread_reset_step_counters(self._h))

    def TotalSteps(self) -> Tuple[int, int]:
        a, b = C.c_ulonglong(), C.c_ulonglong()
        _check(self._lib, self._lib.mpgos_perthread_total_steps(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def Upload(self, variable: int, host_ptr: int, nbytes: int):
        """Asynchronous H2D copy of one variable from a caller-owned host array (address, bytes)."""
        _check(self._lib, self._lib.mpgos_perthread_upload(self._h, int(variable), host_ptr, nbytes))

    def Download(self, variable: int, host_ptr: int, nbytes: int):
        _check(self._lib, self._lib.mpgos_perthread_download(self._h, int(variable), host_ptr, nbytes))

    def SaveState(self):
        """Device-resident checkpoint of all per-instance buffers (asynchronous)."""
        _check(self._lib, self._lib.mpgos_perthread_save_state(self._h))

    def RestoreState(self):
        _check(self._lib, self._lib.mpgos_perthread_restore_state(self._h))

    @property
    def stream(self) -> int:
        """The solver's cudaStream_t as an integer (for torch.cuda.ExternalStream)."""
        return int(self._lib.mpgos_perthread_stream(self._h) or 0)

    def SolveTimed(self, phases: int = 1) -> float:
        """Enqueue ``phases`` Solve() calls bracketed by CUDA events on the solver's stream; returns milliseconds."""
        ms = C.c_float()
        _check(self._lib, self._lib.mpgos_perthread_solve_timed(self._h, self._launcher, phases, C.byref(ms)))
        return ms.value


def apply_workload(ps: ProblemSolver, wl, buffers=None):
    """Fill a solver object from a :class:`workloads.Workload` (options + host mirrors)."""
    b = wl.buffers if buffers is None else buffers
    o = wl.options.resolved(wl.shape, b.NT)
    for name, var in (("TimeDomain", Variables.TimeDomain), ("ActualState", Variables.ActualState),
                      ("ActualTime", Variables.ActualTime), ("ControlParameters", Variables.ControlParameters),
                      ("SharedParameters", Variables.SharedParameters), ("IntegerSharedParameters", Variables.IntegerSharedParameters),
                      ("Accessories", Variables.Accessories), ("IntegerAccessories", Variables.IntegerAccessories),
                      ("DenseIndex", Variables.DenseIndex), ("DenseTime", Variables.DenseTime), ("DenseState", Variables.DenseState)):
        src = getattr(b, name)
        if src.size:
            ps.host(var)[...] = src
    for opt in ("InitialTimeStep", "MaximumTimeStep", "MinimumTimeStep", "TimeStepGrowLimit", "TimeStepShrinkLimit",
                "DenseOutputMinimumTimeStep", "DenseOutputSaveFrequency", "ActiveNumberOfThreads"):
        ps.SolverOption(Options[opt], getattr(o, opt))
    if o.ThreadsPerBlock:
        ps.SolverOption(Options.ThreadsPerBlock, o.ThreadsPerBlock)
    for i in range(wl.shape.SD):
        ps.SolverOption(Options.RelativeTolerance, i, o.RelativeTolerance[i])
        ps.SolverOption(Options.AbsoluteTolerance, i, o.AbsoluteTolerance[i])
    for i in range(wl.shape.NE):
        ps.SolverOption(Options.EventTolerance, i, o.EventTolerance[i])
        ps.SolverOption(Options.EventDirection, i, int(o.EventDirection[i]))


def read_back(ps: ProblemSolver, wl):
    """Copy every mirror back and return them as a :class:`workloads.Buffers`."""
    from .workloads import Buffers
    ps.SynchroniseFromDeviceToHost(Variables.All)
    ps.InsertSynchronisationPoint()
    ps.SynchroniseSolver()
    out = Buffers(wl.shape, ps.desc.NT)
    for name, var in (("TimeDomain", Variables.TimeDomain), ("ActualState", Variables.ActualState),
                      ("ActualTime", Variables.ActualTime), ("ControlParameters", Variables.ControlParameters),
                      ("SharedParameters", Variables.SharedParameters), ("IntegerSharedParameters", Variables.IntegerSharedParameters),
                      ("Accessories", Variables.Accessories), ("IntegerAccessories", Variables.IntegerAccessories),
                      ("DenseIndex", Variables.DenseIndex), ("DenseTime", Variables.DenseTime), ("DenseState", Variables.DenseState)):
        src = ps.host(var)
        if src.size:
            getattr(out, name)[...] = src
    return out


def run_workload(wl, buffers=None, phases: Optional[int] = None, lane_refill: Optional[bool] = None, device: int = 0,
                 threads_per_block: Optional[int] = None, strict: bool = False, devices=None, tile: int = 0):
    """Integrate a workload on the GPU through the public API; returns (buffers_after, attempted, accepted)
    with the same meaning as the checker's ``run_oracle`` / ``run_ref``."""
    b = wl.buffers if buffers is None else buffers
    sh = wl.shape
    with ProblemSolver(wl.name, b.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype, Device=device, strict=strict,
                       Devices=devices, Tile=tile) as ps:
        apply_workload(ps, wl, b)
        if lane_refill is not None:
            ps.SolverOption(Options.LaneRefill, int(lane_refill))
        if threads_per_block is not None:
            ps.SolverOption(Options.ThreadsPerBlock, threads_per_block)
        ps.SynchroniseFromHostToDevice(Variables.All)
        for _ in range(wl.phases if phases is None else phases):
            ps.Solve()
        out = read_back(ps, wl)
        att, acc = ps.StepCounters()
    return out, att.astype(np.int64), acc.astype(np.int64)
