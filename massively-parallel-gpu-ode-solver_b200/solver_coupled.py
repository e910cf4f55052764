"""ctypes binding of the coupled (block-per-system) part of ``libmpgos_b200.so`` and a Python mirror of the
reference's coupled ``ProblemSolver`` (``SourceCodes/CoupledSystems_PerBlock_Interface.cuh:212-247``) for the systems
compiled into the library.  Forwards every call to the CUDA library; no fallback."""
from __future__ import annotations

import ctypes as C
import enum
from typing import Dict, Optional

import numpy as np

from .solver import MpgosError, load_library, _check


class CoupledVariables(enum.IntEnum):     # CoupledSystems_PerBlock_Interface.cuh:35-53
    All = 0
    TimeDomain = 1
    ActualTime = 2
    ActualState = 3
    UnitParameters = 4
    SystemParameters = 5
    GlobalParameters = 6
    IntegerGlobalParameters = 7
    UnitAccessories = 8
    IntegerUnitAccessories = 9
    SystemAccessories = 10
    IntegerSystemAccessories = 11
    CouplingMatrix = 12
    CouplingStrength = 13
    CouplingIndex = 14
    DenseOutput = 15
    DenseIndex = 16
    DenseTime = 17
    DenseState = 18


class CoupledOptions(enum.IntEnum):       # CoupledSystems_PerBlock_Interface.cuh:55-68 (+ DeviceMessages)
    InitialTimeStep = 0
    ActiveSystems = 1
    MaximumTimeStep = 2
    MinimumTimeStep = 3
    TimeStepGrowLimit = 4
    TimeStepShrinkLimit = 5
    RelativeTolerance = 6
    AbsoluteTolerance = 7
    EventTolerance = 8
    EventDirection = 9
    DenseOutputSaveFrequency = 10
    DenseOutputMinimumTimeStep = 11
    SharedGlobalVariables = 12
    SharedCouplingMatrices = 13
    DeviceMessages = 14
    TensorCoreCoupling = 15
    DispatchOrder = 16


COUPLED_ABI_SYMBOLS = (
    "mpgos_coupled_create", "mpgos_coupled_destroy", "mpgos_coupled_host_ptr", "mpgos_coupled_set_matrix_entry", "mpgos_coupled_get_matrix_entry",
    "mpgos_coupled_set_option", "mpgos_coupled_set_option_indexed", "mpgos_coupled_sync_h2d", "mpgos_coupled_sync_d2h", "mpgos_coupled_solve",
    "mpgos_coupled_solve_timed", "mpgos_coupled_insert_sync_point", "mpgos_coupled_sync_solver", "mpgos_coupled_sync_device",
    "mpgos_coupled_read_step_counters", "mpgos_coupled_reset_step_counters", "mpgos_coupled_save_state", "mpgos_coupled_restore_state",
    "mpgos_coupled_stream", "mpgos_coupled_builtin_desc", "mpgos_coupled_builtin_launcher", "mpgos_coupled_builtin_launcher_p",
    "mpgos_coupled_upload", "mpgos_coupled_download", "mpgos_coupled_create_sharded", "mpgos_coupled_shard_count",
)


class _CDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("NS", "UPS", "UD", "TPB", "SPB", "NC", "CBW", "CCI", "NUP", "NSP", "NGP", "NiGP", "NUA", "NiUA", "NSA", "NiSA",
                                        "NE", "NDO", "Algorithm", "PrecisionBytes", "Device", "SystemOffset",
                                        "ShardIndex", "ShardCount", "ShardTile")]


_BOUND = False


def _bind(lib):
    global _BOUND
    if _BOUND:
        return
    v = C.c_void_p
    lib.mpgos_coupled_create.argtypes = [C.POINTER(_CDesc), C.POINTER(v)]
    lib.mpgos_coupled_create_sharded.argtypes = [C.POINTER(_CDesc), C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(v)]
    lib.mpgos_coupled_shard_count.argtypes = [v]
    lib.mpgos_coupled_destroy.argtypes = [v]
    lib.mpgos_coupled_host_ptr.restype = v
    lib.mpgos_coupled_host_ptr.argtypes = [v, C.c_int, C.POINTER(C.c_size_t)]
    lib.mpgos_coupled_set_matrix_entry.argtypes = [v, C.c_int, C.c_int, C.c_int, C.c_double]
    lib.mpgos_coupled_get_matrix_entry.argtypes = [v, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
    lib.mpgos_coupled_set_option.argtypes = [v, C.c_int, C.c_double]
    lib.mpgos_coupled_set_option_indexed.argtypes = [v, C.c_int, C.c_int, C.c_double]
    for f in ("sync_h2d", "sync_d2h"):
        getattr(lib, "mpgos_coupled_" + f).argtypes = [v, C.c_int]
    lib.mpgos_coupled_solve.argtypes = [v, v]
    lib.mpgos_coupled_solve_timed.argtypes = [v, v, C.c_int, C.POINTER(C.c_float)]
    for f in ("insert_sync_point", "sync_solver", "sync_device", "reset_step_counters", "save_state", "restore_state"):
        getattr(lib, "mpgos_coupled_" + f).argtypes = [v]
    lib.mpgos_coupled_read_step_counters.argtypes = [v, v, v]
    lib.mpgos_coupled_upload.argtypes = [v, C.c_int, v, C.c_size_t]
    lib.mpgos_coupled_download.argtypes = [v, C.c_int, v, C.c_size_t]
    lib.mpgos_coupled_stream.restype = v
    lib.mpgos_coupled_stream.argtypes = [v]
    lib.mpgos_coupled_builtin_desc.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_CDesc)]
    lib.mpgos_coupled_builtin_launcher.restype = v
    lib.mpgos_coupled_builtin_launcher.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.mpgos_coupled_builtin_launcher_p.restype = v
    lib.mpgos_coupled_builtin_launcher_p.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    _BOUND = True


class CoupledProblemSolver:
    """``ProblemSolver<NS,UPS,UD,TPB,SPB,NC,CBW,CCI,NUP,...,Algorithm,Precision>`` for a built-in coupled system."""

    def __init__(self, system: str, NS: int, UPS: int, CBW: int = 0, CCI: int = 0, NDO: int = 0, Algorithm: int = 1,
                 Device: int = 0, TPB: int = 64, SPB: int = 1, SystemOffset: int = 0, Precision=np.float64, Devices=None, Tile: int = 0):
        """``Devices``: one solver object over several devices (systems dealt out in tiles of ``Tile``; mpgos_coupled_create_sharded)."""
        self._lib = lib = load_library()
        _bind(lib)
        self._h = None
        d = _CDesc()
        _check(lib, lib.mpgos_coupled_builtin_desc(system.encode(), UPS, CBW, CCI, C.byref(d)))
        self.dtype = np.dtype(Precision)
        d.NS, d.NDO, d.Algorithm, d.PrecisionBytes, d.Device, d.TPB, d.SPB, d.SystemOffset = NS, NDO, int(Algorithm), self.dtype.itemsize, Device, TPB, SPB, SystemOffset
        self.desc = d
        self._launcher = lib.mpgos_coupled_builtin_launcher_p(system.encode(), UPS, CBW, CCI, int(Algorithm), int(NDO > 0), self.dtype.itemsize)
        if not self._launcher:
            raise MpgosError(lib.mpgos_last_error().decode())
        h = C.c_void_p()
        if Devices is not None:
            devs = (C.c_int * len(Devices))(*[int(x) for x in Devices])
            _check(lib, lib.mpgos_coupled_create_sharded(C.byref(d), devs, len(Devices), int(Tile), C.byref(h)))
        else:
            _check(lib, lib.mpgos_coupled_create(C.byref(d), C.byref(h)))
        self._h = h
        self._views: Dict[int, np.ndarray] = {}

    def close(self):
        if self._h is not None:
            self._lib.mpgos_coupled_destroy(self._h)
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

    def host(self, variable: int) -> np.ndarray:
        """numpy view of a pinned mirror: unit scope [serial, system, unit], system scope [serial, system]; CouplingMatrix is
        the PACKED matrix (flat) -- use SetCouplingMatrix / GetCouplingMatrix for (k, row, col) access."""
        variable = int(variable)
        if variable in self._views:
            return self._views[variable]
        n = C.c_size_t()
        p = self._lib.mpgos_coupled_host_ptr(self._h, variable, C.byref(n))
        V = CoupledVariables
        ints = variable in (V.IntegerGlobalParameters, V.IntegerUnitAccessories, V.IntegerSystemAccessories, V.CouplingIndex, V.DenseIndex)
        dt = np.dtype(np.int32) if ints else self.dtype
        if n.value == 0:
            arr = np.zeros(0, dt)
        else:
            if not p:
                raise MpgosError(self._lib.mpgos_last_error().decode())
            arr = np.frombuffer((C.c_char * (n.value * dt.itemsize)).from_address(p), dtype=dt)
        d = self.desc
        NS, U = d.NS, d.UPS
        shape = {V.TimeDomain: (2, NS), V.ActualTime: (NS,), V.ActualState: (d.UD, NS, U), V.UnitParameters: (d.NUP, NS, U),
                 V.SystemParameters: (d.NSP, NS), V.GlobalParameters: (d.NGP,), V.IntegerGlobalParameters: (d.NiGP,),
                 V.UnitAccessories: (d.NUA, NS, U), V.IntegerUnitAccessories: (d.NiUA, NS, U), V.SystemAccessories: (d.NSA, NS),
                 V.IntegerSystemAccessories: (d.NiSA, NS), V.CouplingMatrix: (n.value,), V.CouplingStrength: (d.NC, NS),
                 V.CouplingIndex: (d.NC,), V.DenseIndex: (NS,), V.DenseTime: (d.NDO, NS), V.DenseState: (d.NDO, d.UD, NS, U)}[CoupledVariables(variable)]
        arr = arr.reshape(shape)
        self._views[variable] = arr
        return arr

    # SetHost(CouplingNumber, CouplingMatrix, Row, Col, Value), Interface.cuh:993-1081 (packs by matrix type, rejects out-of-band entries)
    def SetCouplingMatrix(self, k: int, row: int, col: int, value: float):
        _check(self._lib, self._lib.mpgos_coupled_set_matrix_entry(self._h, k, row, col, float(value)))

    def GetCouplingMatrix(self, k: int, row: int, col: int) -> float:
        v = C.c_double()
        _check(self._lib, self._lib.mpgos_coupled_get_matrix_entry(self._h, k, row, col, C.byref(v)))
        return v.value

    def SolverOption(self, option: int, *args):
        if len(args) == 1:
            _check(self._lib, self._lib.mpgos_coupled_set_option(self._h, int(option), float(args[0])))
        elif len(args) == 2:
            _check(self._lib, self._lib.mpgos_coupled_set_option_indexed(self._h, int(option), int(args[0]), float(args[1])))
        else:
            raise TypeError("SolverOption takes (option, value) or (option, index, value)")

    def SynchroniseFromHostToDevice(self, variable: int = CoupledVariables.All):
        _check(self._lib, self._lib.mpgos_coupled_sync_h2d(self._h, int(variable)))

    def SynchroniseFromDeviceToHost(self, variable: int = CoupledVariables.All):
        _check(self._lib, self._lib.mpgos_coupled_sync_d2h(self._h, int(variable)))

    def Solve(self):
        _check(self._lib, self._lib.mpgos_coupled_solve(self._h, self._launcher))

    def SolveTimed(self, phases: int = 1) -> float:
        ms = C.c_float()
        _check(self._lib, self._lib.mpgos_coupled_solve_timed(self._h, self._launcher, phases, C.byref(ms)))
        return ms.value

    def InsertSynchronisationPoint(self):
        _check(self._lib, self._lib.mpgos_coupled_insert_sync_point(self._h))

    def SynchroniseSolver(self):
        _check(self._lib, self._lib.mpgos_coupled_sync_solver(self._h))

    def SynchroniseDevice(self):
        _check(self._lib, self._lib.mpgos_coupled_sync_device(self._h))

    # copies between caller-owned (pinned) host arrays and the device buffers, asynchronous on the solver's stream
    def Upload(self, variable: int, host_ptr: int, nbytes: int):
        _check(self._lib, self._lib.mpgos_coupled_upload(self._h, int(variable), host_ptr, nbytes))

    def Download(self, variable: int, host_ptr: int, nbytes: int):
        _check(self._lib, self._lib.mpgos_coupled_download(self._h, int(variable), host_ptr, nbytes))

    def StepCounters(self):
        att, acc = np.zeros(self.desc.NS, np.uint64), np.zeros(self.desc.NS, np.uint64)
        _check(self._lib, self._lib.mpgos_coupled_read_step_counters(self._h, att.ctypes.data, acc.ctypes.data))
        return att, acc

    def ResetStepCounters(self):
        _check(self._lib, self._lib.mpgos_coupled_reset_step_counters(self._h))

    def SaveState(self):
        _check(self._lib, self._lib.mpgos_coupled_save_state(self._h))

    def RestoreState(self):
        _check(self._lib, self._lib.mpgos_coupled_restore_state(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.mpgos_coupled_stream(self._h) or 0)


_FIELD_TO_VAR = {"TimeDomain": "TimeDomain", "ActualState": "ActualState", "ActualTime": "ActualTime", "UnitParameters": "UnitParameters",
                 "SystemParameters": "SystemParameters", "GlobalParameters": "GlobalParameters", "IntegerGlobalParameters": "IntegerGlobalParameters",
                 "UnitAccessories": "UnitAccessories", "IntegerUnitAccessories": "IntegerUnitAccessories", "SystemAccessories": "SystemAccessories",
                 "IntegerSystemAccessories": "IntegerSystemAccessories", "CouplingStrength": "CouplingStrength", "CouplingIndex": "CouplingIndex",
                 "DenseIndex": "DenseIndex", "DenseTime": "DenseTime", "DenseState": "DenseState"}


def apply_coupled_workload(ps: CoupledProblemSolver, wl, buffers=None):
    b = wl.buffers if buffers is None else buffers
    sh = wl.shape
    o = wl.options.resolved(sh, b.NS)
    for field, var in _FIELD_TO_VAR.items():
        src = getattr(b, field)
        if src.size:
            ps.host(CoupledVariables[var])[...] = src
    # the coupling matrix goes through the packing entry point, as SetHost does in the reference
    band = sh.CBW if sh.CBW > 0 else None
    for k in range(sh.NC):
        for r in range(sh.UPS):
            for c in range(sh.UPS):
                if band is not None:
                    dist = min((c - r) % sh.UPS, (r - c) % sh.UPS)
                    if dist > band:
                        continue
                ps.SetCouplingMatrix(k, r, c, b.CouplingMatrix[k, c, r])
    for opt in ("InitialTimeStep", "ActiveSystems", "MaximumTimeStep", "MinimumTimeStep", "TimeStepGrowLimit", "TimeStepShrinkLimit",
                "DenseOutputSaveFrequency", "DenseOutputMinimumTimeStep"):
        ps.SolverOption(CoupledOptions[opt], getattr(o, opt))
    for i in range(sh.UD):
        ps.SolverOption(CoupledOptions.RelativeTolerance, i, o.RelativeTolerance[i])
        ps.SolverOption(CoupledOptions.AbsoluteTolerance, i, o.AbsoluteTolerance[i])
    for i in range(sh.NE):
        ps.SolverOption(CoupledOptions.EventTolerance, i, o.EventTolerance[i])
        ps.SolverOption(CoupledOptions.EventDirection, i, int(o.EventDirection[i]))


def read_back_coupled(ps: CoupledProblemSolver, wl):
    from .workloads_coupled import CoupledBuffers
    ps.SynchroniseFromDeviceToHost(CoupledVariables.All)
    ps.InsertSynchronisationPoint()
    ps.SynchroniseSolver()
    out = CoupledBuffers(wl.shape, ps.desc.NS)
    for field, var in _FIELD_TO_VAR.items():
        src = ps.host(CoupledVariables[var])
        if src.size:
            getattr(out, field)[...] = src
    out.CouplingMatrix[...] = wl.buffers.CouplingMatrix
    return out


def run_coupled_workload(wl, buffers=None, phases: Optional[int] = None, device: int = 0, tensor_core_coupling: Optional[int] = None,
                         dispatch_order: Optional[int] = None, devices=None, tile: int = 0):
    """Integrate a coupled workload on the GPU through the public API; returns (buffers_after, attempted, accepted)."""
    b = wl.buffers if buffers is None else buffers
    sh = wl.shape
    with CoupledProblemSolver(wl.name, b.NS, sh.UPS, CBW=sh.CBW, CCI=sh.CCI, NDO=sh.NDO, Algorithm=sh.Algorithm, Device=device,
                              TPB=sh.TPB, SPB=sh.SPB, Precision=sh.dtype, Devices=devices, Tile=tile) as ps:
        apply_coupled_workload(ps, wl, b)
        if tensor_core_coupling is not None:
            ps.SolverOption(CoupledOptions.TensorCoreCoupling, tensor_core_coupling)
        if dispatch_order is not None:
            ps.SolverOption(CoupledOptions.DispatchOrder, dispatch_order)
        ps.SynchroniseFromHostToDevice(CoupledVariables.All)
        for _ in range(wl.phases if phases is None else phases):
            ps.Solve()
        out = read_back_coupled(ps, wl)
        att, acc = ps.StepCounters()
    return out, att.astype(np.int64), acc.astype(np.int64)
