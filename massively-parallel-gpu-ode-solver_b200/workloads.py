"""Synthetic parameter sweeps of the BASELINE.json configurations.

Each generator restates what the corresponding reference tutorial's ``main`` / ``FillSolverObject``
puts into the solver object (cited per function) and returns a :class:`Workload`: the template
shape, the solver options and host buffers in the reference's layout (structure of arrays,
instance index fastest, ``SourceCodes/SingleSystem_PerThread_Interface.cuh:712,814``).

Host-side only (numpy); nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import numpy as np

RK4, RKCK45 = 0, 1
ALGORITHM_NAMES = {RK4: "rk4", RKCK45: "rkck45"}


@dataclasses.dataclass
class Shape:
    """Template arguments of the reference's ``ProblemSolver`` (Interface.cuh:117)."""
    SD: int
    NCP: int = 0
    NSP: int = 0
    NISP: int = 0
    NE: int = 0
    NA: int = 0
    NIA: int = 0
    NDO: int = 0
    Algorithm: int = RKCK45
    dtype: type = np.float64


@dataclasses.dataclass
class Options:
    """Run-time solver options with the reference's defaults (Interface.cuh:475-496)."""
    InitialTimeStep: float = 1e-2
    MaximumTimeStep: float = 1.0e6
    MinimumTimeStep: float = 1.0e-12
    TimeStepGrowLimit: float = 5.0
    TimeStepShrinkLimit: float = 0.1
    DenseOutputMinimumTimeStep: float = 0.0
    DenseOutputSaveFrequency: int = 1
    ActiveNumberOfThreads: Optional[int] = None      # None = NT
    ThreadsPerBlock: Optional[int] = None
    RelativeTolerance: Optional[np.ndarray] = None   # [SD], default 1e-8
    AbsoluteTolerance: Optional[np.ndarray] = None   # [SD], default 1e-8
    EventTolerance: Optional[np.ndarray] = None      # [NE], default 1e-6
    EventDirection: Optional[np.ndarray] = None      # [NE], default 0

    def resolved(self, shape: Shape, NT: int) -> "Options":
        o = dataclasses.replace(self)
        dt = shape.dtype
        if o.ActiveNumberOfThreads is None:
            o.ActiveNumberOfThreads = NT
        o.RelativeTolerance = np.full(shape.SD, 1e-8, dt) if o.RelativeTolerance is None else np.asarray(o.RelativeTolerance, dt)
        o.AbsoluteTolerance = np.full(shape.SD, 1e-8, dt) if o.AbsoluteTolerance is None else np.asarray(o.AbsoluteTolerance, dt)
        o.EventTolerance = np.full(shape.NE, 1e-6, dt) if o.EventTolerance is None else np.asarray(o.EventTolerance, dt)
        o.EventDirection = np.zeros(shape.NE, np.int32) if o.EventDirection is None else np.asarray(o.EventDirection, np.int32)
        return o


class Buffers:
    """Host mirrors of one solver object; arrays are [serial, instance] so that element
    (instance, serial) sits at ``instance + serial*NT`` exactly as in the reference."""

    FIELDS = ("TimeDomain", "ActualState", "ActualTime", "ControlParameters", "SharedParameters",
              "IntegerSharedParameters", "Accessories", "IntegerAccessories",
              "DenseIndex", "DenseTime", "DenseState")

    def __init__(self, shape: Shape, NT: int):
        dt = shape.dtype
        self.NT = NT
        self.TimeDomain = np.zeros((2, NT), dt)
        self.ActualState = np.zeros((shape.SD, NT), dt)
        self.ActualTime = np.zeros(NT, dt)
        self.ControlParameters = np.zeros((shape.NCP, NT), dt)
        self.SharedParameters = np.zeros(shape.NSP, dt)
        self.IntegerSharedParameters = np.zeros(shape.NISP, np.int32)
        self.Accessories = np.zeros((shape.NA, NT), dt)
        self.IntegerAccessories = np.zeros((shape.NIA, NT), np.int32)
        self.DenseIndex = np.zeros(NT, np.int32)
        self.DenseTime = np.zeros((shape.NDO, NT), dt)
        self.DenseState = np.zeros((shape.NDO, shape.SD, NT), dt)

    def copy(self) -> "Buffers":
        b = object.__new__(Buffers)
        b.NT = self.NT
        for f in self.FIELDS:
            setattr(b, f, getattr(self, f).copy())
        return b

    def as_dict(self) -> Dict[str, np.ndarray]:
        return {f: getattr(self, f) for f in self.FIELDS}

    def take(self, idx) -> "Buffers":
        """Sub-sweep made of the listed instances (per-instance arrays are sliced)."""
        idx = np.asarray(idx)
        b = object.__new__(Buffers)
        b.NT = int(idx.size)
        for f in self.FIELDS:
            a = getattr(self, f)
            if f in ("SharedParameters", "IntegerSharedParameters"):
                setattr(b, f, a.copy())
            else:
                setattr(b, f, np.ascontiguousarray(a[..., idx]))
        return b


@dataclasses.dataclass
class Workload:
    name: str            # builtin system name understood by the CUDA library and the oracle
    shape: Shape
    options: Options
    buffers: Buffers
    phases: int          # Solve() calls per bench step
    flops_per_step: int  # algorithmic flops per attempted step
    note: str = ""

    @property
    def NT(self) -> int:
        return self.buffers.NT


def single_precision(wl: "Workload", tolerance: float = 1e-5, event_tolerance: float = 1e-4, minimum_time_step: float = 1e-6) -> "Workload":
    """The same sweep for the FP32 instantiation: buffers cast to float32 and tolerances a float can resolve.  The reference
    ships single-precision settings for none of its tutorials, so these are our choice (stated wherever they are used)."""
    shape = dataclasses.replace(wl.shape, dtype=np.float32)
    b = Buffers(shape, wl.NT)
    for f in Buffers.FIELDS:
        getattr(b, f)[...] = getattr(wl.buffers, f)
    o = dataclasses.replace(wl.options)
    o.RelativeTolerance = np.full(shape.SD, tolerance, np.float32)
    o.AbsoluteTolerance = np.full(shape.SD, tolerance, np.float32)
    if shape.NE:
        o.EventTolerance = np.full(shape.NE, event_tolerance, np.float32)
    o.MinimumTimeStep = max(o.MinimumTimeStep, minimum_time_step)
    return Workload(wl.name, shape, o, b, wl.phases, wl.flops_per_step, wl.note + " (single precision)")


def linspace(B: float, E: float, N: int, dtype=np.float64) -> np.ndarray:
    """``Linspace`` of the tutorials (T2_Lorenz/Lorenz.cu:85-101): x[i] = B + i*((E-B)/(N-1)),
    first and last element set exactly."""
    dt = np.dtype(dtype).type
    x = np.empty(N, dtype)
    x[0] = dt(B)
    if N > 1:
        inc = dt((dt(E) - dt(B)) / dt(N - 1))
        x[1:N - 1] = dt(B) + np.arange(1, N - 1).astype(dtype) * inc
        x[N - 1] = dt(E)
    return x


def logspace(B: float, E: float, N: int) -> np.ndarray:
    """``Logspace`` of T5 (KellerMiksis.cu:134-149): 10^(log10 B + i*incr)."""
    x = np.empty(N, np.float64)
    x[0] = B
    if N > 1:
        eb, ee = math.log10(B), math.log10(E)
        incr = (ee - eb) / (N - 1)
        x[1:N - 1] = np.power(10.0, eb + np.arange(1, N - 1) * incr)
        x[N - 1] = E
    return x


def lorenz(NT: int = 1 << 20, algorithm: int = RKCK45, dtype=np.float64, phases: int = 1) -> Workload:
    """C2: T2_Lorenz/Lorenz.cu:46-61,104-121. rho = Linspace(0, 21, NT), x0 = (10,10,10),
    TD = [0, 10], InitialTimeStep 1e-3, everything else default.  FP32 runs use rtol=atol=1e-5
    (1e-8 is below float epsilon; the reference ships no FP32 settings -- our choice)."""
    shape = Shape(SD=3, NCP=1, Algorithm=algorithm, dtype=dtype)
    b = Buffers(shape, NT)
    b.TimeDomain[0, :] = 0
    b.TimeDomain[1, :] = dtype(0.001 * 10000.0)
    b.ActualState[:, :] = 10.0
    b.ControlParameters[0, :] = linspace(0.0, 21.0, NT, dtype)
    opt = Options(InitialTimeStep=1.0e-3, ThreadsPerBlock=64)
    if dtype == np.float32:
        opt.RelativeTolerance = np.full(3, 1e-5, np.float32)
        opt.AbsoluteTolerance = np.full(3, 1e-5, np.float32)
    flops = 236 if algorithm == RKCK45 else 4 * 9 + 13 * 3 + 4
    return Workload("lorenz", shape, opt, b, phases, flops, "T2_Lorenz rho-sweep")


def _duffing_fill(b: Buffers, NT: int):
    b.TimeDomain[0, :] = 0.0
    b.TimeDomain[1, :] = 2 * 3.14159265358979323846
    b.ActualState[0, :] = -0.5
    b.ActualState[1, :] = -0.1
    b.ActualTime[:] = 0.0
    b.ControlParameters[0, :] = linspace(0.2, 0.3, NT)
    b.SharedParameters[0] = 0.3


def duffing_t1(NT: int = 10000, algorithm: int = RK4, phases: int = 64, ndo: int = 200) -> Workload:
    """C1 / C1b: T1_Reference/Reference.cu:45-79,160-179 -- Duffing with 2 events, 3 accessories,
    1 integer accessory and dense output; the tutorial's non-default options are kept."""
    shape = Shape(SD=2, NCP=1, NSP=1, NE=2, NA=3, NIA=1, NDO=ndo, Algorithm=algorithm)
    b = Buffers(shape, NT)
    _duffing_fill(b, NT)
    opt = Options(InitialTimeStep=1.0e-2, MaximumTimeStep=1.0e3, MinimumTimeStep=1.0e-14,
                  TimeStepGrowLimit=10.0, TimeStepShrinkLimit=0.2,
                  DenseOutputMinimumTimeStep=0.0, DenseOutputSaveFrequency=1, ThreadsPerBlock=64,
                  RelativeTolerance=np.full(2, 1e-9), AbsoluteTolerance=np.full(2, 1e-9),
                  EventTolerance=np.full(2, 1e-6), EventDirection=np.array([-1, 0], np.int32))
    flops = 58 if algorithm == RK4 else 6 * 7 + 55 * 2 + 9 + 6
    return Workload("duffing", shape, opt, b, phases, flops, "T1_Reference Duffing k-sweep")


def poincare_t3(NT: int = 10000, algorithm: int = RK4, phases: int = 8) -> Workload:
    """C1a: the events-free Duffing of T3_PoincareSection/Poincare.cu:17-26,45-72 (same RHS as T1),
    used for the strict fixed-step RK4 gate."""
    shape = Shape(SD=2, NCP=1, NSP=1, Algorithm=algorithm)
    b = Buffers(shape, NT)
    _duffing_fill(b, NT)
    opt = Options(InitialTimeStep=1.0e-2, MaximumTimeStep=1.0e3, MinimumTimeStep=1.0e-14,
                  TimeStepGrowLimit=10.0, TimeStepShrinkLimit=0.2, ThreadsPerBlock=64,
                  RelativeTolerance=np.full(2, 1e-9), AbsoluteTolerance=np.full(2, 1e-9))
    flops = 58 if algorithm == RK4 else 6 * 7 + 55 * 2 + 9 + 6
    return Workload("poincare", shape, opt, b, phases, flops, "T3 Duffing without events")


def impact_t6(NT: int = 1 << 20, phases: int = 128, ndo: int = 0) -> Workload:
    """C3: T6_ImpactDynamics/ImpactDynamics.cu:38,58-66,166-186 -- pressure relief valve with
    impact law; q = Linspace(0.2, 10, NT); each phase ends at the first Poincare crossing."""
    shape = Shape(SD=3, NCP=1, NSP=4, NE=2, NA=3, NIA=2, NDO=ndo, Algorithm=RKCK45)      # ndo > 0: not a tutorial setting (store-efficiency probe)
    b = Buffers(shape, NT)
    b.TimeDomain[0, :] = 0.0
    b.TimeDomain[1, :] = 1.0e10
    b.ActualState[0, :] = 0.2
    b.ControlParameters[0, :] = linspace(0.2, 10.0, NT)
    b.SharedParameters[:] = (1.25, 10.0, 20.0, 0.8)
    opt = Options(ThreadsPerBlock=64, RelativeTolerance=np.full(3, 1e-10), AbsoluteTolerance=np.full(3, 1e-10),
                  EventDirection=np.array([-1, -1], np.int32))
    return Workload("impact", shape, opt, b, phases, 230, "T6 impact dynamics q-sweep")


def keller_miksis_t5(NT: int = 1 << 18, phases: int = 32, ndo: int = 200, fast_math: bool = False) -> Workload:
    """C4: T5_LargeTimeScaleDifferences/KellerMiksis.cu:39,54-58,153-215 -- frequency sweep
    Logspace(20, 1000) kHz at PA1 = 1.5 bar, RE = 10 um; 13 control parameters.  The config adds
    dense output (NDO = 200, DenseOutputMinimumTimeStep 5e-3, dense index reset every phase)."""
    shape = Shape(SD=2, NCP=13, NA=1, NDO=ndo, Algorithm=RKCK45)
    b = Buffers(shape, NT)
    PI = 3.14159265358979323846
    P1, P3, P4, P5, P6, P7, P9 = 1.5, 0.0, 0.0, 0.0, 10.0, 1.0, 1.4
    Pv, Rho, ST = 3.166775638952003e+03, 9.970639504998557e+02, 0.071977583160056
    Vis, CL = 8.902125058209557e-04, 1.497251785455527e+03
    P2 = logspace(20.0, 1000.0, NT)
    Pinf, PA1, PA2, RE = P7 * 1.0e5, P1 * 1.0e5, P3 * 1.0e5, P6 / 1.0e6
    f1 = 2.0 * PI * (P2 * 1000)
    f2 = 2.0 * PI * (P4 * 1000)
    sq = np.power(2.0 * PI / RE / f1, 2.0)
    cp = b.ControlParameters
    b.TimeDomain[0, :] = 0.0
    b.TimeDomain[1, :] = 1.0
    b.ActualState[0, :] = 1.0
    b.ActualState[1, :] = 0.0
    cp[0] = (2.0 * ST / RE + Pinf - Pv) * sq / Rho
    cp[1] = (1.0 - 3.0 * P9) * (2 * ST / RE + Pinf - Pv) * (2.0 * PI / RE / f1) / CL / Rho
    cp[2] = (Pinf - Pv) * sq / Rho
    cp[3] = (2.0 * ST / RE / Rho) * sq
    cp[4] = (4.0 * Vis / Rho / math.pow(RE, 2.0)) * (2.0 * PI / f1)
    cp[5] = PA1 * sq / Rho
    cp[6] = PA2 * sq / Rho
    cp[7] = (RE * f1 * PA1 / Rho / CL) * sq
    cp[8] = (RE * f2 * PA2 / Rho / CL) * sq
    cp[9] = RE * f1 / (2.0 * PI) / CL
    cp[10] = 3.0 * P9
    cp[11] = P4 / P2
    cp[12] = P5
    opt = Options(ThreadsPerBlock=64, RelativeTolerance=np.full(2, 1e-10), AbsoluteTolerance=np.full(2, 1e-10))
    if ndo > 0:
        opt.DenseOutputMinimumTimeStep = 5e-3
    # fast_math: the built-in system whose pow / sincospi / sin / cos come from include/mpgos_b200/fast_math.cuh (opt-in, <= 3 ulp)
    name = ("kellermiksis_dense" if ndo > 0 else "kellermiksis") + ("_fm" if fast_math else "")
    return Workload(name, shape, opt, b, phases, 413, "T5 Keller-Miksis frequency sweep")


def tidprobe(NT: int = 4096, algorithm: int = RKCK45) -> Workload:
    """Test system "tidprobe" (csrc/builtin_systems.cuh): x' = k (tid / NT - x); accessories record the id and NT the callbacks saw."""
    shape = Shape(SD=1, NCP=1, NA=1, NIA=2, Algorithm=algorithm)
    b = Buffers(shape, NT)
    b.TimeDomain[1, :] = 3.0
    b.ActualState[0, :] = linspace(-1.0, 1.0, NT)
    b.ControlParameters[0, :] = linspace(0.5, 4.0, NT)
    return Workload("tidprobe", shape, Options(InitialTimeStep=1e-2), b, 1, 0, "instance-id probe")


def vanderpol_example(NT: int = 1536, algorithm: int = RKCK45, phases: int = 2) -> Workload:
    """The user-written system of examples/vanderpol (drop-in test; not a BASELINE configuration):
    same options and initial data as examples/vanderpol/vanderpol_dropin.cu."""
    shape = Shape(SD=2, NCP=1, NSP=2, NISP=1, NE=2, NA=2, NIA=2, NDO=64, Algorithm=algorithm)
    b = Buffers(shape, NT)
    b.TimeDomain[1, :] = 1.0e3
    b.ActualState[0, :] = 2.0
    b.ControlParameters[0, :] = linspace(0.1, 4.0, NT)
    b.SharedParameters[:] = (0.5, 1.1)
    b.IntegerSharedParameters[:] = 3
    opt = Options(InitialTimeStep=1.0e-2, ThreadsPerBlock=64, DenseOutputMinimumTimeStep=0.0537, DenseOutputSaveFrequency=2,
                  RelativeTolerance=np.full(2, 1e-9), AbsoluteTolerance=np.full(2, 1e-9),
                  EventTolerance=np.full(2, 1e-7), EventDirection=np.array([-1, 0], np.int32))
    return Workload("vanderpol", shape, opt, b, phases, 6 * 9 + 55 * 2 + 9 + 6, "user-defined Van der Pol (examples/vanderpol)")


GENERATORS = {
    "lorenz": lorenz,
    "duffing": duffing_t1,
    "poincare": poincare_t3,
    "impact": impact_t6,
    "kellermiksis": keller_miksis_t5,
}
