"""Synthetic inputs of the coupled (block-per-system) configurations and
of the ring tutorial used for the banded / circulant / events / dense-output coverage.

Layouts at this level (the checker ABI and the product's C-ABI use the same):
  unit scope [serial, system, unit], system scope [serial, system], coupling matrix ORIGINAL
  (unpacked) [k, col, row] i.e. element (k, row, col) at row + col*UPS + k*UPS*UPS.
Host-side only (numpy)."""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import numpy as np

RK4, RKCK45 = 0, 1


@dataclasses.dataclass
class CoupledShape:
    """Template arguments of the reference's coupled ``ProblemSolver`` (CoupledSystems_PerBlock_Interface.cuh:135)."""
    UPS: int
    UD: int
    NC: int = 1
    CBW: int = 0
    CCI: int = 0
    NUP: int = 0
    NSP: int = 0
    NGP: int = 0
    NiGP: int = 0
    NUA: int = 0
    NiUA: int = 0
    NSA: int = 0
    NiSA: int = 0
    NE: int = 0
    NDO: int = 0
    Algorithm: int = RKCK45
    TPB: int = 64       # the reference's launch hints (ours chooses its own block shape)
    SPB: int = 1
    dtype: type = np.float64


@dataclasses.dataclass
class CoupledOptions:
    """Run-time options with the reference's defaults (CoupledSystems_PerBlock_Interface.cuh:631-656)."""
    InitialTimeStep: float = 1e-2
    MaximumTimeStep: float = 1.0e6
    MinimumTimeStep: float = 1.0e-12
    TimeStepGrowLimit: float = 5.0
    TimeStepShrinkLimit: float = 0.1
    DenseOutputMinimumTimeStep: float = 0.0
    DenseOutputSaveFrequency: int = 1
    ActiveSystems: Optional[int] = None
    RelativeTolerance: Optional[np.ndarray] = None
    AbsoluteTolerance: Optional[np.ndarray] = None
    EventTolerance: Optional[np.ndarray] = None
    EventDirection: Optional[np.ndarray] = None

    def resolved(self, shape: CoupledShape, NS: int) -> "CoupledOptions":
        o = dataclasses.replace(self)
        if o.ActiveSystems is None:
            o.ActiveSystems = NS
        o.RelativeTolerance = np.full(shape.UD, 1e-8) if o.RelativeTolerance is None else np.asarray(o.RelativeTolerance, np.float64)
        o.AbsoluteTolerance = np.full(shape.UD, 1e-8) if o.AbsoluteTolerance is None else np.asarray(o.AbsoluteTolerance, np.float64)
        o.EventTolerance = np.full(shape.NE, 1e-6) if o.EventTolerance is None else np.asarray(o.EventTolerance, np.float64)
        o.EventDirection = np.zeros(shape.NE, np.int32) if o.EventDirection is None else np.asarray(o.EventDirection, np.int32)
        return o


class CoupledBuffers:
    FIELDS = ("TimeDomain", "ActualState", "ActualTime", "UnitParameters", "SystemParameters", "GlobalParameters",
              "IntegerGlobalParameters", "UnitAccessories", "IntegerUnitAccessories", "SystemAccessories",
              "IntegerSystemAccessories", "CouplingMatrix", "CouplingStrength", "CouplingIndex",
              "DenseIndex", "DenseTime", "DenseState")
    PER_SYSTEM = ("TimeDomain", "ActualState", "ActualTime", "UnitParameters", "SystemParameters", "UnitAccessories",
                  "IntegerUnitAccessories", "SystemAccessories", "IntegerSystemAccessories", "CouplingStrength",
                  "DenseIndex", "DenseTime", "DenseState")

    def __init__(self, sh: CoupledShape, NS: int):
        f, i = np.dtype(sh.dtype), np.int32
        self.NS = NS
        self.TimeDomain = np.zeros((2, NS), f)
        self.ActualState = np.zeros((sh.UD, NS, sh.UPS), f)
        self.ActualTime = np.zeros(NS, f)
        self.UnitParameters = np.zeros((sh.NUP, NS, sh.UPS), f)
        self.SystemParameters = np.zeros((sh.NSP, NS), f)
        self.GlobalParameters = np.zeros(sh.NGP, f)
        self.IntegerGlobalParameters = np.zeros(sh.NiGP, i)
        self.UnitAccessories = np.zeros((sh.NUA, NS, sh.UPS), f)
        self.IntegerUnitAccessories = np.zeros((sh.NiUA, NS, sh.UPS), i)
        self.SystemAccessories = np.zeros((sh.NSA, NS), f)
        self.IntegerSystemAccessories = np.zeros((sh.NiSA, NS), i)
        self.CouplingMatrix = np.zeros((sh.NC, sh.UPS, sh.UPS), f)      # [k, col, row]
        self.CouplingStrength = np.zeros((sh.NC, NS), f)
        self.CouplingIndex = np.zeros(sh.NC, i)
        self.DenseIndex = np.zeros(NS, i)
        self.DenseTime = np.zeros((sh.NDO, NS), f)
        self.DenseState = np.zeros((sh.NDO, sh.UD, NS, sh.UPS), f)

    def copy(self) -> "CoupledBuffers":
        b = object.__new__(CoupledBuffers)
        b.NS = self.NS
        for f in self.FIELDS:
            setattr(b, f, getattr(self, f).copy())
        return b

    def take(self, idx) -> "CoupledBuffers":
        """Sub-sweep made of the listed systems."""
        idx = np.asarray(idx)
        b = object.__new__(CoupledBuffers)
        b.NS = int(idx.size)
        for f in self.FIELDS:
            a = getattr(self, f)
            if f not in self.PER_SYSTEM:
                setattr(b, f, a.copy())
            elif f in ("ActualState", "UnitParameters", "UnitAccessories", "IntegerUnitAccessories"):
                setattr(b, f, np.ascontiguousarray(a[:, idx, :]))
            elif f == "DenseState":
                setattr(b, f, np.ascontiguousarray(a[:, :, idx, :]))
            else:
                setattr(b, f, np.ascontiguousarray(a[..., idx]))
        return b


@dataclasses.dataclass
class CoupledWorkload:
    name: str
    shape: CoupledShape
    options: CoupledOptions
    buffers: CoupledBuffers
    phases: int
    flops_per_step: int      # algorithmic flops per attempted SYSTEM step
    note: str = ""

    @property
    def NS(self) -> int:
        return self.buffers.NS


def _linspace(B, E, N):
    x = np.empty(N)
    x[0] = B
    if N > 1:
        x[1:N - 1] = B + np.arange(1, N - 1) * ((E - B) / (N - 1))
        x[N - 1] = E
    return x


def bubble_geometry(UPS: int, seed: int = 20200818):
    """Bubble radii and positions: the reference draws them with rand()/random_device seeded by the
    clock (BubbleEnsemble.cu:196-221), i.e. not reproducibly; here: mt19937_64-style generator with a fixed seed, radii
    on the grid 3 + 5 j/10001 um, positions N(0, 10 mm), redrawn while any pair is closer than 100 um."""
    rng = np.random.Generator(np.random.MT19937(seed))
    radii = 3.0 + 5.0 * rng.integers(0, 10001, UPS) / 10001.0
    while True:
        pos = rng.normal(0.0, 10.0, (3, UPS))
        d = np.sqrt(((pos[:, :, None] - pos[:, None, :]) ** 2).sum(axis=0)) * 1000.0
        if (d + np.eye(UPS) * 1e9).min() >= 100.0:
            return radii, pos, d


def bubble_ensemble(NS: int = 1 << 16, UPS: int = 64, algorithm: int = RKCK45, phases: int = 1, ndo: int = 0,
                    amplitude=(0.0, 5.0), frequency_khz: float = 1000.0, dtype=np.float64, fast_math: bool = False) -> CoupledWorkload:
    """C5: Tutorials_CoupledSystems_PerBlock/T2_BubbleEnsemble/BubbleEnsemble.cu:88-97, 225-353 -- NS systems of UPS
    coupled Keller-Miksis bubbles, pressure amplitude = Linspace(0, 5, NS) bar, full dense coupling matrix
    A[r][c] = RE_c^3 / RE_r^2 / dist_um shared by all systems, coupling strength 1 into component 1."""
    sh = CoupledShape(UPS=UPS, UD=2, NC=1, NUP=21, NDO=ndo, Algorithm=algorithm, TPB=UPS, SPB=1, dtype=dtype)
    b = CoupledBuffers(sh, NS)
    radii, pos, dist = bubble_geometry(UPS)
    PI = 3.14159265358979323846
    P3, P4, P5, P7, P9 = 0.0, 0.0, 0.0, 1.0, 1.4
    Pv, Rho, ST, Vis, CL = 3.166775638952003e+03, 9.970639504998557e+02, 0.071977583160056, 8.902125058209557e-04, 1.497251785455527e+03
    P1 = _linspace(amplitude[0], amplitude[1], NS)[:, None]          # [NS, 1]
    P2 = frequency_khz
    P6 = radii[None, :]                                             # [1, UPS]
    Pinf, PA1, PA2, RE = P7 * 1e5, P1 * 1e5, P3 * 1e5, P6 / 1e6
    f1, f2 = 2.0 * PI * (P2 * 1000), 2.0 * PI * (P4 * 1000)
    sq = np.power(2.0 * PI / RE / f1, 2.0)
    up = b.UnitParameters
    one = np.ones((NS, UPS))
    up[0] = (2.0 * ST / RE + Pinf - Pv) * sq / Rho * one
    up[1] = (1.0 - 3.0 * P9) * (2 * ST / RE + Pinf - Pv) * (2.0 * PI / RE / f1) / CL / Rho * one
    up[2] = (Pinf - Pv) * sq / Rho * one
    up[3] = (2.0 * ST / RE / Rho) * sq * one
    up[4] = (4.0 * Vis / Rho / np.power(RE, 2.0)) * (2.0 * PI / f1) * one
    up[5] = PA1 * sq / Rho
    up[6] = PA2 * sq / Rho * one
    up[7] = (RE * f1 * PA1 / Rho / CL) * sq
    up[8] = (RE * f2 * PA2 / Rho / CL) * sq * one
    up[9] = RE * f1 / (2.0 * PI) / CL * one
    up[10] = 3.0 * P9
    up[11] = P4 / P2
    up[12] = P5
    up[13] = 2.0 * PI / f1
    up[14] = RE * one
    up[15] = P1 * one
    up[16] = P2
    up[17], up[18], up[19] = P3, P4, P5
    up[20] = P6 * one
    b.TimeDomain[1, :] = 1.0
    b.ActualState[0] = 1.0
    b.CouplingStrength[:] = 1.0
    b.CouplingIndex[:] = 1
    with np.errstate(divide="ignore"):
        A = np.power(radii[None, :], 3) / np.power(radii[:, None], 2) / dist      # A[row, col]
    A[np.arange(UPS), np.arange(UPS)] = 0.0
    b.CouplingMatrix[0] = A.T                                                    # stored [col, row]
    opt = CoupledOptions(InitialTimeStep=1e-3, MinimumTimeStep=1e-13,
                         RelativeTolerance=np.array([1e-10, 1e-8]), AbsoluteTolerance=np.array([1e-10, 1e-8]))
    if np.dtype(dtype) == np.float32:      # the reference ships no single-precision settings: tolerances a float can resolve (our choice)
        opt = CoupledOptions(InitialTimeStep=1e-3, MinimumTimeStep=1e-7,
                             RelativeTolerance=np.array([1e-5, 1e-5]), AbsoluteTolerance=np.array([1e-5, 1e-5]))
    flops = 6 * UPS * (48 + 7 + 2 * UPS + 3) + UPS * 110 + 9 + UPS * 6
    # fast_math: the built-in system with the opt-in transcendentals of include/mpgos_b200/fast_math.cuh (UPS 16 and 64)
    return CoupledWorkload("bubble_ensemble_fm" if fast_math else "bubble_ensemble", sh, opt, b, phases, flops, "T2_BubbleEnsemble amplitude sweep")


def duffing_ring(NS: int = 9, CBW: int = 1, CCI: int = 1, algorithm: int = RKCK45, phases: int = 1, ndo: int = 100,
                 TPB: int = 32, SPB: int = 3, UPS: int = 14, dtype=np.float64) -> CoupledWorkload:
    """Tutorials_CoupledSystems_PerBlock/T1_RingOfDuffingOscillators/RingOfDuffingOscillators.cu:79-106, 215-300 -- rings
    of 14 Duffing oscillators, two coupling terms through a banded / circulant matrix, one event per unit (second
    local maximum of any unit ends the phase), system accessories tracking min/max dt, dense output."""
    sh = CoupledShape(UPS=UPS, UD=2, NC=2, CBW=CBW, CCI=CCI, NUP=3, NGP=1, NUA=3, NiUA=1, NSA=2, NiSA=1, NE=1, NDO=ndo,
                      Algorithm=algorithm, TPB=TPB, SPB=SPB, dtype=dtype)
    b = CoupledBuffers(sh, NS)
    n = int(math.ceil(math.sqrt(NS)))
    icx, icy = _linspace(-8.0, 8.0, n), _linspace(-20.0, 40.0, n)
    x0 = np.repeat(icx, n)[:NS]
    y0 = np.tile(icy, n)[:NS]
    PI = 3.14159265358979323846
    b.TimeDomain[1, :] = 10.0 * 2.0 * PI / 0.5
    bw = max(CBW, 1)
    b.CouplingStrength[:] = 0.004 / (2.0 * bw)
    b.SystemAccessories[:] = 1.0e-2
    b.ActualState[:] = -5.0
    b.UnitAccessories[0:2] = -5.0
    b.ActualState[0, :, 11] = x0
    b.ActualState[1, :, 11] = y0
    b.UnitAccessories[0, :, 11] = x0
    b.UnitAccessories[1, :, 11] = y0
    b.UnitParameters[0], b.UnitParameters[1], b.UnitParameters[2] = 0.5, 0.24, 13.633
    b.CouplingIndex[:] = (0, 1)
    b.GlobalParameters[0] = 0.004
    A = np.zeros((UPS, UPS))
    for r in range(UPS):
        for d in range(-bw, bw + 1):
            if d != 0:
                A[r, (r + d) % UPS] = 1.0
    b.CouplingMatrix[0] = A.T
    b.CouplingMatrix[1] = A.T
    opt = CoupledOptions(InitialTimeStep=1e-2, RelativeTolerance=np.array([1e-8, 1e-8]), AbsoluteTolerance=np.array([1e-8, 1e-8]),
                         EventDirection=np.array([-1], np.int32), DenseOutputMinimumTimeStep=1e-6)
    if np.dtype(dtype) == np.float32:
        opt = dataclasses.replace(opt, RelativeTolerance=np.array([1e-5, 1e-5]), AbsoluteTolerance=np.array([1e-5, 1e-5]), MinimumTimeStep=1e-6,
                                  EventTolerance=np.array([1e-4]), DenseOutputMinimumTimeStep=1e-4)
    return CoupledWorkload("duffing_ring", sh, opt, b, phases, 0, "T1_RingOfDuffingOscillators")


def swing_ring(NS: int = 37, phases: int = 2, t_end: float = 30.0) -> CoupledWorkload:
    """examples/swing_ring: the user-defined coupled system of the drop-in example (ring of second-order phase
    oscillators, two couplings through a banded wrap-around matrix, one event per unit, the third local maximum of
    unit 0 ends a phase, Finalization moves the time domain).  Same inputs as examples/swing_ring/swing_ring_dropin.cu
    receives from the test."""
    UPS, CBW = 40, 2
    sh = CoupledShape(UPS=UPS, UD=2, NC=2, CBW=CBW, CCI=0, NUP=2, NSP=1, NGP=1, NiGP=1, NUA=1, NiUA=1, NSA=1, NiSA=1, NE=1, NDO=32,
                      Algorithm=RKCK45, TPB=64, SPB=1)
    b = CoupledBuffers(sh, NS)
    u = np.arange(UPS)
    b.TimeDomain[1, :] = t_end
    b.UnitParameters[0] = (0.4 * np.sin(2.0 * np.pi * u / UPS + 1.0))[None, :]
    b.UnitParameters[1] = (0.25 + 0.1 * np.cos(2.0 * np.pi * u / UPS))[None, :]
    gain = _linspace(0.6, 3.0, NS)
    b.CouplingStrength[0], b.CouplingStrength[1] = gain, gain
    b.ActualState[0] = 0.8 * np.cos(2.0 * np.pi * u / UPS)[None, :] * (1.0 + np.arange(NS)[:, None] / NS)
    b.GlobalParameters[0] = 1.3
    b.IntegerGlobalParameters[0] = 1
    b.CouplingIndex[:] = (1, 1)
    A = np.zeros((2, UPS, UPS))
    for r in range(UPS):
        for d in range(-CBW, CBW + 1):
            if d != 0:
                A[0, r, (r + d) % UPS] = 1.0 / (1.0 + abs(d))
                A[1, r, (r + d) % UPS] = (1.0 + 0.05 * d) / (1.0 + abs(d))
    b.CouplingMatrix[0] = A[0].T
    b.CouplingMatrix[1] = A[1].T
    opt = CoupledOptions(InitialTimeStep=1e-3, RelativeTolerance=np.array([1e-9, 1e-9]), AbsoluteTolerance=np.array([1e-9, 1e-9]),
                         EventTolerance=np.array([1e-7]), EventDirection=np.array([-1], np.int32), DenseOutputMinimumTimeStep=0.05)
    return CoupledWorkload("swing_ring", sh, opt, b, phases, 0, "examples/swing_ring")


def parameter_drift(NS: int = 40) -> CoupledWorkload:
    """Test system "parameter_drift" (csrc/builtin_coupled_systems.cuh): callbacks write unit and system parameters during the phase."""
    sh = CoupledShape(UPS=16, UD=1, NC=1, NUP=1, NSP=1, Algorithm=RKCK45, TPB=16, SPB=1)
    b = CoupledBuffers(sh, NS)
    rng = np.random.default_rng(5)
    b.TimeDomain[1, :] = _linspace(2.0, 6.0, NS)
    b.ActualState[0] = rng.uniform(-1.0, 1.0, (NS, 16))
    b.UnitParameters[0] = rng.uniform(0.5, 3.0, (NS, 16))
    b.SystemParameters[0] = _linspace(0.0, 1.0, NS)
    b.CouplingStrength[0] = 0.05
    b.CouplingIndex[:] = 0
    A = rng.uniform(0.0, 1.0, (16, 16)) * (1.0 - np.eye(16))
    b.CouplingMatrix[0] = A.T
    opt = CoupledOptions(InitialTimeStep=1e-3, RelativeTolerance=np.array([1e-9]), AbsoluteTolerance=np.array([1e-9]))
    return CoupledWorkload("parameter_drift", sh, opt, b, 1, 0, "parameter-writing callbacks")
