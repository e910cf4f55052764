"""Small parity cases shared by the golden-vector generator, the CPU tests and the GPU tests.

Each case is a sub-sweep of a BASELINE configuration (same generators, same options, fewer
instances / phases) that the CPU checker finishes in about a second."""
import copy

import numpy as np


def subsample(wl, n, seed=None):
    """Every k-th instance of the full sweep (keeps first and last), as a new workload."""
    import copy
    idx = np.unique(np.linspace(0, wl.NT - 1, n).round().astype(np.int64))
    out = copy.copy(wl)
    out.buffers = wl.buffers.take(idx)
    return out


def make_cases(W):
    """name -> (workload, phases).  W is the product's ``workloads`` module."""
    c = {}
    c["lorenz_rk4_f64"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RK4), 96), 1)
    c["lorenz_rkck45_f64"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RKCK45), 192), 1)
    c["lorenz_rkck45_f32"] = (subsample(W.lorenz(NT=1 << 14, algorithm=W.RKCK45, dtype=np.float32), 96), 1)
    c["poincare_rk4_f64"] = (subsample(W.poincare_t3(NT=4096, algorithm=W.RK4), 96), 4)
    c["poincare_rkck45_f64"] = (subsample(W.poincare_t3(NT=4096, algorithm=W.RKCK45), 96), 4)
    c["duffing_rk4_f64"] = (subsample(W.duffing_t1(NT=4096, algorithm=W.RK4), 96), 3)
    c["duffing_rkck45_f64"] = (subsample(W.duffing_t1(NT=4096, algorithm=W.RKCK45), 96), 3)
    c["impact_rkck45_f64"] = (subsample(W.impact_t6(NT=1 << 14), 128), 6)
    c["kellermiksis_rkck45_f64"] = (subsample(W.keller_miksis_t5(NT=4096, ndo=0), 64), 2)
    c["kellermiksis_dense_rkck45_f64"] = (subsample(W.keller_miksis_t5(NT=4096, ndo=200), 64), 2)
    return c


def random_options(rng, wl, trial):
    """A copy of the workload's options with every run-time knob redrawn; extreme draws are deliberate."""
    sh = wl.shape
    o = copy.copy(wl.options)
    o.InitialTimeStep = float(10.0 ** rng.uniform(-5, -1))
    o.MaximumTimeStep = float(10.0 ** rng.uniform(-2, 3))
    # one trial in three pushes the lower limit up into the working range: steps are then accepted by force
    o.MinimumTimeStep = float(10.0 ** (rng.uniform(-5, -3) if trial % 3 == 0 else rng.uniform(-14, -8)))
    o.TimeStepGrowLimit = float(rng.uniform(1.2, 10.0))
    o.TimeStepShrinkLimit = float(rng.uniform(0.05, 0.7))
    o.RelativeTolerance = 10.0 ** rng.uniform(-11, -4, sh.SD)
    o.AbsoluteTolerance = 10.0 ** rng.uniform(-11, -4, sh.SD)
    if sh.NE:
        o.EventTolerance = 10.0 ** rng.uniform(-8, -3, sh.NE)
        if wl.name != "impact":       # the impact law needs its falling-edge detection to keep the valve above the seat
            o.EventDirection = rng.integers(-1, 2, sh.NE).astype(np.int32)
    if sh.NDO:
        o.DenseOutputSaveFrequency = int(rng.integers(1, 6))
        o.DenseOutputMinimumTimeStep = float(rng.choice([0.0, 10.0 ** rng.uniform(-4, -1)]))
    o.ActiveNumberOfThreads = int(rng.integers(wl.NT // 2, wl.NT + 1))
    if wl.name.startswith("kellermiksis"):
        # dimensionless bubble dynamics: keep the controller inside the range where the collapse stays finite
        o.MaximumTimeStep = 1.0e6
        o.MinimumTimeStep = min(o.MinimumTimeStep, 1e-10)
    return o


RESULT_FIELDS = ("TimeDomain", "ActualState", "ActualTime", "ControlParameters", "Accessories",
                 "IntegerAccessories", "DenseIndex", "DenseTime", "DenseState")


def make_coupled_cases(WC):
    """name -> (reference-host build tag, workload, phases) for the block-per-system module.  Between them the builds
    exercise all four launch variants of the reference (SingleSystem/MultipleSystems x Single/MultipleBlockLaunch) and
    all four coupling-matrix layouts (full, banded, circulant-banded, full circulant)."""
    c = {}
    c["bubble_u8_rkck45"] = ("bubble_ensemble_u8_rkck45", WC.bubble_ensemble(NS=4, UPS=8, amplitude=(0.2, 1.2)), 1)
    c["bubble_u16_rkck45"] = ("bubble_ensemble_u16_rkck45", WC.bubble_ensemble(NS=5, UPS=16, amplitude=(0.2, 1.2)), 2)
    c["bubble_u16_rk4"] = ("bubble_ensemble_u16_rk4", WC.bubble_ensemble(NS=2, UPS=16, algorithm=WC.RK4, amplitude=(0.2, 0.6)), 1)
    c["ring_b1c1_rkck45"] = ("duffing_ring_u14_b1c1_rkck45", WC.duffing_ring(NS=7, CBW=1, CCI=1), 2)
    c["ring_b1c0_rkck45"] = ("duffing_ring_u14_b1c0_rkck45", WC.duffing_ring(NS=5, CBW=1, CCI=0), 1)
    c["ring_b0c1_rk4"] = ("duffing_ring_u14_b0c1_rk4", WC.duffing_ring(NS=3, CBW=0, CCI=1, algorithm=WC.RK4), 1)
    c["ring_b0c0_rkck45"] = ("duffing_ring_u14_b0c0_rkck45", WC.duffing_ring(NS=4, CBW=0, CCI=0), 2)
    # single precision: the reference's own float instantiation of the ring tutorial (its bubble tutorial does not compile in float)
    c["ring_b1c1_rkck45_f32"] = ("duffing_ring_u14_b1c1_rkck45_f32", WC.duffing_ring(NS=7, CBW=1, CCI=1, dtype=np.float32), 2)
    return c


COUPLED_RESULT_FIELDS = ("TimeDomain", "ActualState", "ActualTime", "UnitAccessories", "IntegerUnitAccessories", "SystemAccessories",
                         "IntegerSystemAccessories", "DenseIndex", "DenseTime", "DenseState")


def random_coupled_options(rng, wl, trial):
    """The block-per-system module's run-time knobs (CP/Interface.cuh:1083-1215) redrawn; same intent as random_options."""
    sh = wl.shape
    o = copy.copy(wl.options)
    # fixed-step RK4 takes (time domain) / InitialTimeStep steps: keep it within what the host emulation of the reference can do
    o.InitialTimeStep = float(10.0 ** (rng.uniform(-2.7, -2) if sh.Algorithm == 0 else rng.uniform(-5, -2)))
    o.MaximumTimeStep = float(10.0 ** rng.uniform(-1, 3))
    o.MinimumTimeStep = float(10.0 ** (rng.uniform(-4.5, -3) if trial % 3 == 0 else rng.uniform(-13, -8)))
    o.TimeStepGrowLimit = float(rng.uniform(1.2, 10.0))
    o.TimeStepShrinkLimit = float(rng.uniform(0.05, 0.7))
    o.RelativeTolerance = 10.0 ** rng.uniform(-10, -4, sh.UD)
    o.AbsoluteTolerance = 10.0 ** rng.uniform(-10, -4, sh.UD)
    if sh.NE:
        o.EventTolerance = 10.0 ** rng.uniform(-7, -3, sh.NE)
        o.EventDirection = rng.integers(-1, 2, sh.NE).astype(np.int32)
    if sh.NDO:
        o.DenseOutputSaveFrequency = int(rng.integers(1, 5))
        o.DenseOutputMinimumTimeStep = float(rng.choice([0.0, 10.0 ** rng.uniform(-4, -1)]))
    o.ActiveSystems = int(rng.integers(max(1, wl.buffers.NS - 1), wl.buffers.NS + 1))
    if wl.name.startswith("bubble"):
        o.MaximumTimeStep = 1.0e6
        o.MinimumTimeStep = min(o.MinimumTimeStep, 1e-10)
    return o
