"""Chaotic cases (north_star: "compared over a short horizon and on ensemble statistics such as Poincare-section
histograms").  A chaotic orbit amplifies the O(1e-16) differences between two correct realisations of the same algorithm
(FMA contraction, libm vs the CUDA math library, 1-ulp step-size factors) until individual instances decorrelate; what
must agree is (a) the trajectory while the amplification is still small and (b) the statistics of the attractor.
The thresholds are calibrated on two CPU realisations of the reference arithmetic (with and without FMA contraction):
they agree instance by instance on ~90 % of the impact sweep (periodic windows) and to 0.007 in the L1 distance of
the bifurcation-diagram histogram."""
import copy

import numpy as np
import pytest

from cases import subsample

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


def test_impact_bifurcation_diagram_statistics(pkg, workloads, oracle):
    """C3 (T6 impact dynamics): after a transient of 48 phases, the maximum valve lift of each of the next 32 phases
    (one phase = one Poincare return; ACC[1]) over a q-sweep -- the bifurcation diagram of the reference's
    PressureReliefValveTestBifurcationdiagram.m -- as a 2-D histogram, GPU against the CPU checker."""
    wl = subsample(workloads.impact_t6(NT=1 << 14), 768)
    transient, keep = 48, 32

    def collect(run):
        b = run(wl.buffers.copy(), transient)
        vals = []
        for _ in range(keep):
            b = run(b, 1)
            vals.append(b.Accessories[1].copy())
        return np.array(vals)                      # [phase, instance]

    gpu = collect(lambda b, p: pkg.run_workload(wl, buffers=b, phases=p)[0])
    cpu = collect(lambda b, p: oracle.run_oracle(wl, buffers=b, phases=p, host_threads=8)[0])
    q = wl.buffers.ControlParameters[0]
    assert np.isfinite(gpu).all() and (gpu >= 0).all()

    # periodic windows: the same numbers instance by instance, phase by phase
    per_instance = np.abs(gpu - cpu).max(axis=0)
    assert (per_instance < 1e-6).mean() >= 0.85
    # everywhere, chaotic windows included: the same attractor
    top = 1.01 * max(gpu.max(), cpu.max())

    def hist(v):
        H, _, _ = np.histogram2d(np.tile(q, (keep, 1)).ravel(), v.ravel(), bins=[48, 64], range=[[q.min(), q.max()], [0.0, top]])
        return H / H.sum()

    assert np.abs(hist(gpu) - hist(cpu)).sum() < 0.03
    # and the same envelope of the diagram (largest lift per q bin)
    bins = np.minimum(((q - q.min()) / (q.max() - q.min()) * 48).astype(int), 47)
    env_g = np.array([gpu[:, bins == k].max() for k in range(48)])
    env_c = np.array([cpu[:, bins == k].max() for k in range(48)])
    assert rel_err(env_g, env_c).max() < 0.05


def test_lorenz_strange_attractor_short_horizon_and_statistics(pkg, workloads, oracle):
    """Lorenz at rho = 28 (strange attractor; the C2 sweep stops at rho = 21 to stay comparable point by point).
    Short horizon (T = 2, amplification ~e^(0.9*2)): endpoints to 1e-7 and identical step counts on at least 90 % of the instances;
    long horizon (T = 40): instances decorrelate, the distribution of z on the attractor must not."""
    NT = 2048
    wl = workloads.lorenz(NT=NT, algorithm=workloads.RKCK45)
    wl = copy.copy(wl)
    wl.buffers = wl.buffers.copy()
    wl.buffers.ControlParameters[0, :] = 28.0
    wl.buffers.ActualState[0, :] = 10.0 + 1e-3 * np.arange(NT)          # a line of initial conditions
    wl.buffers.TimeDomain[1, :] = 2.0
    ref, ratt, _ = oracle.run_oracle(wl, phases=1, host_threads=8)
    out, att, _ = pkg.run_workload(wl, phases=1)
    assert np.array_equal(out.ActualTime, ref.ActualTime)
    assert rel_err(out.ActualState, ref.ActualState).max() < 1e-7
    assert (att == ratt).mean() >= 0.9

    wl.buffers.TimeDomain[1, :] = 40.0
    ref, _, _ = oracle.run_oracle(wl, phases=1, host_threads=8)
    out, _, _ = pkg.run_workload(wl, phases=1)
    assert np.isfinite(out.ActualState).all()
    assert np.median(rel_err(out.ActualState, ref.ActualState)) > 1e-4     # the instances did decorrelate: this is the chaotic regime
    zg, zc = np.sort(out.ActualState[2]), np.sort(ref.ActualState[2])
    assert 20.0 < zg.mean() < 27.0 and abs(zg.mean() - zc.mean()) < 0.6     # <z> on the attractor ~ 23.5
    assert abs(zg.std() - zc.std()) < 0.6
    # two-sample Kolmogorov-Smirnov distance between the z distributions of the two ensembles (critical value at
    # alpha = 1e-3 for n = m = 2048 is 1.95 sqrt(2/n) = 0.061)
    grid = np.concatenate([zg, zc])
    ks = np.abs(np.searchsorted(zg, grid, side="right") / NT - np.searchsorted(zc, grid, side="right") / NT).max()
    assert ks < 0.061
    # the trapping region of the attractor
    assert (np.abs(out.ActualState[0]) < 25).all() and (out.ActualState[2] > 0).all() and (out.ActualState[2] < 55).all()
