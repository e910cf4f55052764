"""Asynchronous orchestration semantics of the host object: every copy and launch is asynchronous on the
object's own stream, InsertSynchronisationPoint / SynchroniseSolver are a per-object event, so one host thread can drive
several solver objects -- double buffering on one GPU as Tutorials_SingleSystem_PerThread/T7_DoubleBuffering/
DoubleBuffering.cu:89-144 does, and one object per device as T8_MultiGPUSingleNode/MultiGPUSingleNode.cu:85-135 does."""
import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _halves(W, n=8192):
    full = W.lorenz(NT=n, algorithm=W.RKCK45)
    parts = []
    for idx in (np.arange(0, n // 2), np.arange(n // 2, n)):
        w = copy.copy(full)
        w.buffers = full.buffers.take(idx)
        parts.append(w)
    return full, parts


def _interleaved(pkg, parts, devices, rounds=3):
    """The tutorial loop: launch both objects back to back, then -- object by object -- wait for ITS event, read its results
    and refill its host mirrors for the next round while the other object may still be running."""
    V = pkg.Variables
    solvers = []
    for w, dev in zip(parts, devices):
        sh = w.shape
        ps = pkg.ProblemSolver(w.name, w.NT, NDO=sh.NDO, Algorithm=sh.Algorithm, Precision=sh.dtype, Device=dev)
        pkg.apply_workload(ps, w)
        solvers.append(ps)
    results = [[] for _ in solvers]
    for r in range(rounds):
        for ps in solvers:
            ps.SynchroniseFromHostToDevice(V.All)
            ps.Solve()
            ps.SynchroniseFromDeviceToHost(V.ActualState)
            ps.SynchroniseFromDeviceToHost(V.ActualTime)
            ps.InsertSynchronisationPoint()
        for i, (ps, w) in enumerate(zip(solvers, parts)):
            ps.SynchroniseSolver()
            results[i].append((ps.host(V.ActualState).copy(), ps.host(V.ActualTime).copy()))
            # next round: new control parameter, tutorial initial condition again
            ps.host(V.ControlParameters)[...] = w.buffers.ControlParameters + 0.25 * (r + 1)
            ps.host(V.ActualState)[...] = w.buffers.ActualState
            ps.host(V.ActualTime)[...] = w.buffers.ActualTime
            ps.host(V.TimeDomain)[...] = w.buffers.TimeDomain
    for ps in solvers:
        ps.close()
    return results


def _sequential(pkg, full, rounds=3):
    out = []
    for r in range(rounds):
        w = copy.copy(full)
        w.buffers = full.buffers.copy()
        w.buffers.ControlParameters[...] = full.buffers.ControlParameters + 0.25 * r
        res, _, _ = pkg.run_workload(w, phases=1)
        out.append((res.ActualState, res.ActualTime))
    return out


def test_double_buffering_two_solver_objects_on_one_gpu(pkg, workloads):
    full, parts = _halves(workloads)
    inter = _interleaved(pkg, parts, devices=(0, 0))
    seq = _sequential(pkg, full)
    h = full.NT // 2
    for r, (state, time) in enumerate(seq):
        assert np.array_equal(inter[0][r][0], state[:, :h]) and np.array_equal(inter[1][r][0], state[:, h:])
        assert np.array_equal(inter[0][r][1], time[:h]) and np.array_equal(inter[1][r][1], time[h:])


def test_one_solver_object_per_device_from_one_host_thread(pkg, workloads):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    full, parts = _halves(workloads)
    inter = _interleaved(pkg, parts, devices=(0, 1))
    seq = _sequential(pkg, full)
    h = full.NT // 2
    for r, (state, time) in enumerate(seq):
        assert np.array_equal(inter[0][r][0], state[:, :h]) and np.array_equal(inter[1][r][0], state[:, h:])
