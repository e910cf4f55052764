"""The source-level drop-in: a user program written in the reference's style (its own system
definition with the six PerThread_* callbacks, then `#include "SingleSystem_PerThread_Interface.cuh"`,
then ProblemSolver<...> driven with SolverOption / SetHost / Synchronise* / Solve / GetHost) compiled
by nvcc against include/ and linked with the runtime library.  The system (forced Van der Pol, two
events, accessories, integer shared parameter, dense output, Finalization that moves the time domain)
is NOT one of the built-ins: the callbacks are inlined into a kernel instantiated in the user's
translation unit.  Checked against the CPU restatement of the reference's algorithm."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "examples", "vanderpol", "libvanderpol_dropin.so")


def run_example(algorithm, phases, mu, shape):
    lib = C.CDLL(LIB)
    NT = lib.vanderpol_instances()
    assert NT == mu.size
    out = dict(state=np.zeros((shape.SD, NT)), time=np.zeros(NT), acc=np.zeros((shape.NA, NT)), acci=np.zeros((shape.NIA, NT), np.int32),
               didx=np.zeros(NT, np.int32), dtime=np.zeros((shape.NDO, NT)), dstate=np.zeros((shape.NDO, shape.SD, NT)),
               att=np.zeros(NT, np.uint64), accd=np.zeros(NT, np.uint64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "/tmp")
    try:
        rc = lib.vanderpol_run(algorithm, phases, p(mu), p(out["state"]), p(out["time"]), p(out["acc"]), p(out["acci"]), p(out["didx"]),
                               p(out["dtime"]), p(out["dstate"]), p(out["att"]), p(out["accd"]))
        files = {f: open(f).read() for f in ("ActualStateInSolverObject.txt", "DenseOutput_7.txt") if os.path.exists(f)}
    finally:
        os.chdir(cwd)
    assert rc == 0
    return out, files


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)


@pytest.mark.parametrize("algorithm", [0, 1])
def test_user_defined_system_through_the_dropin_header(algorithm, workloads, oracle):
    if not os.path.exists(LIB):
        pytest.fail("examples/vanderpol/libvanderpol_dropin.so missing: run __graft_entry__.build()")
    wl = workloads.vanderpol_example(algorithm=algorithm)
    ref, ratt, racc = oracle.run_oracle(wl, phases=2)
    out, files = run_example(algorithm, 2, np.ascontiguousarray(wl.buffers.ControlParameters[0]), wl.shape)
    # event counts, user-defined termination after 3 crossings, accepted-step counter kept by the callback
    assert np.array_equal(out["acci"][0], ref.IntegerAccessories[0]) and (out["acci"][0] == 3).all()
    same_steps = (out["att"].astype(np.int64) == ratt)
    assert same_steps.mean() > 0.97 and np.abs(out["att"].astype(np.int64) - ratt).max() <= 6
    assert np.abs(out["acci"][1] - ref.IntegerAccessories[1]).max() <= 6
    tol = 1e-9 if algorithm == 0 else 1e-7
    assert rel_err(out["time"], ref.ActualTime).max() < tol          # event location (secant refinement to 1e-7)
    assert rel_err(out["state"], ref.ActualState).max() < 10 * tol
    assert rel_err(out["acc"], ref.Accessories).max() < 10 * tol
    assert np.array_equal(out["didx"], ref.DenseIndex)
    ok = np.where(same_steps)[0]
    n = int(ref.DenseIndex.min())
    # stored points sit on accepted steps.  RK4: the same steps, point by point.  RKCK45: two correct realisations of the same
    # controller may reject a step at different places and still take the same number of steps, so the bulk of the instances
    # (95 %) must agree point by point -- times to 10 tol, states to 1e-4 (the state moves with speed O(60) in the fast
    # segments) -- and the rest must stay on the same orbit (1e-2)
    et = rel_err(out["dtime"][:n][:, ok], ref.DenseTime[:n][:, ok]).max(axis=0)
    es = rel_err(out["dstate"][:n][:, :, ok], ref.DenseState[:n][:, :, ok]).max(axis=(0, 1))
    if algorithm == 0:
        assert et.max() < 10 * tol and es.max() < 1e-8
    else:
        assert (et < 10 * tol).mean() >= 0.95 and (es < 1e-4).mean() >= 0.95
        assert et.max() < 1e-3 and es.max() < 1e-2
    # Print(): the reference's file formats (instance per row, %18.10e, comma separated; dense output with headers)
    rows = files["ActualStateInSolverObject.txt"].rstrip("\n").split("\n")
    assert len(rows) == wl.NT and len(rows[0].split(",")) == 2 and len(rows[0].split(",")[0]) == 18
    assert abs(float(rows[5].split(",")[0]) - out["state"][0, 5]) < 1e-9 * max(1, abs(out["state"][0, 5]))
    dense = files["DenseOutput_7.txt"]
    for header in ("ControlParameters:", "SharedParameters:", "IntegerSharedParameters:", "Accessories:", "IntegerAccessories:", "Time series:"):
        assert header in dense
    series = dense.split("Time series:\n")[1].strip().split("\n")
    assert len(series) == int(out["didx"][7]) and len(series[0].split(",")) == 3


# ---- the coupled module through its drop-in header ----------------------------------------------------------------

SWING_LIB = os.path.join(ROOT, "examples", "swing_ring", "libswing_ring_dropin.so")


def test_user_defined_coupled_system_through_the_dropin_header(pkg, oracle):
    """examples/swing_ring: a user program against include/CoupledSystems_PerBlock_Interface.cuh (all six SetHost /
    GetHost overloads, banded wrap-around coupling matrices through SetHost(k, CouplingMatrix, row, col, v), two
    couplings, unit / system / global / integer-global scopes, events with user-defined termination, dense output,
    Finalization that moves the time domain, Print formats).  Its system is not built into the library.  Checked
    against the CPU restatement of the reference's coupled algorithm."""
    import importlib
    if not os.path.exists(SWING_LIB):
        pytest.fail("examples/swing_ring/libswing_ring_dropin.so missing: run __graft_entry__.build()")
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    wl = WC.swing_ring()
    sh, b = wl.shape, wl.buffers
    NS, UPS, UD, NDO = b.NS, sh.UPS, sh.UD, sh.NDO
    ref, ratt, racc = oracle.run_coupled_oracle(wl, phases=wl.phases, host_threads=8)

    lib = C.CDLL(SWING_LIB)
    shape = (C.c_int * 16)()
    lib.swing_ring_shape(shape)
    assert list(shape) == [NS, UPS, UD, sh.NC, sh.CBW, sh.CCI, sh.NUP, sh.NSP, sh.NGP, sh.NiGP, sh.NUA, sh.NiUA, sh.NSA, sh.NiSA, sh.NE, NDO]
    out = dict(state=np.zeros((UD, NS, UPS)), time=np.zeros(NS), td=np.zeros((2, NS)), uacc=np.zeros((NS, UPS)), iuacc=np.zeros((NS, UPS), np.int32),
               sacc=np.zeros(NS), isacc=np.zeros(NS, np.int32), didx=np.zeros(NS, np.int32), dtime=np.zeros((NDO, NS)),
               dstate=np.zeros((NDO, UD, NS, UPS)), att=np.zeros(NS, np.uint64), acc=np.zeros(NS, np.uint64))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    omega = np.ascontiguousarray(b.UnitParameters[0, 0]); damping = np.ascontiguousarray(b.UnitParameters[1, 0])
    gain = np.ascontiguousarray(b.CouplingStrength[0]); theta0 = np.ascontiguousarray(b.ActualState[0])
    matrix = np.ascontiguousarray(np.transpose(b.CouplingMatrix, (0, 2, 1)))     # [k][row][col]
    lib.swing_ring_run.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 12
    cwd = os.getcwd()
    os.chdir(os.path.join(ROOT, "gpurun_out") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "/tmp")
    try:
        rc = lib.swing_ring_run(wl.phases, p(omega), p(damping), p(gain), p(theta0), p(matrix), float(b.GlobalParameters[0]), int(b.IntegerGlobalParameters[0]),
                                float(b.TimeDomain[1, 0]), p(out["state"]), p(out["time"]), p(out["td"]), p(out["uacc"]), p(out["iuacc"]), p(out["sacc"]),
                                p(out["isacc"]), p(out["didx"]), p(out["dtime"]), p(out["dstate"]), p(out["att"]), p(out["acc"]))
        files = {f: open(f).read() for f in ("ActualStateInSolverObject.txt", "CouplingStrengthInSolverObject.txt", "CouplingMatrixInSolverObject_0.txt",
                                             "DenseOutput_5.txt") if os.path.exists(f)}
    finally:
        os.chdir(cwd)
    assert rc == 0
    assert np.array_equal(out["iuacc"], ref.IntegerUnitAccessories[0])                 # event counts of every unit
    assert (out["iuacc"][:, 0] == 3).any() and (out["iuacc"][:, 0] < 3).any()          # some phases end by the user-defined termination, others at TD[1]
    assert np.abs(out["att"].astype(np.int64) - ratt).max() <= 3 and np.abs(out["isacc"] - ref.IntegerSystemAccessories[0]).max() <= 3
    assert rel_err(out["time"], ref.ActualTime).max() < 1e-6                           # event-terminated phases (secant refinement to 1e-7)
    assert rel_err(out["td"], ref.TimeDomain).max() < 1e-6                             # moved by Finalization
    assert rel_err(out["sacc"], ref.SystemAccessories[0]).max() < 1e-6
    assert rel_err(out["state"], ref.ActualState).max() < 1e-5
    # max |v| over the ACCEPTED steps: a sampled maximum, it moves with the step sequence (median agreement 1e-13)
    assert rel_err(out["uacc"], ref.UnitAccessories[0]).max() < 1e-3 and np.median(rel_err(out["uacc"], ref.UnitAccessories[0])) < 1e-9
    same = out["didx"] == ref.DenseIndex
    assert same.mean() > 0.9
    n = int(min(out["didx"].min(), ref.DenseIndex.min())) - 1
    cols = np.where(same)[0]
    # stored points sit on accepted steps; after an event the step sequences agree to the event tolerance (1e-7 in v)
    assert rel_err(out["dtime"][:n][:, cols], ref.DenseTime[:n][:, cols]).max() < 1e-5
    assert rel_err(out["dstate"][:n][:, :, cols], ref.DenseState[:n][:, :, cols]).max() < 1e-3
    assert np.median(rel_err(out["dstate"][:n][:, :, cols], ref.DenseState[:n][:, :, cols])) < 1e-9
    # Print(): the reference's formats -- unit scope: one row per logical thread INCLUDING its per-block padding rows
    rows = files["ActualStateInSolverObject.txt"].rstrip("\n").split("\n")
    assert len(rows) == NS * 64 and rows[40].count("PADDING") == UD and len(rows[0].split(",")[0]) == 18
    assert abs(float(rows[64 * 3 + 7].split(",")[1]) - out["state"][1, 3, 7]) < 1e-9 * max(1.0, abs(out["state"][1, 3, 7]))
    rows = files["CouplingMatrixInSolverObject_0.txt"].rstrip("\n").split("\n")
    assert len(rows) == UPS and abs(float(rows[5].split(",")[6]) - matrix[0, 5, 6]) < 1e-12 and float(rows[5].split(",")[20]) == 0.0
    dense = files["DenseOutput_5.txt"]
    for header in ("UnitParameters:", "UnitAccessories:", "IntegerUnitAccessories:", "SystemParameters:", "SystemAccessories:",
                   "IntegerSystemAccessories:", "GlobalParameters:", "IntegerGlobalParameters:", "Time series:"):
        assert header in dense
    series = dense.split("Time series:\n")[1].strip().split("\n")
    assert len(series) == 1 + int(out["didx"][5]) and len(series[1].split(",")) == 1 + UPS * UD
