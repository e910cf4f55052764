"""One solver object over several devices (mpgos_perthread_create_sharded / mpgos_coupled_create_sharded): the tile map on the CPU,
and -- on the GPU -- that a sharded solve is the unsharded solve bit for bit, including for a system whose callbacks use the
instance id (`tidprobe`).  A device may appear several times in the device list, so the whole mechanism (strided copies, one launch
per part, gather) is exercised on a single GPU as well; with two GPUs present the parts really sit on different devices."""
import numpy as np
import pytest


def test_tile_map_is_a_partition_of_the_sweep(pkg):
    """mpgos_shard_instances / mpgos_shard_global_id (no device needed): every global id is owned by exactly one part, local ids are
    dense, ids increase with the local index, ragged last tiles included."""
    lib = pkg.load_library()
    for NT, tile, count in [(1 << 14, 4096, 2), (10000, 4096, 2), (10000, 32, 8), (4097, 4096, 2), (96, 32, 3), (1 << 20, 4096, 8), (12345, 64, 5)]:
        owners = np.full(NT, -1)
        total = 0
        for index in range(count):
            n = lib.mpgos_shard_instances(NT, tile, count, index)
            assert n >= 0
            total += n
            sample = np.unique(np.append(np.arange(0, n, max(1, n // 257)), max(n - 1, 0)))[:n]
            g = np.array([lib.mpgos_shard_global_id(tile, count, index, int(i)) for i in sample])
            assert (g < NT).all() and (g >= 0).all() and (np.diff(g) > 0).all()
            i = np.arange(n)
            gid = ((i // tile) * count + index) * tile + i % tile          # the formula of mpgos_b200.h
            assert (gid < NT).all()
            assert (owners[gid] == -1).all()
            owners[gid] = index
            if n:
                assert lib.mpgos_shard_global_id(tile, count, index, n - 1) == gid[-1] and lib.mpgos_shard_global_id(tile, count, index, 0) == gid[0]
        assert total == NT and (owners >= 0).all()
    assert lib.mpgos_shard_instances(100, 0, 2, 0) == -1 and lib.mpgos_shard_global_id(32, 2, 2, 0) == -1


def _devices(parts):
    import torch
    n = torch.cuda.device_count()
    return [i % n for i in range(parts)]


@pytest.mark.gpu
@pytest.mark.parametrize("NT,tile,parts", [(8192, 1024, 2), (5000, 512, 3), (4097, 4096, 2)])
def test_sharded_solver_shows_callbacks_the_global_instance_id(NT, tile, parts, pkg):
    """`tidprobe`: right-hand side, Initialization and Finalization use tid and NT.  Sharded == unsharded bit for bit, and the
    accessories hold exactly the ids of the whole sweep (ragged tiles and a sweep shorter than parts x tile included)."""
    wl = pkg.workloads.tidprobe(NT=NT)
    one, oatt, oacc = pkg.run_workload(wl, phases=2)
    many, matt, macc = pkg.run_workload(wl, phases=2, devices=_devices(parts), tile=tile)
    assert np.array_equal(one.IntegerAccessories[0], np.arange(NT)) and (one.IntegerAccessories[1] == NT).all()
    for f in ("ActualState", "ActualTime", "Accessories", "IntegerAccessories", "TimeDomain", "ControlParameters"):
        assert np.array_equal(getattr(one, f), getattr(many, f)), f
    assert np.array_equal(oatt, matt) and np.array_equal(oacc, macc)
    assert np.abs(many.ActualState[0] - np.arange(NT) / NT).max() < 0.8          # relaxing towards tid / NT


@pytest.mark.gpu
def test_sharded_solver_is_the_unsharded_solver_bit_for_bit(pkg):
    """Events, accessories, dense output, several phases, fewer active instances than allocated: Duffing (T1 feature set) on 3 parts."""
    wl = pkg.workloads.duffing_t1(NT=3000, algorithm=pkg.workloads.RKCK45, ndo=50)
    wl.options.ActiveNumberOfThreads = 2777
    one, oatt, oacc = pkg.run_workload(wl, phases=3)
    many, matt, macc = pkg.run_workload(wl, phases=3, devices=_devices(3), tile=256)
    for f in ("ActualState", "ActualTime", "Accessories", "IntegerAccessories", "DenseIndex", "DenseTime", "DenseState"):
        assert np.array_equal(getattr(one, f), getattr(many, f), equal_nan=True), f
    assert np.array_equal(oatt, matt) and np.array_equal(oacc, macc)
    assert (matt[2777:] == 0).all() and (matt[:2777] > 0).all()


@pytest.mark.gpu
def test_sharded_solver_on_two_devices(pkg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    wl = pkg.workloads.lorenz(NT=1 << 16)
    one, oatt, _ = pkg.run_workload(wl, phases=1)
    two, tatt, _ = pkg.run_workload(wl, phases=1, devices=[0, 1])
    assert np.array_equal(one.ActualState, two.ActualState) and np.array_equal(oatt, tatt)
    with pkg.ProblemSolver("lorenz", wl.NT, Devices=[0, 1]) as ps:
        lib = pkg.load_library()
        assert lib.mpgos_perthread_shard_count(ps._h) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ring_events_dense", "bubbles_tensor_core"])
def test_sharded_coupled_solver_is_the_unsharded_one_bit_for_bit(case, pkg):
    """Block-per-system module over 3 parts (ragged tiles): ring of Duffing oscillators with events, system accessories and dense
    output in two phases; 64-unit bubble ensembles through the tensor-core coupling."""
    import importlib
    WC = importlib.import_module(pkg.__name__ + ".workloads_coupled")
    if case == "ring_events_dense":
        wl, phases, kw, tile = WC.duffing_ring(NS=23, CBW=1, CCI=1), 2, {}, 4
    else:
        wl, phases, kw, tile = WC.bubble_ensemble(NS=37, UPS=64, amplitude=(0.2, 1.5)), 1, dict(tensor_core_coupling=1, dispatch_order=0), 8
    one, oatt, oacc = pkg.run_coupled_workload(wl, phases=phases, **kw)
    many, matt, macc = pkg.run_coupled_workload(wl, phases=phases, devices=_devices(3), tile=tile, **kw)
    for f in ("ActualState", "ActualTime", "UnitAccessories", "IntegerUnitAccessories", "SystemAccessories", "IntegerSystemAccessories", "DenseIndex", "DenseTime", "DenseState"):
        assert np.array_equal(getattr(one, f), getattr(many, f), equal_nan=True), f
    assert np.array_equal(oatt, matt) and np.array_equal(oacc, macc)
