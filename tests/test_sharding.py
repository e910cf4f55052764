"""Host-side logic of the multi-GPU path on CPU: tile-cyclic sharding partitions the sweep exactly,
and the cross-rank reduction (world_size 2, gloo) takes MAX of timings and SUM of step counts."""
import importlib
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKAGE = "massively-parallel-gpu-ode-solver_b200"


def test_tile_cyclic_shards_partition_the_sweep(pkg):
    sharding = importlib.import_module(PACKAGE + ".sharding")
    for world in (2, 4, 8):
        per_rank = 3 * sharding.TILE
        ids = np.concatenate([sharding.shard_indices(per_rank, r, world) for r in range(world)])
        assert np.array_equal(np.sort(ids), np.arange(per_rank * world))
    full = pkg.workloads.lorenz(NT=2 * 8192)
    parts = [sharding.shard_workload(full, 8192, r, 2, tile=256) for r in range(2)]
    rho = np.concatenate([p.buffers.ControlParameters[0] for p in parts])
    assert np.array_equal(np.sort(rho), full.buffers.ControlParameters[0])
    # balanced: both shards see the whole range of the sweep
    assert all(p.buffers.ControlParameters[0].min() < 0.7 and p.buffers.ControlParameters[0].max() > 20.3 for p in parts)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module(PACKAGE + ".sharding")
    W = importlib.import_module(PACKAGE + ".workloads")
    wl = sharding.shard_workload(W.lorenz(NT=world * 4096), 4096, rank, world, tile=512)
    maxima, sums = sharding.reduce_over_ranks([10.0 + rank, 1.0], [float(wl.buffers.ControlParameters.sum()), 1.0])
    q.put((rank, maxima, sums))
    dist.barrier()
    dist.destroy_process_group()


def test_reduction_over_two_ranks_gloo(pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total_rho = pkg.workloads.lorenz(NT=2 * 4096).buffers.ControlParameters.sum()
    for rank, maxima, sums in results:
        assert maxima == [11.0, 1.0]                        # MAX over ranks
        assert abs(sums[0] - total_rho) < 1e-6 and sums[1] == 2.0   # SUM over ranks; shards cover the sweep


def test_coupled_systems_shard_like_instances(pkg):
    """bench.py --workload coupled --gpus N: systems are dealt to the ranks in tiles of 64; unit-scope arrays keep their
    [serial, system, unit] shape, the coupling matrix and the coupling index (shared by all systems) are replicated."""
    sharding = importlib.import_module(PACKAGE + ".sharding")
    WC = pkg.workloads_coupled
    world, per_rank, tile = 2, 128, 64
    full = WC.bubble_ensemble(NS=world * per_rank, UPS=16)
    parts = [full.buffers.take(sharding.shard_indices(per_rank, r, world, tile)) for r in range(world)]
    for r, b in enumerate(parts):
        assert b.NS == per_rank and b.UnitParameters.shape == (21, per_rank, 16) and b.ActualState.shape == (2, per_rank, 16)
        assert np.array_equal(b.CouplingMatrix, full.buffers.CouplingMatrix) and np.array_equal(b.CouplingIndex, full.buffers.CouplingIndex)
        ids = sharding.shard_indices(per_rank, r, world, tile)
        assert np.array_equal(b.UnitParameters[5], full.buffers.UnitParameters[5][ids])          # the swept pressure amplitude
        assert np.array_equal(b.CouplingStrength, full.buffers.CouplingStrength[:, ids])
    amplitude = np.concatenate([b.UnitParameters[5][:, 0] for b in parts])
    assert np.array_equal(np.sort(amplitude), np.sort(full.buffers.UnitParameters[5][:, 0]))
    # both ranks see both ends of the sweep
    assert all(b.UnitParameters[5][:, 0].max() > 0.7 * full.buffers.UnitParameters[5][:, 0].max() for b in parts)
