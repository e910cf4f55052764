"""Sharding of a parameter sweep over one solver object per GPU (one process per GPU).

Instances are independent, so there is no collective on the data path: each rank integrates its
own shard and the only cross-rank traffic is the reduction of step counts / timings for reporting
(``torch.distributed``: NCCL on GPUs, gloo in the CPU tests).  The reference's only multi-GPU
mechanism is one ProblemSolver per device driven from one host thread
(``Tutorials_SingleSystem_PerThread/T8_MultiGPUSingleNode/MultiGPUSingleNode.cu:55-56, 85-135``).

The shard is tile-cyclic rather than one contiguous slab per GPU because cost varies smoothly along a
sweep (Keller-Miksis: 28x between the two ends), which would leave most GPUs idle behind
the most expensive slab."""
from __future__ import annotations

import copy
from typing import Sequence

import numpy as np

TILE = 4096


def shard_indices(instances_per_rank: int, rank: int, world: int, tile: int = TILE) -> np.ndarray:
    """Global instance ids owned by ``rank``: tiles of ``tile`` consecutive instances dealt round-robin.
    Every rank owns whole tiles: ``instances_per_rank`` must be a multiple of ``tile`` (callers fall back to tile = 1)."""
    if instances_per_rank % tile != 0:
        raise ValueError(f"shard_indices: {instances_per_rank} instances per rank is not a multiple of the tile ({tile})")
    i = np.arange(instances_per_rank, dtype=np.int64)
    return (i // tile) * (world * tile) + rank * tile + (i % tile)


def shard_workload(full, instances_per_rank: int, rank: int, world: int, tile: int = TILE):
    """This rank's part of a workload generated for ``instances_per_rank * world`` instances."""
    if world == 1:
        return full
    assert full.NT == instances_per_rank * world
    wl = copy.copy(full)
    wl.buffers = full.buffers.take(shard_indices(instances_per_rank, rank, world, tile))
    return wl


def reduce_over_ranks(maxima: Sequence[float], sums: Sequence[float], device=None):
    """(max over ranks of each entry of ``maxima``, sum over ranks of each entry of ``sums``).
    Timings are reduced with MAX (the job is as slow as its slowest rank), work counts with SUM."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(maxima), list(sums)
    m = torch.tensor(list(maxima), dtype=torch.float64, device=device)
    s = torch.tensor(list(sums), dtype=torch.float64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return m.tolist(), s.tolist()
