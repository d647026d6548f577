"""Multi-GPU decomposition of the month.

Stages 1-2 are independent per granule, stage 3 is a sum over granules and
stage 4 is per cell after the sum.  So: granules are dealt to ranks by day
(round-robin, keeps every rank's share of orbits contiguous in time), each rank
runs K0..K4 into its LOCAL [10][n_cell] accumulator block, ONE all-reduce(sum)
merges the blocks (counts are exact integers held in float64, so one dtype and
one call suffice), and the OI stage is replicated.  No other collective exists
on the path.  The helpers below are backend-agnostic (NCCL on the GPUs, gloo in
the CPU test-suite).
"""
from __future__ import annotations

import datetime


def day_of(granule_time: datetime.datetime) -> int:
    return granule_time.toordinal()


def assign(times, rank: int, world: int):
    """Indices of the granules rank `rank` owns: day d goes to rank d % world
    (days numbered from the first day present)."""
    if world <= 1:
        return list(range(len(times)))
    days = sorted({day_of(t) for t in times})
    owner = {d: i % world for i, d in enumerate(days)}
    return [i for i, t in enumerate(times) if owner[day_of(t)] == rank]


def merge_accumulators(acc, group=None):
    """In-place sum of the accumulator block over the ranks of `group`."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return acc
    dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc
