"""Inputs over GPUs: the only multi-GPU logic this path has.

A ``device_t`` (one input and its channels) shares no mutable state with any other (boondock_airband.h:272-292; the
reference already runs one demod thread per device, boondock_airband.cpp:1088-1122), so ranks own disjoint sets of inputs
and nothing crosses between them on the data path.  ``torch.distributed`` is used for the start/stop barrier and for the
max-over-ranks time only.
"""
from __future__ import annotations

from typing import List, Sequence


def inputs_of_rank(n_inputs_total: int, world: int, rank: int) -> List[int]:
    """Static partition of a fixed set of inputs: input i runs on GPU i mod G (strong scaling, BASELINE cfg 3)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d of %d" % (rank, world))
    return list(range(rank, n_inputs_total, world))


def first_input_of_rank(inputs_per_gpu: int, rank: int) -> int:
    """Weak scaling (bench.py): every rank brings `inputs_per_gpu` inputs of its own; rank r owns [r*n, (r+1)*n)."""
    return rank * inputs_per_gpu


def max_over_ranks(seconds: float, dist=None, device=None) -> float:
    """The job takes as long as its slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(seconds)
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_msps(world: int, steps: int, samples_per_step_per_rank: int, seconds: float) -> float:
    """Whole-job throughput: what all ranks consumed over the slowest rank's time."""
    return world * steps * samples_per_step_per_rank / seconds / 1e6


def gather_shards(per_rank: Sequence[Sequence[int]]) -> List[int]:
    """Sorted union of the ranks' input lists; raises if two ranks claim the same input."""
    seen = {}
    for r, lst in enumerate(per_rank):
        for i in lst:
            if i in seen:
                raise ValueError("input %d on ranks %d and %d" % (i, seen[i], r))
            seen[i] = r
    return sorted(seen)
