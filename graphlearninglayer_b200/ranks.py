"""Host-side logic for running the layer data-parallel: one process per GPU, every rank owns independent graphs
(the layer has no parameters, so there is NO collective in the data path).  The only
communication is what a benchmark or trainer needs around the layer: a barrier and a max/sum of scalars.

Kept free of CUDA so that the multi-rank plumbing is testable on CPU with the gloo backend.
"""
from __future__ import annotations

import os
from typing import Optional


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))


def rank_seed(base_seed: int, rank: int) -> int:
    """Every rank draws a different graph: seeds are disjoint across ranks for any base seed < 10**6."""
    return base_seed + 1_000_003 * rank


def units_for_rank(n_units: int, rank: int, world: int):
    """Contiguous block of independent units (graphs) for this rank; sizes differ by at most one."""
    lo = n_units * rank // world
    hi = n_units * (rank + 1) // world
    return range(lo, hi)


def aggregate_throughput(units_this_rank: int, ms_this_rank: float, group=None, device: Optional[str] = None):
    """Whole-job throughput: (units processed by all ranks) / (max over ranks of the device time).

    Returns (units_total, ms_max, units_per_second).  Uses all_reduce(SUM) and all_reduce(MAX) on a 2-element tensor
    pair; with no process group it degenerates to the single-rank numbers.
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return units_this_rank, ms_this_rank, units_this_rank / (ms_this_rank * 1e-3)
    dev = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    s = torch.tensor([float(units_this_rank)], dtype=torch.float64, device=dev)
    t = torch.tensor([float(ms_this_rank)], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(round(s.item())), float(t.item()), s.item() / (t.item() * 1e-3)
