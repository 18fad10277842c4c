"""Multi-GPU plumbing for batched branch-and-bound.

The LP hot path never shards: every rank solves whole LPs.  The only exchange between ranks is the
incumbent - an 8-byte ``all_reduce(MIN)`` of the incumbent objective per round and, when some rank
improved it, a broadcast of the incumbent vector from the winning rank so that the column reductions
of the reference's driver (/root/reference/src/sypha_solver_bnb_driver.cpp:906-929) stay consistent on
all replicas - plus max/sum reductions of the timing counters for the bench.  Works over NCCL
(device tensors) and over gloo (CPU tensors, used by the world_size-2 tests).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
import torch.distributed as dist


def _dev(group=None):
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def partition_round_robin(items: Sequence, rank: int, world: int):
    """Initial split of open nodes (or LP instances): item i goes to rank i % world."""
    return [it for i, it in enumerate(items) if i % world == rank]


def exchange_incumbent(obj: float, x: Optional[torch.Tensor], n: int, group=None):
    """All ranks call this once per round with their local incumbent (``obj = +inf``, ``x = None`` if
    they have none).  Returns (best_obj, best_x, owner_rank); ties go to the lowest rank so every rank
    takes the same decision."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = _dev(group)
    best = torch.tensor([obj if obj is not None else math.inf], dtype=torch.float64, device=dev)
    dist.all_reduce(best, op=dist.ReduceOp.MIN, group=group)
    best_obj = float(best.item())
    if not math.isfinite(best_obj):
        return math.inf, None, -1
    cand = torch.tensor([rank if (obj is not None and obj == best_obj) else world], dtype=torch.int64, device=dev)
    dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=group)
    owner = int(cand.item())
    buf = torch.zeros(n, dtype=torch.float64, device=dev)
    if rank == owner:
        buf.copy_(x.to(dev, torch.float64))
    dist.broadcast(buf, src=owner, group=group)
    return best_obj, buf, owner


def global_lower_bound(local_bound: float, group=None) -> float:
    """min over ranks of the smallest open-node dual bound (gap test, bnb_driver.cpp:713-723)."""
    t = torch.tensor([local_bound], dtype=torch.float64, device=_dev(group))
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return float(t.item())


def reduce_counters(elapsed_s: float, counts: Sequence[float], group=None):
    """Bench timing contract: elapsed = MAX over ranks, work counters = SUM over ranks."""
    dev = _dev(group)
    t = torch.tensor([elapsed_s], dtype=torch.float64, device=dev)
    c = torch.tensor(list(counts), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), [float(v) for v in c.tolist()]
