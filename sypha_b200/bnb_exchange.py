"""Multi-GPU plumbing for batched branch-and-bound.

The LP hot path never shards: every rank solves whole LPs.  The only exchange between ranks is the
incumbent - an 8-byte ``all_reduce(MIN)`` of the incumbent objective per round and, when some rank
improved it, a broadcast of the incumbent vector from the winning rank so that the column reductions
of the reference's driver (/root/reference/src/sypha_solver_bnb_driver.cpp:906-929) stay consistent on
all replicas - plus max/sum reductions of the timing counters for the bench.  Works over NCCL
(device tensors) and over gloo (CPU tensors, used by the world_size-2 tests).
"""
from __future__ import annotations

import collections
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


def plan_transfers(sizes: Sequence[int]):
    """Deterministic balancing plan from the frontier sizes of all ranks: returns (targets, moves) with
    ``moves`` a list of (src, dst, count).  Targets differ by at most one node; ranks are matched in rank
    order, so every rank derives the same plan from the same ``all_gather``."""
    world = len(sizes)
    total = int(sum(sizes))
    targets = [total // world + (1 if r < total % world else 0) for r in range(world)]
    surplus = [[r, int(sizes[r]) - targets[r]] for r in range(world) if sizes[r] > targets[r]]
    deficit = [[r, targets[r] - int(sizes[r])] for r in range(world) if sizes[r] < targets[r]]
    moves = []
    i = j = 0
    while i < len(surplus) and j < len(deficit):
        k = min(surplus[i][1], deficit[j][1])
        moves.append((surplus[i][0], deficit[j][0], k))
        surplus[i][1] -= k
        deficit[j][1] -= k
        if surplus[i][1] == 0:
            i += 1
        if deficit[j][1] == 0:
            j += 1
    return targets, moves


def rebalance_frontier(nodes: list, max_depth: int, min_imbalance: int = 1, group=None):
    """Node donation between ranks (frontier sizes gathered, nodes moved from the long
    frontiers to the short ones).  A node is just its decision list and its parent bound
    (/root/reference/src/sypha_solver_heuristics.h:9-30), so it travels as ``max_depth + 2`` 8-byte words:
    [n_decisions, bits of the bound, var * 2 + fix ...].

    ``nodes``: this rank's open nodes as (decisions, bound) pairs in processing order; donated nodes are
    taken from the END (processed last) and appended at the end of the receiver's list.  Nodes deeper than
    ``max_depth`` stay where they are.  All ranks must call this in the same round.  Returns
    (new_nodes, global_open_count, n_sent, n_received)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = _dev(group)
    mine = torch.tensor([len(nodes)], dtype=torch.int64, device=dev)
    sizes_t = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes_t, mine, group=group)
    sizes = [int(t.item()) for t in sizes_t]
    total = sum(sizes)
    if world == 1 or max(sizes) - min(sizes) <= min_imbalance:
        return nodes, total, 0, 0
    _, moves = plan_transfers(sizes)
    cap = max(sum(k for s, _, k in moves if s == r) for r in range(world))     # rows every rank contributes
    width = max_depth + 3
    out = torch.zeros((cap, width), dtype=torch.int64)
    out[:, 0] = -1                                                              # destination; -1 = empty row
    keep = list(nodes)
    sent = 0
    for src, dst, k in moves:
        if src != rank:
            continue
        for _ in range(k):
            # last donatable node (depth <= max_depth)
            idx = next((i for i in range(len(keep) - 1, -1, -1) if len(keep[i][0]) <= max_depth), None)
            if idx is None:
                break
            dec, bound = keep.pop(idx)
            row = out[sent]
            row[0] = dst
            row[1] = len(dec)
            row[2] = torch.tensor([bound], dtype=torch.float64).view(torch.int64)[0]
            if dec:
                row[3:3 + len(dec)] = torch.tensor([2 * int(v) + int(f) for v, f in dec], dtype=torch.int64)
            sent += 1
    out = out.to(dev)
    gathered = [torch.zeros_like(out) for _ in range(world)]
    dist.all_gather(gathered, out, group=group)
    received = 0
    for r in range(world):
        if r == rank:
            continue
        rows = gathered[r].cpu()
        for row in rows[rows[:, 0] == rank]:
            d = int(row[1])
            bound = float(row[2:3].view(torch.float64)[0])
            dec = tuple((int(w) // 2, int(w) % 2) for w in row[3:3 + d].tolist())
            keep.append((dec, bound))
            received += 1
    return keep, total, sent, received


class AsyncBoundExchange:
    """Incumbent bounds between ranks WITHOUT a per-round barrier (BASELINE.json north_star item 4: NCCL only
    carries incumbent bounds and pruning information).

    Every round a rank posts one non-blocking ``all_gather`` of three doubles - its incumbent objective, its
    number of open nodes and its processed-node count - and collects the gather it posted ``lag`` rounds earlier,
    which by then has normally completed: the host never waits on the slowest rank's current round, ranks may
    drift up to ``lag`` rounds apart, and because every rank reads the SAME gathered rows for round r they all
    take the same decisions from them (adopt the best bound, rebalance the frontiers, stop).  The incumbent
    VECTOR does not travel during the search - only its 8-byte objective prunes - and is fetched once at the end
    from the rank that owns it (``final_incumbent``).

    Over NCCL the gather runs on NCCL's own stream and its result reaches the host through a pinned buffer; over
    gloo (CPU tests) the tensors are host tensors."""

    WORDS = 3

    def __init__(self, lag: int = 2, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.dev = _dev(group)
        self.lag = max(0, int(lag))
        self.pending = collections.deque()
        self.posted = 0
        self.bytes_posted = 0
        self.wait_s = 0.0          # host time spent blocked in collect(): the rank's idle time due to the exchange

    def post(self, incumbent: float, open_nodes: int, processed: int):
        src = torch.tensor([incumbent, float(open_nodes), float(processed)], dtype=torch.float64).to(self.dev)
        out = torch.empty(self.world * self.WORDS, dtype=torch.float64, device=self.dev)
        work = dist.all_gather_into_tensor(out, src, group=self.group, async_op=True)
        self.pending.append((work, out, src))
        self.posted += 1
        self.bytes_posted += 8 * self.WORDS

    def collect(self, drain: bool = False):
        """-> list of gathered [world, 3] tensors (host) that are due: everything older than ``lag`` rounds,
        or everything posted when ``drain``."""
        import time as _t
        due = []
        while self.pending and (drain or len(self.pending) > self.lag):
            work, out, _src = self.pending.popleft()
            t0 = _t.perf_counter()
            work.wait()
            rows = out.cpu().view(self.world, self.WORDS)
            self.wait_s += _t.perf_counter() - t0
            due.append(rows)
        return due

    def final_incumbent(self, obj: float, x, n: int):
        """End of the search: the best objective over all ranks and its vector, from the owning rank."""
        xt = None if x is None else torch.as_tensor(x, dtype=torch.float64)
        best, bx, owner = exchange_incumbent(obj, xt, n, group=self.group)
        return best, (None if bx is None else bx.cpu().numpy()), owner
