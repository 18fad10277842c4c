"""Batched branch-and-bound node loop on top of the LP hot path.

The reference's driver (/root/reference/src/sypha_solver_bnb_driver.cpp:698-1046) pops ONE node at a time
from a FIFO frontier, rebuilds its model on the host (``build_branch_model``,
/root/reference/src/sypha_solver_bnb.cpp:418-490), re-uploads it and solves its LP.  Its search logic is
out of scope here; what this module restructures is the node BODY: a window of K open nodes is popped
per round and their LP relaxations are solved concurrently on one GPU (one persistent workspace and
stream per slot, ``sb200_solve_batch``), and with several ranks every rank works on its own part of the
frontier and only the incumbent travels (``bnb_exchange``).  The combinatorial parts kept on the host
are deliberately the plain versions of the reference's: breadth-first order, most-fractional branching
(sypha_solver_heuristics.cpp ``MostFractional``), nearest-integer rounding with greedy repair for
incumbents, pruning by the parent / node dual bound with integer costs.
"""
from __future__ import annotations

import collections
import itertools
import dataclasses
import math
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .instances import ScpModel
from .solver import (CODE_SUCCESSFUL, IpmWorkspace, SolverExecutionConfig, SolverGapStagnationConfig,
                     SyphaEnvironment, SyphaNodeSparse, get_cover, get_primal, get_rounded, initializeIpmWorkspace,
                     last_window, node_heuristics, releaseIpmWorkspace, window_begin, window_finish, set_heuristic_rules, solve_batch, solve_batch_nodes,
                     workspace_for_nodes)

TERM_CONVERGED, TERM_MAX_ITER, TERM_GAP_STALLED, TERM_NUMERICAL = 0, 1, 2, 3


def build_branch_model(base: ScpModel, decisions: Sequence[Tuple[int, int]]) -> ScpModel:
    """Node model = base + one row per branching decision (bnb.cpp:453-468): row = (fix == 0 ? -1 : +1) at
    ``var`` and -1 at a fresh slack column, rhs = fix, slack cost 0."""
    k = len(decisions)
    if k == 0:
        return base
    nnz = base.nnz
    offs = np.concatenate([base.offs.astype(np.int64), base.offs[-1] + 2 * np.arange(1, k + 1, dtype=np.int64)])
    inds = np.empty(nnz + 2 * k, dtype=np.int32)
    vals = np.empty(nnz + 2 * k, dtype=np.float64)
    inds[:nnz], vals[:nnz] = base.inds, base.vals
    var = np.fromiter((d[0] for d in decisions), dtype=np.int32, count=k)
    fix = np.fromiter((d[1] for d in decisions), dtype=np.float64, count=k)
    inds[nnz::2], vals[nnz::2] = var, np.where(fix == 0.0, -1.0, 1.0)
    inds[nnz + 1::2], vals[nnz + 1::2] = base.n + np.arange(k, dtype=np.int32), -1.0
    return ScpModel(base.m + k, base.n + k, base.n_orig, offs.astype(np.int32), inds, vals,
                    np.concatenate([base.c, np.zeros(k)]), np.concatenate([base.b, fix]), base.name + f"+{k}br")


def greedy_cover(base: ScpModel) -> Tuple[float, Optional[np.ndarray]]:
    """The reference's first incumbent (greedy_set_cover_heuristic, sypha_preprocessor.cpp:11-96): columns sorted
    by (cost ascending, rows covered descending) - full ties here by column index, the reference leaves them to
    std::sort (unstable, so implementation-defined) - and scanned once; a column is taken when it covers a row that is
    still uncovered.  Against the reference's own function (compiled in place for tests/test_bnb_host.py): identical
    where no two columns tie, the same objective on the bench's instances (scpnre1 38, scpnrg1 266); on instances with
    many full ties (unit costs) the two covers differ.  The C++ node loop calls the reference's function itself."""
    m, n0 = base.m, base.n_orig
    mask = (base.inds < n0) & (base.vals > 0.0)
    rows = np.repeat(np.arange(m), np.diff(base.offs))[mask]
    cols = base.inds[mask]
    order = np.argsort(cols, kind="stable")
    col_rows, col_ptr = rows[order], np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n0))])
    cnt = np.diff(col_ptr)
    covered = np.zeros(m, dtype=bool)
    left, total, x = m, 0.0, np.zeros(n0)
    for j in np.lexsort((np.arange(n0), -cnt, base.c[:n0])):
        if left <= 0:
            break
        r = col_rows[col_ptr[j]:col_ptr[j + 1]]
        new = r[~covered[r]]
        if len(new):
            covered[new] = True
            left -= len(new)
            total += float(base.c[j])
            x[j] = 1.0
    return (total, x) if left == 0 else (math.inf, None)


def reduce_by_incumbent(base: ScpModel, incumbent: float, tol: float = 1e-12) -> Tuple[ScpModel, np.ndarray]:
    """Drop every original column whose cost alone reaches the incumbent (SyphaNodeSparse::reduceByIncumbent,
    sypha_node_sparse.cpp:284-332, called at bnb_driver.cpp:297-306): the node LPs of configs[4] are solved on
    this model, not on the file's (scpnre1: 5000 -> 1775 columns).  Returns (reduced model, kept -> original
    column index)."""
    n0, m = base.n_orig, base.m
    keep = np.nonzero(base.c[:n0] + tol < incumbent)[0]
    if len(keep) == n0 or len(keep) == 0 or not math.isfinite(incumbent):
        return base, np.arange(n0)
    new_of = np.full(base.n, -1, dtype=np.int64)
    new_of[keep] = np.arange(len(keep))
    new_of[n0:] = len(keep) + np.arange(base.n - n0)
    ent_keep = new_of[base.inds] >= 0
    row_of = np.repeat(np.arange(m), np.diff(base.offs))
    offs = np.concatenate([[0], np.cumsum(np.bincount(row_of[ent_keep], minlength=m))]).astype(np.int32)
    red = ScpModel(m, len(keep) + (base.n - n0), len(keep), offs, new_of[base.inds[ent_keep]].astype(np.int32),
                   base.vals[ent_keep].copy(), np.concatenate([base.c[keep], base.c[n0:]]), base.b.copy(),
                   base.name + f"[{len(keep)} of {n0} columns]")
    return red, keep


class CoverHeuristic:
    """Nearest-integer rounding of the LP point, greedy repair of uncovered rows by cost per newly covered
    row (gains kept up to date incrementally: O(nnz) per call), then removal of redundant columns (most
    expensive first).  Host-side NumPy on the CSR / CSC lists of A0.  The product path runs the same rules on
    the device (``sb200_node_heuristics``, csrc/sb200_heur.cu); this class is what the tests check that kernel
    against, and what ``device_heuristics=False`` (the reference's host-side arrangement) uses."""

    def __init__(self, base: ScpModel):
        import scipy.sparse as sp
        m, n0 = base.m, base.n_orig
        A = sp.csr_matrix((base.vals, base.inds, base.offs), shape=(m, base.n))[:, :n0]
        self.A = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=(m, n0))
        At = self.A.T.tocsr()
        self.col_ptr, self.col_rows = At.indptr, At.indices
        self.row_ptr, self.row_cols = self.A.indptr, self.A.indices
        self.c = base.c[:n0]
        self.m, self.n0 = m, n0

    def _rows(self, j):
        return self.col_rows[self.col_ptr[j]:self.col_ptr[j + 1]]

    def __call__(self, x_lp: np.ndarray, fixed_zero=()) -> Tuple[float, Optional[np.ndarray]]:
        x = (x_lp[:self.n0] >= 0.5).astype(np.float64)
        banned = np.zeros(self.n0, dtype=bool)
        if len(fixed_zero):
            banned[list(fixed_zero)] = True
            x[banned] = 0.0
        cover = self.A @ x
        unc = cover < 0.5
        if unc.any():
            gain = self.A.T @ unc.astype(np.float64)           # rows each column would newly cover
            usable = ~banned & (x == 0.0)
            while unc.any():
                score = np.where(usable & (gain > 0.5), self.c / np.maximum(gain, 0.5), np.inf)
                j = int(np.argmin(score))
                if not np.isfinite(score[j]):
                    return math.inf, None                      # infeasible under the fixings
                x[j] = 1.0
                usable[j] = False
                rows = self._rows(j)
                cover[rows] += 1.0
                new = rows[unc[rows]]
                unc[new] = False
                if len(new):                                   # those rows no longer count for anybody's gain
                    touched = np.concatenate([self.row_cols[self.row_ptr[i]:self.row_ptr[i + 1]] for i in new])
                    gain -= np.bincount(touched, minlength=self.n0)
        chosen = np.nonzero(x)[0]
        for j in chosen[np.argsort(-self.c[chosen], kind="stable")]:   # drop redundant columns, dearest first
            rows = self._rows(j)
            if np.all(cover[rows] >= 1.5):
                x[j] = 0.0
                cover[rows] -= 1.0
        return float(self.c @ x), x


@dataclasses.dataclass
class BnbNode:
    decisions: Tuple[Tuple[int, int], ...]
    parent_bound: float
    warm: Optional[tuple] = None        # (device tensor x | y | s of the parent's LP, n, m): the child's warm start


@dataclasses.dataclass
class BnbStats:
    processed: int = 0
    lp_iterations: int = 0
    pruned_by_bound: int = 0
    infeasible: int = 0
    integral: int = 0
    rounds: int = 0
    incumbent: float = math.inf
    greedy_incumbent: float = math.inf
    lp_device_ms: float = 0.0
    kernels_launched: int = 0
    delta_rows: int = 0
    wall_s: float = 0.0
    open_nodes: int = 0
    root_bound: float = -math.inf
    nodes_sent: int = 0
    nodes_received: int = 0
    maxiter_nodes: int = 0          # node LPs that stopped at the iteration cap
    gap_stalled_nodes: int = 0      # node LPs that left through the gap-stagnation exit (kept, parent's bound)
    exchange_wait_ms: float = 0.0   # host time blocked in inter-rank collectives (this rank's idle time)
    round_max_iterations: int = 0   # sum over rounds of the longest LP of the window (what a round waits for)
    round_ms: list = dataclasses.field(default_factory=list)
    # where a round's wall time goes (ms): node deltas + the window's launch + its wait | of which the window's kernel on the
    # device (CUDA events) | the node-heuristics launch + read-back | host bookkeeping (bounds, branching, frontier)
    solve_ms: list = dataclasses.field(default_factory=list)
    window_ms: list = dataclasses.field(default_factory=list)
    heur_ms: list = dataclasses.field(default_factory=list)


class BatchedBnb:
    """K LP slots on one GPU.  ``run(max_nodes)`` processes the frontier in windows of K nodes."""

    def __init__(self, base: ScpModel, slots: int = 8, device: int = 0, max_iter: int = 100,
                 exchange=None, integer_costs: bool = True, device_nodes: bool = True, max_depth: int = 64,
                 heuristic_threads: int = 0, device_heuristics: bool = True, rebalance=None, rebalance_every: int = 1,
                 share_gpu: bool = True, poll_every: int = 1, node_lp: str = "reference", async_exchange=None,
                 rebalance_min_imbalance: Optional[int] = None, heuristic_rules: str = "reference",
                 branch_rule: str = "most_fractional", warm_start: bool = False, warm_floor: float = 0.1,
                 pipeline: int = 1):
        self.base = base
        self.device_nodes = device_nodes      # False: the reference's way (host CSR per node + full upload)
        self.max_depth = max_depth
        # branching variable + rounding/repair incumbent per node on the device, behind the node's LP on its own
        # stream (sb200_node_heuristics); False: the host-side NumPy versions on a host copy of x
        self.device_heuristics = device_heuristics and device_nodes
        # poll_every > 1: the host enqueues that many iterations per LP before it reads the scalar block back
        # (kernels of an LP that has finished return at once), halving the host work per iteration of a window
        self.env = SyphaEnvironment(cudaDeviceId=device, pollEvery=max(1, poll_every))
        # node LP configuration.  "reference": what the reference's driver passes for every node
        # (bnb_driver.cpp:833-837): gap-stagnation early exit on, window kBnbGapStallBranchIters = 5, minimum
        # improvement kBnbGapStallMinImprovPct = 1 % (sypha_environment_defaults.h:34-35).  "converged": every node
        # LP runs to mu <= mu_tol.  Either way only a CONVERGED LP with dual <= primal bounds its node
        # (boundIsReliableForPruning, bnb_driver.cpp:866-873); any other successful LP keeps its parent's bound
        # and is still branched on.
        if node_lp not in ("reference", "converged"):
            raise ValueError("node_lp must be 'reference' or 'converged'")
        self.node_lp = node_lp
        gs = SolverGapStagnationConfig(True, 5, 1.0) if node_lp == "reference" else SolverGapStagnationConfig(False, 0, 0.0)
        self.cfg = SolverExecutionConfig(maxIterations=max_iter, gapStagnation=gs)
        self.slots = slots
        # pipeline = 2: two sets of `slots` workspaces used alternately (sb200_window_begin / _finish): while the host
        # reads one window's results, branches and stages the next deltas, the other window's thread blocks are running and
        # the first set's next window is already queued behind it - no idle GPU between windows, and a window's late LPs
        # share the GPU with the next window's early ones.  Needs the device node path with the device rules.
        self.pipeline = max(1, pipeline) if (device_nodes and device_heuristics and not warm_start and slots > 1) else 1
        self.ws: List[IpmWorkspace] = []
        self.base_node = SyphaNodeSparse.from_csr(base.m, base.n, base.n_orig, base.offs, base.inds, base.vals,
                                                  base.c, base.b, self.env)
        for _ in range(slots * self.pipeline):
            if device_nodes:
                self.ws.append(workspace_for_nodes(self.base_node, max_depth, device))   # base model resident
            else:
                w = IpmWorkspace()
                initializeIpmWorkspace(w, device=device)
                self.ws.append(w)
        # per-node rules on the device: the reference's own (NearestIntegerFixing, then DualGuidedCoverRepair with the
        # node's duals; bnb_driver.cpp:879-905 tries them in that order and takes the first that improves) or the
        # plain rounding / greedy repair of round 1 ("plain", also what device_heuristics=False runs on the host)
        # children start from their parent's final iterate floored at `warm_floor` instead of the Mehrotra starting
        # point (sb200_node_delta.warm_start): fewer iterations per node LP and no starting-point
        # factorisation.  Needs the device-resident node path and the throughput form (several slots).
        self.warm_start = warm_start and self.device_nodes
        self.warm_floor = warm_floor
        self._export: List = [None] * slots
        self._pool_free: List = []
        self._arena = None
        self.warm_pool = 4096               # export buffers kept for open nodes (a node without one starts its children cold)
        self.heuristic_rules = heuristic_rules if self.device_heuristics else "plain"
        if self.device_heuristics:
            for w in self.ws:
                set_heuristic_rules(w, self.heuristic_rules, branch_rule, 1e-6)
        if share_gpu and slots > 1:              # K LPs in flight: fewer CTAs per factorisation (throughput over latency)
            from . import _lib as L
            for w in self.ws:
                L.load().sb200_set_concurrency_hint(w.handle, slots)
        self._sets = [self.ws[i * slots:(i + 1) * slots] for i in range(self.pipeline)]
        self._free_sets = list(range(self.pipeline))
        self._inflight: collections.deque = collections.deque()      # (batch, NodeWindow, set index, begin time)
        self._cur_ws = self.ws                                       # the set whose results are being processed
        self.heur = CoverHeuristic(base)
        self.frontier: collections.deque = collections.deque([BnbNode((), -math.inf)])
        self.incumbent = math.inf
        self.incumbent_x: Optional[np.ndarray] = None
        self.integer_costs = integer_costs
        self.exchange = exchange            # callable(obj, x) -> (obj, x) across ranks, or None
        # node donation between ranks (bnb_exchange.rebalance_frontier): callable(list of (decisions, bound)) ->
        # (list, open nodes over all ranks, sent, received); every rank calls it in the same rounds
        self.rebalance = rebalance
        self.rebalance_every = max(1, rebalance_every)
        self.global_open: Optional[int] = None
        # bnb_exchange.AsyncBoundExchange: objective-only exchange without a per-round barrier; replaces
        # `exchange` (blocking, carries the vector every round) when given
        self.async_exchange = async_exchange
        self.rebalance_min_imbalance = slots // 2 if rebalance_min_imbalance is None else rebalance_min_imbalance
        self.global_processed: Optional[int] = None
        self.stats = BnbStats()
        # optional: incumbent heuristics of round r on host threads WHILE the GPU solves round r+1 (the ctypes
        # call releases the GIL); results are folded in at a fixed point (after that solve), so the search stays
        # deterministic.  Off by default: measured slower on the B200 box (16 slots: 379 vs 411 nodes/s - the
        # threads delay the host loop that feeds the K LP streams)
        import concurrent.futures
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=heuristic_threads) if heuristic_threads > 0 else None
        self._pending = []

    @classmethod
    def with_reference_presolve(cls, model: ScpModel, **kw) -> "BatchedBnb":
        """The prelude of the reference's driver that shapes its node LPs (bnb_driver.cpp:262-306): greedy cover
        as the first incumbent, then every column whose cost reaches it is dropped.  The search runs on the
        reduced model; ``incumbent_in_input_space()`` maps the answer back."""
        obj, x = greedy_cover(model)
        red, keep = reduce_by_incumbent(model, obj)
        drv = cls(red, **kw)
        drv.input_model, drv.kept_cols = model, keep
        if x is not None:
            drv.incumbent, drv.incumbent_x = obj, x[keep]        # the greedy cover only uses columns cheaper than itself
            drv.stats.greedy_incumbent = obj
        return drv

    def incumbent_in_input_space(self) -> Optional[np.ndarray]:
        if self.incumbent_x is None:
            return None
        keep = getattr(self, "kept_cols", None)
        if keep is None:
            return self.incumbent_x
        x = np.zeros(self.input_model.n_orig)
        x[keep] = self.incumbent_x
        return x

    def _fold_pending(self):
        for fut in self._pending:
            self._offer(*fut.result())
        self._pending = []

    def drain(self):
        """Finish and process the windows still in flight (pipeline > 1)."""
        while self._inflight:
            batch, w, si = self._inflight.popleft()
            results, heur = window_finish(w)
            self._cur_ws = self._sets[si]
            self._process_batch(batch, results, heur)
            self._free_sets.append(si)

    def _open_nodes(self) -> int:
        return len(self.frontier) + sum(len(b) for b, _, _ in self._inflight)

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
        try:
            self.drain()
        except Exception:
            pass
        for w in self.ws:
            releaseIpmWorkspace(w)
        self.ws = []

    # a node whose bound cannot beat the incumbent is dropped (integer costs: bound rounds up)
    def _prunable(self, bound: float) -> bool:
        if not math.isfinite(self.incumbent) or not math.isfinite(bound):
            return bound == math.inf
        # the LP objective at mu <= 1e-4 is good to ~1e-5 relative: keep a 1e-4 safety margin
        bound = bound - 1e-4 * max(1.0, abs(bound))
        b = math.ceil(bound) if self.integer_costs else bound
        return b >= self.incumbent - (1e-9 if self.integer_costs else 1e-6 * max(1.0, abs(self.incumbent)))

    def _offer(self, obj: float, x: Optional[np.ndarray]):
        if x is not None and obj < self.incumbent:
            self.incumbent, self.incumbent_x = obj, x

    def _node_bound(self, nd: BnbNode, ok: bool, reason: int, primal: float, dual: float) -> Optional[float]:
        """The reference's node rule (bnb_driver.cpp:843-877).  None: the LP failed (status != SUCCESSFUL) and the
        non-root node is skipped.  Otherwise the node's dual bound: the LP's dual objective when the bound is
        reliable (CONVERGED, finite, dual <= primal), else the parent's bound - a MAX_ITER or GAP_STALLED node is
        kept and branched on."""
        if not ok:
            self.stats.infeasible += 1
            return None
        reliable = (reason == TERM_CONVERGED and np.isfinite(dual) and np.isfinite(primal)
                    and dual <= primal + 1e-9 * max(1.0, abs(primal)))
        if reason == TERM_MAX_ITER:
            self.stats.maxiter_nodes += 1
        elif reason == TERM_GAP_STALLED:
            self.stats.gap_stalled_nodes += 1
        return max(nd.parent_bound, dual) if reliable else nd.parent_bound

    def _branch_from_device(self, nd: BnbNode, slot: int, bound: float, feasible: bool, cover_obj: float,
                            branch_var: int, branch_frac: float, rounded_obj: float, nif_feasible: bool = False,
                            nif_obj: float = math.inf):
        """What follows a converged, unpruned node LP when the heuristics ran on the device: incumbent offers
        (the cover, and c.rint(x) for an integral LP point) and the two children.  Shared by the window and the
        continuous drivers."""
        # heuristics in the reference's order; the first one that improves the incumbent is taken (bnb_driver.cpp:885-903)
        if nif_feasible and nif_obj < self.incumbent:
            self._offer(nif_obj, get_rounded(self._cur_ws[slot], self.base.n_orig))
        elif feasible and cover_obj < self.incumbent:
            self._offer(cover_obj, get_cover(self._cur_ws[slot], self.base.n_orig))
        if branch_var < 0 or branch_frac < 1e-6:           # integral LP point
            self.stats.integral += 1
            if rounded_obj < self.incumbent:
                x = get_primal(self._cur_ws[slot], self.base.n + len(nd.decisions))[:self.base.n_orig]
                self._offer(rounded_obj, np.round(x))
            return
        warm = None
        if self.warm_start and self._export[slot] is not None:
            k = len(nd.decisions)
            warm = (self._export[slot], self.base.n + k, self.base.m + k)
        self.frontier.append(BnbNode(nd.decisions + ((branch_var, 0),), bound, warm))
        self.frontier.append(BnbNode(nd.decisions + ((branch_var, 1),), bound, warm))

    def _new_export(self, slot: int, depth: int):
        """device buffer that receives the node's final x | y | s (kept alive by the children that start from it).
        The buffers are views of ONE arena allocated up front: an allocation inside the search synchronises the device
        (measured: 66-450 ms spikes in 35-110 ms windows).  An empty pool means the node's children start cold."""
        if not self.warm_start:
            return None
        import torch
        if self._arena is None:
            size = 2 * (self.base.n + self.max_depth + 1) + self.base.m + self.max_depth + 1
            count = max(4 * self.slots, self.warm_pool)
            self._arena = torch.empty((count, size), dtype=torch.float64, device=torch.device("cuda", self.env.cudaDeviceId))
            self._pool_free = [self._arena[i] for i in range(count)]
        old = self._export[slot]
        self._export[slot] = self._pool_free.pop() if self._pool_free else None
        self._retire(old)
        t = self._export[slot]
        return None if t is None else t.data_ptr()

    def _retire(self, t):
        """a buffer goes back to the pool when neither a slot nor an open node refers to it any more"""
        import sys
        if t is not None and sys.getrefcount(t) <= 3:       # `t`, the getrefcount argument, the caller's local
            self._pool_free.append(t)

    @staticmethod
    def _warm_arg(nd: "BnbNode"):
        return None if nd.warm is None else (nd.warm[0].data_ptr(), nd.warm[1], nd.warm[2])

    def _pop_batch(self) -> List[BnbNode]:
        batch: List[BnbNode] = []
        while self.frontier and len(batch) < self.slots:
            nd = self.frontier.popleft()                       # FIFO, bnb.cpp:42-43
            if self._prunable(nd.parent_bound):                # bnb_driver.cpp:797
                self.stats.pruned_by_bound += 1
                continue
            batch.append(nd)
        return batch

    def _process_batch(self, batch, results, heur):
        """bounds, incumbent offers and children of a solved window (bnb_driver.cpp:844-1005)"""
        self.stats.round_max_iterations += max(r.iterations for r in results)
        for slot, (nd, res) in enumerate(zip(batch, results)):
            self.stats.processed += 1
            self.stats.lp_iterations += res.iterations
            self.stats.lp_device_ms += res.msStart + res.msSetup + res.msLoop
            self.stats.kernels_launched += int(res.kernelsLaunched)
            bound = self._node_bound(nd, res.status == CODE_SUCCESSFUL, res.terminationReason, res.primalObj,
                                     res.dualObj)
            if bound is None:                              # failed non-root node is skipped (bnb_driver.cpp:844-859)
                continue
            if not nd.decisions:
                self.stats.root_bound = bound
            if self._prunable(bound):
                self.stats.pruned_by_bound += 1
                continue
            if heur is not None:                           # same rules, computed on the device
                h = heur[slot]
                self.stats.kernels_launched += 1
                self._branch_from_device(nd, slot, bound, h.feasible, h.coverObj, h.branchVar, h.branchFrac,
                                         h.roundedObj, h.nifFeasible, h.nifObj)
                continue
            x = res.primalSolution[:self.base.n_orig]
            zero_fixed = [v for v, f in nd.decisions if f == 0]
            if self._pool is not None:
                self._pending.append(self._pool.submit(self.heur, x.copy(), zero_fixed))
            else:
                self._offer(*self.heur(x, zero_fixed))
            frac = np.abs(x - np.round(x))
            j = int(np.argmax(frac))
            if frac[j] < 1e-6:                             # integral LP point
                self.stats.integral += 1
                self._offer(float(self.base.c[:self.base.n_orig] @ np.round(x)), np.round(x))
                continue
            self.frontier.append(BnbNode(nd.decisions + ((j, 0),), bound))
            self.frontier.append(BnbNode(nd.decisions + ((j, 1),), bound))

    def _round_pipelined(self) -> int:
        """One window's worth of results per call, with up to ``pipeline`` windows in flight (sb200_window_begin /
        _finish over alternating workspace sets)."""
        t_round = time.perf_counter()
        while self._free_sets and self.frontier:
            if any(len(nd.decisions) > self.max_depth for nd in itertools.islice(self.frontier, self.slots)):
                break                                          # deeper than the slots take: drained below, then the plain path
            batch = self._pop_batch()
            if not batch:
                continue
            si = self._free_sets.pop()
            w = window_begin(self.base_node, [nd.decisions for nd in batch], self.cfg, self._sets[si], with_rules=True)
            if w is None:                                      # not a one-launch window: solve it here and now
                results = solve_batch_nodes(self.base_node, [nd.decisions for nd in batch], self.cfg, self._sets[si],
                                            fetch_solutions=False, fetch_trace=False)
                heur = node_heuristics(self._sets[si][:len(batch)])
                self._cur_ws = self._sets[si]
                self._process_batch(batch, results, heur)
                self._free_sets.append(si)
                continue
            self.stats.delta_rows += sum(len(nd.decisions) for nd in batch)
            self._inflight.append((batch, w, si))
        done = 0
        if self._inflight:
            batch, w, si = self._inflight.popleft()
            t_s = time.perf_counter()
            results, heur = window_finish(w)
            self.stats.solve_ms.append(round(1e3 * (time.perf_counter() - t_s), 2))      # time blocked waiting for the window
            self.stats.window_ms.append(round(last_window(self._sets[si][0])[0], 2))
            self.stats.heur_ms.append(0.0)
            self._cur_ws = self._sets[si]
            self._process_batch(batch, results, heur)
            self._free_sets.append(si)
            done = len(batch)
        elif self.frontier:                                    # nodes deeper than max_depth: the plain path takes them
            self._cur_ws = self.ws
            return self._round_plain()
        self._collectives()
        self.stats.rounds += 1
        self.stats.round_ms.append(round(1e3 * (time.perf_counter() - t_round), 2))
        return done

    def round(self) -> int:
        """Pop up to K nodes, solve their LPs as one batch, branch.  Returns the number processed."""
        if self.pipeline > 1 and self.device_nodes and self.device_heuristics:
            return self._round_pipelined()
        return self._round_plain()

    def _round_plain(self) -> int:
        t_round = time.perf_counter()
        self._cur_ws = self.ws
        batch = self._pop_batch()
        if batch:
            if self.device_nodes and all(len(nd.decisions) <= self.max_depth for nd in batch):
                on_dev = self.device_heuristics
                export = [self._new_export(i, len(nd.decisions)) for i, nd in enumerate(batch)] if self.warm_start else None
                warm = [self._warm_arg(nd) for nd in batch] if self.warm_start else None
                t_s = time.perf_counter()
                results = solve_batch_nodes(self.base_node, [nd.decisions for nd in batch], self.cfg, self.ws,
                                            fetch_solutions=not on_dev, warm=warm, export=export,
                                            warm_floor=self.warm_floor, fetch_trace=False)
                t_h = time.perf_counter()
                heur = node_heuristics(self.ws[:len(batch)]) if on_dev else None
                self.stats.solve_ms.append(round(1e3 * (t_h - t_s), 2))
                self.stats.heur_ms.append(round(1e3 * (time.perf_counter() - t_h), 2))
                self.stats.window_ms.append(round(last_window(self.ws[0])[0], 2))
                self.stats.delta_rows += sum(len(nd.decisions) for nd in batch)
            else:
                nodes = []
                for nd in batch:
                    mdl = build_branch_model(self.base, nd.decisions)
                    nodes.append(SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals,
                                                          mdl.c, mdl.b, self.env))
                results = solve_batch(nodes, self.cfg, self.ws[:len(nodes)])
                heur = None
                self.device_nodes = self.device_heuristics = False      # the slots no longer hold the base model
            self._fold_pending()               # heuristics of the previous round (ran beside this solve)
            self._process_batch(batch, results, heur)
            if self.warm_start:
                for nd in batch:                               # children solved: their parent's buffer may be free now
                    if nd.warm is not None:
                        t, nd.warm = nd.warm[0], None
                        self._retire(t)
                        del t
        else:
            self._fold_pending()
        self._collectives()
        self.stats.rounds += 1
        self.stats.round_ms.append(round(1e3 * (time.perf_counter() - t_round), 2))
        return len(batch)

    def _collectives(self):
        ax = self.async_exchange
        if ax is not None:
            # post this round's (incumbent, open, processed); act on the gather posted `lag` rounds ago, which
            # every rank reads identically - so every rank adopts the same bound and takes the same decision to
            # rebalance, without waiting for anybody's current round
            t0 = time.perf_counter()
            ax.post(self.incumbent, self._open_nodes(), self.stats.processed)
            for rows in ax.collect():
                self._apply_gather(rows)
            self.stats.exchange_wait_ms += 1e3 * (time.perf_counter() - t0)
            return
        if self.exchange is not None:                          # every rank calls it once per round
            self.incumbent, self.incumbent_x = self.exchange(self.incumbent, self.incumbent_x)
        if self.rebalance is not None and self.stats.rounds % self.rebalance_every == 0:
            nodes, self.global_open, sent, recv = self.rebalance([(nd.decisions, nd.parent_bound) for nd in self.frontier])
            if sent or recv:
                self.frontier = collections.deque(BnbNode(d, b) for d, b in nodes)
            self.stats.nodes_sent += sent
            self.stats.nodes_received += recv

    def _apply_gather(self, rows):
        """rows: [world, 3] = (incumbent objective, open nodes, processed nodes) of every rank at one round."""
        best = float(rows[:, 0].min())
        if best < self.incumbent:                              # somebody else's incumbent: only its value prunes
            self.incumbent, self.incumbent_x = best, None
        sizes = [int(v) for v in rows[:, 1].tolist()]
        self.global_open = sum(sizes)
        self.global_processed = int(rows[:, 2].sum())
        if self.rebalance is not None and max(sizes) - min(sizes) > self.rebalance_min_imbalance:
            # every rank saw the same sizes, so every rank enters the (blocking) donation here
            nodes, _total, sent, recv = self.rebalance([(nd.decisions, nd.parent_bound) for nd in self.frontier])
            if sent or recv:
                self.frontier = collections.deque(BnbNode(d, b) for d, b in nodes)
            self.stats.nodes_sent += sent
            self.stats.nodes_received += recv

    def finish_exchange(self):
        """End of a multi-rank search: drain the posted gathers and fetch the best incumbent with its vector."""
        ax = self.async_exchange
        if ax is None:
            return
        for rows in ax.collect(drain=True):
            best = float(rows[:, 0].min())
            if best < self.incumbent:
                self.incumbent, self.incumbent_x = best, None
        own = self.incumbent if self.incumbent_x is not None else math.inf
        best, bx, _owner = ax.final_incumbent(own, self.incumbent_x, self.base.n_orig)
        if bx is not None and best <= self.incumbent:
            self.incumbent, self.incumbent_x = best, bx

    def stream_round(self, node_limit: int) -> int:
        """Continuous batching (``sb200_solve_stream``): up to ``node_limit`` nodes are started, each slot taking
        the next open node the moment its LP and its heuristics kernel are done - no slot waits for the slowest
        LP of a window.  Children join the frontier while other nodes are still in flight, so the ORDER of the
        search depends on completion times (the optimum does not).  Ends with the same collectives as
        ``round``.  Returns the number of nodes processed."""
        import ctypes as C
        from . import _lib as L
        from .solver import _params_from
        if not (self.device_nodes and self.device_heuristics):
            raise RuntimeError("stream_round needs device-resident node models and device heuristics")
        lib = L.load()
        t_round = time.perf_counter()
        p = _params_from(self.base_node, self.cfg)
        k = self.slots
        handles = (C.c_void_p * k)(*[ws.handle for ws in self.ws])
        in_slot: List[Optional[BnbNode]] = [None] * k
        keep = [None] * k                       # the decision arrays of the node in each slot
        started = [0]
        failure: List[BaseException] = []
        deep: List[BnbNode] = []
        st = self.stats

        def next_cb(_user, slot, delta):
            try:
                if failure or started[0] >= node_limit:
                    return 0
                while self.frontier:
                    nd = self.frontier.popleft()
                    if self._prunable(nd.parent_bound):
                        st.pruned_by_bound += 1
                        continue
                    d = len(nd.decisions)
                    if d > self.max_depth:                     # deeper than the workspaces were sized for
                        deep.append(nd)
                        continue
                    var = np.fromiter((v for v, _ in nd.decisions), dtype=np.int32, count=d)
                    fix = np.fromiter((f for _, f in nd.decisions), dtype=np.float64, count=d)
                    coef = np.where(fix == 0.0, -1.0, 1.0)
                    keep[slot] = (var, coef, fix)
                    delta[0].n_extra_rows = d
                    delta[0].var = var.ctypes.data_as(C.POINTER(C.c_int))
                    delta[0].coef = coef.ctypes.data_as(C.POINTER(C.c_double))
                    delta[0].rhs = fix.ctypes.data_as(C.POINTER(C.c_double))
                    if self.warm_start:
                        w = self._warm_arg(nd)
                        if w is not None:
                            delta[0].warm_start, delta[0].warm_n, delta[0].warm_m = w
                            delta[0].warm_floor = self.warm_floor
                        keep[slot] = keep[slot] + (nd.warm,)            # the parent's buffer lives until this LP is done
                        delta[0].export_xys = self._new_export(slot, d)
                    in_slot[slot] = nd
                    started[0] += 1
                    st.delta_rows += d
                    return 1
                return 0
            except BaseException as e:          # an exception must not unwind through the C frame
                failure.append(e)
                return 0

        def done_cb(_user, slot, res_p, heur_p):
            try:
                nd, r, h = in_slot[slot], res_p[0], heur_p[0]
                in_slot[slot] = None
                st.processed += 1
                st.lp_iterations += r.iterations
                st.lp_device_ms += r.ms_start + r.ms_setup + r.ms_loop
                st.kernels_launched += int(r.kernels_launched) + 1
                bound = self._node_bound(nd, r.status == L.SB200_OK, r.reason, r.primal_obj, r.dual_obj)
                if bound is None:
                    return
                if not nd.decisions:
                    st.root_bound = bound
                if self._prunable(bound):
                    st.pruned_by_bound += 1
                    return
                self._branch_from_device(nd, slot, bound, bool(h.feasible), h.cover_obj, h.branch_var, h.branch_frac,
                                         h.rounded_obj, bool(h.nif_feasible), h.nif_obj)
            except BaseException as e:
                failure.append(e)

        before = st.processed
        ncb, dcb = L.NEXT_NODE_FN(next_cb), L.NODE_DONE_FN(done_cb)
        rc = lib.sb200_solve_stream(handles, k, C.byref(p), ncb, dcb, None)
        if failure:
            raise failure[0]
        if rc != L.SB200_OK:
            msgs = "; ".join(lib.sb200_last_error(ws.handle).decode() for ws in self.ws)
            raise RuntimeError(f"sb200_solve_stream failed (code {rc}): {msgs}")
        self.frontier.extendleft(reversed(deep))
        self._collectives()
        st.rounds += 1
        st.round_ms.append(round(1e3 * (time.perf_counter() - t_round), 2))
        return st.processed - before

    def run(self, max_nodes: int, rounds: Optional[int] = None, stream_nodes: int = 0) -> BnbStats:
        """``stream_nodes`` > 0: every round is a ``stream_round`` of that many nodes (continuous batching)
        instead of one window of K nodes."""
        t0 = time.perf_counter()
        r = 0
        # with several ranks the stop test must be the same on every rank (the collectives of a round are
        # matched): open nodes over ALL ranks, as of the last rebalance
        multi = self.exchange is not None or self.rebalance is not None or self.async_exchange is not None
        if multi and rounds is None and self.async_exchange is None and not (self.rebalance is not None and self.rebalance_every == 1):
            # the collectives of a round are matched calls: a rank-local stop test would leave ranks in different
            # rounds and deadlock them
            raise ValueError("multi-rank run to completion needs async_exchange, or rebalance with rebalance_every=1, "
                             "or a fixed number of rounds")

        def more():
            if rounds is not None:
                return r < rounds
            if self.async_exchange is not None:                # the same lagged gather on every rank
                if self.global_open is None:
                    return True
                return self.global_open > 0 and (self.global_processed or 0) < max_nodes
            if self.rebalance is not None and self.rebalance_every == 1 and self.global_open is not None:
                return self.global_open > 0
            return (bool(self.frontier) or bool(self._inflight)) and self.stats.processed < max_nodes
        while more():
            if stream_nodes > 0 and self.device_nodes and self.device_heuristics:
                if self.stream_round(stream_nodes) == 0 and self.frontier and rounds is None:
                    self.round()                               # only nodes deeper than max_depth are left
            else:
                self.round()
            r += 1
            if not self.frontier and self._pending:            # last word of the heuristics before stopping
                self._fold_pending()
        if rounds is None:
            self.drain()
        self._fold_pending()
        if rounds is None:
            self.finish_exchange()
        self.stats.wall_s += time.perf_counter() - t0
        self.stats.incumbent = self.incumbent
        self.stats.open_nodes = self._open_nodes()
        return self.stats
