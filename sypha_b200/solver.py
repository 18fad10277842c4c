"""Host-side mirror of the reference's operator interface for the IPM hot path.

Same names, argument meaning and error behaviour as the reference
(/root/reference/src/sypha_solver_sparse.h:13-51, src/sypha_solver.h:76-108,
src/sypha_node_sparse.h:26-119), over the C ABI of libsypha_b200.so.  Python only marshals
buffers; every number is computed by the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import List, Optional

import numpy as np

from . import _lib as L

# enum SyphaStatus (src/common.h:24-29)
CODE_SUCCESSFUL, CODE_GENERIC_ERROR, CODE_MODEL_TYPE_NOT_FOUND = range(3)
# enum SolverTerminationReason (src/sypha_solver_sparse.h:13-20)
(SOLVER_TERM_CONVERGED, SOLVER_TERM_MAX_ITER, SOLVER_TERM_GAP_STALLED,
 SOLVER_TERM_INFEASIBLE_OR_NUMERICAL, SOLVER_TERM_TIME_LIMIT) = range(5)


class Sb200Error(RuntimeError):
    """A CUDA / library failure.  The reference treats these as fatal (checkCudaErrors -> exit,
    src/sypha_cuda_helper.h:19-31); the Python shim raises instead."""


@dataclasses.dataclass
class SolverGapStagnationConfig:            # src/sypha_solver_sparse.h:22-27
    enabled: bool = False
    windowIterations: int = 0
    minImprovementPct: float = 0.0


@dataclasses.dataclass
class SolverExecutionConfig:                # src/sypha_solver_sparse.h:29-36
    maxIterations: int = 25
    gapStagnation: SolverGapStagnationConfig = dataclasses.field(default_factory=SolverGapStagnationConfig)
    bnbNodeOrdinal: int = 0
    denseSelectionLogEveryNodes: int = 1
    skipGpuMemorySampling: bool = False


@dataclasses.dataclass
class SolverExecutionResult:                # src/sypha_solver_sparse.h:38-48
    status: int = CODE_GENERIC_ERROR
    terminationReason: int = SOLVER_TERM_MAX_ITER
    iterations: int = 0
    primalObj: float = 0.0
    dualObj: float = 0.0
    relativeGap: float = math.inf
    primalSolution: Optional[np.ndarray] = None
    dualSolution: Optional[np.ndarray] = None
    # extras (not in the reference struct)
    slackSolution: Optional[np.ndarray] = None
    mu: float = math.nan
    msStart: float = 0.0
    msSetup: float = 0.0
    msLoop: float = 0.0
    strategyUsed: int = 0
    cgIterations: int = 0
    kernelsLaunched: int = 0
    trace: Optional[np.ndarray] = None


@dataclasses.dataclass
class SyphaEnvironment:
    """The parameter surface of src/sypha_environment.h:89-100 that the hot path reads
    (defaults: src/sypha_environment_defaults.h:14-30)."""
    mehrotraMaxIter: int = 25
    mehrotraEta: float = 0.95
    mehrotraMuTol: float = 1e-4
    linearSolverStrategy: str = "auto"      # auto | cholesky | syrk | pcg (krylov)
    krylovMaxCgIter: int = 500
    krylovCgTolInitial: float = 1e-2
    krylovCgTolFinal: float = 1e-8
    krylovCgTolDecayRate: float = 0.5
    cudaDeviceId: int = 0
    pollEvery: int = 1
    useGraph: bool = True
    stopRequested: Optional[C.c_int] = None     # logger watchdog flag (src/sypha_logger.h:69)


class IpmWorkspace:
    """Persistent device workspace (src/sypha_solver.h:76-105).  Grow-only; reused across LPs."""

    def __init__(self):
        self.handle = C.c_void_p()
        self.isAllocated = False
        self.device = 0

    def __del__(self):
        try:
            releaseIpmWorkspace(self)
        except Exception:
            pass


def initializeIpmWorkspace(ws: IpmWorkspace, maxKktNrows: int = 0, maxKktNnz: int = 0, maxNcols: int = 0,
                           device: int = 0):
    """src/sypha_solver.h:107.  The sizing arguments are KKT-shaped in the reference
    (src/sypha_solver_bnb_driver.cpp:620-626: rows 2n+m, nnz 2nnz+3n); recover m, n, nnz."""
    lib = L.load()
    caps = None
    if maxKktNrows > 0 and maxNcols > 0 and maxKktNnz > 0:
        m_max = maxKktNrows - 2 * maxNcols
        nnz_max = (maxKktNnz - 3 * maxNcols) // 2
        if m_max > 0 and nnz_max > 0:
            caps = L.sb200_caps(m_max, maxNcols, nnz_max)
    rc = lib.sb200_ws_create(device, C.byref(caps) if caps else None, C.byref(ws.handle))
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_ws_create failed (code {rc}): no CUDA device or out of memory; "
                         "there is no CPU fallback")
    ws.isAllocated = True
    ws.device = device


def releaseIpmWorkspace(ws: IpmWorkspace):
    """src/sypha_solver.h:108."""
    if ws.isAllocated and ws.handle:
        L.load().sb200_ws_destroy(ws.handle)
    ws.handle = C.c_void_p()
    ws.isAllocated = False


class SyphaNodeSparse:
    """The model container fields the solver reads/writes (src/sypha_node_sparse.h:26-119)."""

    def __init__(self, env: Optional[SyphaEnvironment] = None):
        self.env = env or SyphaEnvironment()
        self.ncols = self.nrows = self.ncolsOriginal = self.nnz = 0
        self.hCsrMatInds = self.hCsrMatOffs = self.hCsrMatVals = None
        self.hObjDns = self.hRhsDns = None
        self.hX = self.hY = self.hS = None
        self.objvalPrim = self.objvalDual = 0.0
        self.mipGap = math.inf
        self.iterations = 0
        self.timeStartSol = self.timePreSol = self.timeSolver = 0.0
        self._loaded_into = None

    @classmethod
    def from_csr(cls, m, n, n_orig, offs, inds, vals, c, b, env=None):
        node = cls(env)
        node.nrows, node.ncols, node.ncolsOriginal = int(m), int(n), int(n_orig)
        node.hCsrMatOffs = np.ascontiguousarray(offs, dtype=np.int32)
        node.hCsrMatInds = np.ascontiguousarray(inds, dtype=np.int32)
        node.hCsrMatVals = np.ascontiguousarray(vals, dtype=np.float64)
        node.hObjDns = np.ascontiguousarray(c, dtype=np.float64)
        node.hRhsDns = np.ascontiguousarray(b, dtype=np.float64)
        node.nnz = int(node.hCsrMatInds.shape[0])
        return node

    def copyModelOnDevice(self, workspace: IpmWorkspace, strategy: Optional[str] = None):
        """src/sypha_node_sparse.cpp:156-198: upload CSR, c, b (and build CSC + symbolic M)."""
        lib = L.load()
        strat = L.STRATEGY_NAMES[(strategy or self.env.linearSolverStrategy).lower()]
        rc = lib.sb200_load_model(
            workspace.handle, self.nrows, self.ncols, self.ncolsOriginal, self.nnz,
            self.hCsrMatOffs.ctypes.data, self.hCsrMatInds.ctypes.data, self.hCsrMatVals.ctypes.data,
            self.hObjDns.ctypes.data, self.hRhsDns.ctypes.data, 0, strat)
        if rc != L.SB200_OK:
            raise Sb200Error(f"sb200_load_model failed (code {rc}): "
                             f"{lib.sb200_last_error(workspace.handle).decode()}")
        self._loaded_into = workspace
        return CODE_SUCCESSFUL


def _params_from(node: SyphaNodeSparse, config: SolverExecutionConfig) -> L.sb200_params:
    lib = L.load()
    p = L.sb200_params()
    lib.sb200_default_params(C.byref(p))
    env = node.env
    # src/sypha_solver.cpp:488
    p.max_iter = config.maxIterations if config.maxIterations > 0 else env.mehrotraMaxIter
    p.eta = env.mehrotraEta
    p.mu_tol = env.mehrotraMuTol
    g = config.gapStagnation
    p.gap_enabled = 1 if g.enabled else 0
    p.gap_window = g.windowIterations
    p.gap_min_improv_pct = g.minImprovementPct
    p.cg_max_iter = env.krylovMaxCgIter
    p.cg_tol_initial = env.krylovCgTolInitial
    p.cg_tol_final = env.krylovCgTolFinal
    p.cg_tol_decay = env.krylovCgTolDecayRate
    p.poll_every = env.pollEvery
    p.use_graph = 1 if env.useGraph else 0
    if env.stopRequested is not None:
        p.stop_flag = C.pointer(env.stopRequested)
    return p


def _fill_result(node, ws, res: L.sb200_result, result: SolverExecutionResult, x, y, s, fetch_trace=True):
    lib = L.load()
    result.status = CODE_SUCCESSFUL if res.status == L.SB200_OK else CODE_GENERIC_ERROR
    result.terminationReason = res.reason
    result.iterations = res.iterations
    result.primalObj, result.dualObj, result.relativeGap = res.primal_obj, res.dual_obj, res.rel_gap
    result.primalSolution, result.dualSolution, result.slackSolution = x, y, s
    result.mu = res.mu
    result.msStart, result.msSetup, result.msLoop = res.ms_start, res.ms_setup, res.ms_loop
    result.strategyUsed = res.strategy_used
    result.cgIterations = res.cg_iterations
    result.kernelsLaunched = res.kernels_launched
    if fetch_trace:
        tr = np.zeros((max(res.iterations, 1), L.TRACE_COLS))
        rows = lib.sb200_get_trace(ws.handle, tr.ctypes.data, tr.shape[0])
        result.trace = tr[:rows]
    else:
        result.trace = None
    # node.* outputs, src/sypha_solver.cpp:774-797
    node.iterations = res.iterations
    node.objvalPrim, node.objvalDual = res.primal_obj, res.dual_obj
    node.mipGap = math.inf
    node.timeStartSol, node.timePreSol, node.timeSolver = res.ms_start / 1e3, res.ms_setup / 1e3, res.ms_loop / 1e3


def solver_sparse_mehrotra_run(node: SyphaNodeSparse, config: SolverExecutionConfig,
                               result: Optional[SolverExecutionResult] = None,
                               workspace: Optional[IpmWorkspace] = None) -> int:
    """SyphaStatus solver_sparse_mehrotra_run(node, config, result, workspace)
    - src/sypha_solver_sparse.h:51.  With ``workspace=None`` everything is allocated and freed
    inside the call (src/sypha_solver.cpp:197-203, :841-871)."""
    lib = L.load()
    own = workspace is None or not workspace.isAllocated
    ws = workspace
    if own:
        ws = IpmWorkspace()
        initializeIpmWorkspace(ws, device=node.env.cudaDeviceId)
    try:
        if node._loaded_into is not ws:
            node.copyModelOnDevice(ws)
        p = _params_from(node, config)
        res = L.sb200_result()
        x = np.empty(node.ncols)
        y = np.empty(node.nrows)
        s = np.empty(node.ncols)
        res.x_host, res.y_host, res.s_host = x.ctypes.data, y.ctypes.data, s.ctypes.data
        rc = lib.sb200_solve(ws.handle, C.byref(p), C.byref(res))
        if rc != L.SB200_OK:
            raise Sb200Error(f"sb200_solve failed (code {rc}): {lib.sb200_last_error(ws.handle).decode()}")
        if result is None:
            result = SolverExecutionResult()
        _fill_result(node, ws, res, result, x, y, s)
        return result.status
    finally:
        if own:
            node._loaded_into = None
            releaseIpmWorkspace(ws)


def solver_sparse_mehrotra(node: SyphaNodeSparse) -> int:
    """Thin wrapper, src/sypha_solver.cpp:25-40."""
    config = SolverExecutionConfig(maxIterations=node.env.mehrotraMaxIter)
    result = SolverExecutionResult()
    status = solver_sparse_mehrotra_run(node, config, result)
    return status if status != CODE_SUCCESSFUL else result.status


def solve_batch(nodes, config: SolverExecutionConfig, workspaces, host_bufs=None):
    """Solve independent LPs concurrently (one workspace/stream each) - the B&B node body of
    src/sypha_solver_bnb_driver.cpp:789-859 batched per GPU.  ``host_bufs[i]`` = (x, y, s) float64 arrays that receive
    LP i's solution (e.g. views of pinned memory); allocated here when absent."""
    lib = L.load()
    k = len(nodes)
    for node, ws in zip(nodes, workspaces):
        if node._loaded_into is not ws:
            node.copyModelOnDevice(ws)
    p = _params_from(nodes[0], config)
    handles = (C.c_void_p * k)(*[ws.handle for ws in workspaces])
    res = (L.sb200_result * k)()
    bufs = []
    for i, node in enumerate(nodes):
        if host_bufs is not None:
            x, y, s = host_bufs[i]
            if x.size < node.ncols or y.size < node.nrows or s.size < node.ncols:
                raise ValueError(f"host_bufs[{i}] is smaller than LP {i}")
        else:
            x, y, s = np.empty(node.ncols), np.empty(node.nrows), np.empty(node.ncols)
        res[i].x_host, res[i].y_host, res[i].s_host = x.ctypes.data, y.ctypes.data, s.ctypes.data
        bufs.append((x, y, s))
    rc = lib.sb200_solve_batch(handles, k, None, C.byref(p), res)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_solve_batch failed (code {rc})")
    out = []
    for i, node in enumerate(nodes):
        r = SolverExecutionResult()
        _fill_result(node, workspaces[i], res[i], r, *bufs[i])
        out.append(r)
    return out


def workspace_for_nodes(base: SyphaNodeSparse, max_depth: int, device: int = 0) -> IpmWorkspace:
    """A workspace sized like the B&B driver sizes its IpmWorkspace (src/sypha_solver_bnb_driver.cpp:618-627,
    KKT-shaped arguments) for nodes up to ``max_depth`` appended rows, with the BASE model resident."""
    n_max, m_max, nnz_max = base.ncols + max_depth, base.nrows + max_depth, base.nnz + 2 * max_depth
    ws = IpmWorkspace()
    initializeIpmWorkspace(ws, maxKktNrows=2 * n_max + m_max, maxKktNnz=2 * nnz_max + 3 * n_max, maxNcols=n_max,
                           device=device)
    base.copyModelOnDevice(ws, "cholesky")
    rc = L.load().sb200_prepare_nodes(ws.handle, max_depth)      # nothing is allocated at the workspace's first node
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_prepare_nodes failed (code {rc}): {L.load().sb200_last_error(ws.handle).decode()}")
    return ws


@dataclasses.dataclass
class NodeHeuristicResult:
    """What ``sb200_node_heuristics`` returns per node (include/sypha_b200.h ``sb200_heur_result``)."""
    feasible: bool
    coverObj: float
    nChosen: int
    branchVar: int
    branchFrac: float
    roundedObj: float
    repairSteps: int
    nifFeasible: bool = False          # NearestIntegerFixingHeuristic (reference rules): rounding + decisions covers
    nifObj: float = float("inf")


FORM_LATENCY, FORM_THROUGHPUT = 0, 1


def last_window(workspace):
    """(device milliseconds, LPs) of the last one-launch window whose FIRST workspace was ``workspace``: CUDA events on the
    launching stream around the kernel (sb200_last_window)."""
    lib = L.load()
    ms, k = C.c_double(), C.c_int()
    lib.sb200_last_window(workspace.handle, C.byref(ms), C.byref(k))
    return ms.value, k.value


def set_solver_form(workspace, form: str = "latency"):
    """"latency": one LP over the whole GPU, ~12 kernels per iteration (default).  "throughput": the whole LP by one
    thread block in one launch (csrc/sb200_cta.cu) - the form for many LPs in flight (B&B nodes)."""
    lib = L.load()
    rc = lib.sb200_set_solver_form(workspace.handle, FORM_THROUGHPUT if form == "throughput" else FORM_LATENCY)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_set_solver_form failed (code {rc})")


HEUR_PLAIN, HEUR_REFERENCE = 0, 1
BRANCH_RULES = {"most_fractional": 0, "highest_cost_fractional": 1}


def set_heuristic_rules(workspace, rules: str = "reference", branch_rule: str = "most_fractional", tol: float = 1e-6):
    """Which per-node rules ``node_heuristics`` runs on this workspace: the reference's own ("reference":
    NearestIntegerFixing + DualGuidedCoverRepair, sypha_solver_heuristics.cpp:53-292, selector by
    ``bnb_var_selection``) or the plain rounding / greedy repair ("plain")."""
    lib = L.load()
    rc = lib.sb200_set_heuristic_rules(workspace.handle, HEUR_REFERENCE if rules == "reference" else HEUR_PLAIN,
                                       BRANCH_RULES[branch_rule], tol)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_set_heuristic_rules failed (code {rc})")


def node_heuristics(workspaces) -> List[NodeHeuristicResult]:
    """Branching variable + rounding/repair incumbent of the nodes whose LPs were just solved in
    ``workspaces`` - on the device, concurrently (reference: host loop, bnb_driver.cpp:861-1005)."""
    lib = L.load()
    k = len(workspaces)
    handles = (C.c_void_p * k)(*[ws.handle for ws in workspaces])
    out = (L.sb200_heur_result * k)()
    rc = lib.sb200_node_heuristics(handles, k, out)
    if rc != L.SB200_OK:
        msgs = "; ".join(lib.sb200_last_error(ws.handle).decode() for ws in workspaces)
        raise Sb200Error(f"sb200_node_heuristics failed (code {rc}): {msgs}")
    return [NodeHeuristicResult(bool(o.feasible), o.cover_obj, o.n_chosen, o.branch_var, o.branch_frac,
                                o.rounded_obj, o.repair_steps, bool(o.nif_feasible), o.nif_obj) for o in out]


def get_primal(workspace: IpmWorkspace, n: int) -> np.ndarray:
    """Host copy of the resident primal iterate (length n = columns of the model as last solved)."""
    lib = L.load()
    x = np.empty(n)
    rc = lib.sb200_get_iterates(workspace.handle, x.ctypes.data, None, None)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_get_iterates failed (code {rc}): {lib.sb200_last_error(workspace.handle).decode()}")
    return x


def get_cover(workspace: IpmWorkspace, n_orig: int) -> np.ndarray:
    """The 0/1 cover found by the last ``node_heuristics`` on this workspace (float64, length n_orig)."""
    lib = L.load()
    buf = np.empty(n_orig, dtype=np.uint8)
    rc = lib.sb200_get_cover(workspace.handle, buf.ctypes.data)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_get_cover failed (code {rc}): {lib.sb200_last_error(workspace.handle).decode()}")
    return buf.astype(np.float64)


def get_rounded(workspace: IpmWorkspace, n_orig: int) -> np.ndarray:
    """The NearestIntegerFixing rounding of the last ``node_heuristics`` (reference rules), float64, length n_orig."""
    lib = L.load()
    buf = np.empty(n_orig, dtype=np.uint8)
    rc = lib.sb200_get_rounded(workspace.handle, buf.ctypes.data)
    if rc != L.SB200_OK:
        raise Sb200Error(f"sb200_get_rounded failed (code {rc}): {lib.sb200_last_error(workspace.handle).decode()}")
    return buf.astype(np.float64)


def solve_batch_nodes(base: SyphaNodeSparse, decisions_list, config: SolverExecutionConfig, workspaces,
                      fetch_solutions: bool = True, warm=None, export=None, warm_floor: float = 0.1,
                      fetch_trace: bool = True):
    """B&B node body, batched and device-resident: workspace i holds the base model; node i = base + one row
    per (var, fix) decision (bnb.cpp:453-468) is formed on the device (``sb200_node_delta``) and the LPs are
    solved concurrently.  Returns one SolverExecutionResult per node (solutions have the node's dimensions).

    ``export[i]``: device address (or None) that receives node i's final x | y | s packed; ``warm[i]``: (device
    address, n, m) of the PARENT's packed iterate (or None): the child then starts from it instead of the Mehrotra
    starting point (throughput form only; see sb200_node_delta in the header)."""
    lib = L.load()
    k = len(decisions_list)
    p = _params_from(base, config)
    handles = (C.c_void_p * k)(*[ws.handle for ws in workspaces[:k]])
    deltas = (L.sb200_node_delta * k)()
    res = (L.sb200_result * k)()
    bufs = []
    # every node's (variable, fixing) pairs in three flat arrays; the deltas point into them
    lens = [len(dec) for dec in decisions_list]
    flat = np.array([pair for dec in decisions_list for pair in dec], dtype=np.float64).reshape(-1, 2)
    var = np.ascontiguousarray(flat[:, 0].astype(np.int32))
    fix = np.ascontiguousarray(flat[:, 1])
    coef = np.where(fix == 0.0, -1.0, 1.0)
    keep = (var, coef, fix)
    vb, cb, fb = var.ctypes.data, coef.ctypes.data, fix.ctypes.data
    PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)
    o = 0
    for i, d in enumerate(lens):
        deltas[i].n_extra_rows = d
        deltas[i].var = C.cast(vb + 4 * o, PI)
        deltas[i].coef = C.cast(cb + 8 * o, PD)
        deltas[i].rhs = C.cast(fb + 8 * o, PD)
        o += d
        if warm is not None and warm[i] is not None:
            deltas[i].warm_start, deltas[i].warm_n, deltas[i].warm_m = warm[i]
            deltas[i].warm_floor = warm_floor
        if export is not None and export[i] is not None:
            res[i].xys_device = export[i]
        if fetch_solutions:
            x, y, s = np.empty(base.ncols + d), np.empty(base.nrows + d), np.empty(base.ncols + d)
            res[i].x_host, res[i].y_host, res[i].s_host = x.ctypes.data, y.ctypes.data, s.ctypes.data
            bufs.append((x, y, s))
        else:
            bufs.append((None, None, None))       # the iterates stay on the device (sb200_node_heuristics reads them)
    rc = lib.sb200_solve_batch(handles, k, deltas, C.byref(p), res)
    if rc != L.SB200_OK:
        msgs = "; ".join(lib.sb200_last_error(ws.handle).decode() for ws in workspaces[:k])
        raise Sb200Error(f"sb200_solve_batch failed (code {rc}): {msgs}")
    out = []
    for i in range(k):
        r = SolverExecutionResult()
        shadow = SyphaNodeSparse(base.env)
        _fill_result(shadow, workspaces[i], res[i], r, *bufs[i], fetch_trace=fetch_trace)
        out.append(r)
    del keep
    return out


class NodeWindow:
    """A window of node LPs in flight between ``window_begin`` and ``window_finish`` (keeps the ctypes arrays alive)."""
    __slots__ = ("handles", "k", "deltas", "res", "keep", "workspaces", "base", "with_rules", "params")


def window_begin(base: SyphaNodeSparse, decisions_list, config: SolverExecutionConfig, workspaces,
                 with_rules: bool = True) -> Optional["NodeWindow"]:
    """``sb200_window_begin``: apply the nodes' deltas, launch the window (one thread block per node LP) and - with
    ``with_rules`` - the node-rules kernel behind it, and return WITHOUT waiting.  ``None`` when the batch cannot be a
    one-launch window; the deltas are applied then and ``solve_batch(..)``-style synchronous solving is the caller's
    fallback (``window_finish`` accepts that case too)."""
    lib = L.load()
    k = len(decisions_list)
    w = NodeWindow()
    w.k, w.base, w.with_rules, w.workspaces = k, base, with_rules, list(workspaces[:k])
    w.params = _params_from(base, config)
    w.handles = (C.c_void_p * k)(*[ws.handle for ws in w.workspaces])
    w.deltas = (L.sb200_node_delta * k)()
    w.res = (L.sb200_result * k)()
    lens = [len(dec) for dec in decisions_list]
    flat = np.array([pair for dec in decisions_list for pair in dec], dtype=np.float64).reshape(-1, 2)
    var = np.ascontiguousarray(flat[:, 0].astype(np.int32))
    fix = np.ascontiguousarray(flat[:, 1])
    coef = np.where(fix == 0.0, -1.0, 1.0)
    w.keep = (var, coef, fix)
    vb, cb, fb = var.ctypes.data, coef.ctypes.data, fix.ctypes.data
    PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)
    o = 0
    for i, d in enumerate(lens):
        w.deltas[i].n_extra_rows = d
        w.deltas[i].var = C.cast(vb + 4 * o, PI)
        w.deltas[i].coef = C.cast(cb + 8 * o, PD)
        w.deltas[i].rhs = C.cast(fb + 8 * o, PD)
        o += d
    rc = lib.sb200_window_begin(w.handles, k, w.deltas, C.byref(w.params), w.res, 1 if with_rules else 0)
    if rc == L.SB200_ERR_UNSUPPORTED:
        return None
    if rc != L.SB200_OK:
        msgs = "; ".join(lib.sb200_last_error(ws.handle).decode() for ws in w.workspaces)
        raise Sb200Error(f"sb200_window_begin failed (code {rc}): {msgs}")
    return w


def window_finish(w: "NodeWindow"):
    """``sb200_window_finish``: wait for the window and return (results, node-rule results or None); solutions stay on the
    device (``get_primal`` / ``get_cover`` read them)."""
    lib = L.load()
    heur = (L.sb200_heur_result * w.k)() if w.with_rules else None
    rc = lib.sb200_window_finish(w.handles, w.k, w.res, heur)
    if rc != L.SB200_OK:
        msgs = "; ".join(lib.sb200_last_error(ws.handle).decode() for ws in w.workspaces)
        raise Sb200Error(f"sb200_window_finish failed (code {rc}): {msgs}")
    out = []
    for i in range(w.k):
        r = SolverExecutionResult()
        _fill_result(SyphaNodeSparse(w.base.env), w.workspaces[i], w.res[i], r, None, None, None, fetch_trace=False)
        out.append(r)
    rules = None
    if heur is not None:
        rules = [NodeHeuristicResult(bool(h.feasible), h.cover_obj, h.n_chosen, h.branch_var, h.branch_frac,
                                     h.rounded_obj, h.repair_steps, bool(h.nif_feasible), h.nif_obj) for h in heur]
    return out, rules
