"""CPU: host logic of the batched B&B node loop (sypha_b200/bnb.py) - the branch-model builder against the
oracle's restatement of build_branch_model, and the incumbent heuristic."""
import numpy as np

from conftest import load_golden
from oracle import scp_io
from sypha_b200 import bnb
from sypha_b200.instances import ScpModel, gen_scp


def _as_model(inst):
    return ScpModel(inst.m, inst.n, inst.n_orig, inst.offs, inst.inds, inst.vals, inst.c, inst.b, inst.name)


def test_branch_model_matches_oracle_restatement():
    inst, _ = load_golden("scp41")
    dec = [(5, 1), (17, 0), (400, 1), (3, 0)]
    ref = scp_io.append_branch_rows(inst, dec)
    got = bnb.build_branch_model(_as_model(inst), dec)
    assert (got.m, got.n, got.n_orig) == (ref.m, ref.n, ref.n_orig)
    for k in ("offs", "inds", "vals", "c", "b"):
        assert np.array_equal(np.asarray(getattr(got, k), dtype=np.float64), np.asarray(getattr(ref, k), dtype=np.float64)), k
    assert bnb.build_branch_model(_as_model(inst), []) is not None


def test_cover_heuristic_is_feasible_and_respects_fixings():
    mdl = gen_scp(60, 400, 0.05, 3)
    h = bnb.CoverHeuristic(mdl)
    r = np.random.default_rng(0)
    x = r.uniform(0, 1, mdl.n)
    obj, sol = h(x)
    assert sol is not None and np.all(h.A @ sol >= 1.0)
    assert abs(obj - mdl.c[:mdl.n_orig] @ sol) < 1e-12
    # no column is redundant
    for j in np.nonzero(sol)[0]:
        y = sol.copy()
        y[j] = 0
        assert np.any(h.A @ y < 1.0)
    banned = list(np.nonzero(sol)[0][:3])
    obj2, sol2 = h(x, banned)
    assert sol2 is None or (np.all(sol2[banned] == 0) and np.all(h.A @ sol2 >= 1.0))


def test_pruning_rule_uses_integer_costs():
    mdl = gen_scp(10, 30, 0.3, 1)

    class Dummy(bnb.BatchedBnb):
        def __init__(self):           # no GPU: only the rule
            self.incumbent, self.integer_costs = 20.0, True
    d = Dummy()
    assert d._prunable(19.2) and d._prunable(20.0) and not d._prunable(18.9) and not d._prunable(19.001)
    d.incumbent = float("inf")
    assert not d._prunable(1e9)
    assert mdl.m == 10


def test_node_bound_rule():
    """bnb_driver.cpp:843-877 as restated in BatchedBnb._node_bound: a failed LP is skipped, a converged one with
    dual <= primal bounds the node, a MAX_ITER / GAP_STALLED one keeps its parent's bound and is still branched."""
    class Dummy(bnb.BatchedBnb):
        def __init__(self):
            self.stats = bnb.BnbStats()
    d = Dummy()
    nd = bnb.BnbNode(((1, 0),), 40.0)
    assert d._node_bound(nd, False, bnb.TERM_NUMERICAL, 1.0, 1.0) is None and d.stats.infeasible == 1
    assert d._node_bound(nd, True, bnb.TERM_CONVERGED, 43.2, 42.9) == 42.9
    assert d._node_bound(nd, True, bnb.TERM_CONVERGED, 39.0, 38.0) == 40.0
    assert d._node_bound(nd, True, bnb.TERM_CONVERGED, 42.0, 43.0) == 40.0
    assert d._node_bound(nd, True, bnb.TERM_MAX_ITER, 50.0, 45.0) == 40.0 and d.stats.maxiter_nodes == 1
    assert d._node_bound(nd, True, bnb.TERM_GAP_STALLED, 50.0, 45.0) == 40.0 and d.stats.gap_stalled_nodes == 1


def test_pipelined_rounds_with_a_stand_in_for_the_device(monkeypatch):
    """BatchedBnb(pipeline=2) on the CPU: the window calls (sb200_window_begin / _finish) replaced by a stand-in that
    solves each node's LP with SciPy and applies the host rules.  Checks the host logic of the windows in flight: never
    more than two windows open, sets used alternately, every window finished, the MILP optimum reached, the same
    optimum as with one window at a time."""
    import scipy.sparse as sp
    from scipy.optimize import Bounds, LinearConstraint, linprog, milp
    from sypha_b200 import solver as S

    mdl = gen_scp(30, 100, 0.1, 1)              # LP bound 179.33, integer optimum 183: the search has to branch
    A = sp.csr_matrix((mdl.vals, mdl.inds, mdl.offs), shape=(mdl.m, mdl.n))[:, :mdl.n_orig]
    c = mdl.c[:mdl.n_orig]
    opt = milp(c, constraints=LinearConstraint(A, lb=1.0), integrality=np.ones(mdl.n_orig), bounds=Bounds(0, 1))
    assert opt.success
    cover = bnb.CoverHeuristic(mdl)
    log = {"open": 0, "max_open": 0, "begun": 0, "finished": 0, "sets": []}

    class FakeWs:
        def __init__(self):
            self.cover = None
            self.x = None

    def fake_window_begin(base, decisions_list, cfg, workspaces, with_rules=True):
        assert len(decisions_list) >= 1
        log["open"] += 1
        log["begun"] += 1
        log["max_open"] = max(log["max_open"], log["open"])
        log["sets"].append(id(workspaces[0]))
        return (list(decisions_list), list(workspaces))

    def fake_window_finish(w):
        decs, wss = w
        log["open"] -= 1
        log["finished"] += 1
        results, rules = [], []
        for dec, ws in zip(decs, wss):
            lo, hi = np.zeros(mdl.n_orig), np.ones(mdl.n_orig)
            for v, f in dec:
                lo[v] = hi[v] = f
            r = linprog(c, A_ub=-A, b_ub=-np.ones(mdl.m), bounds=list(zip(lo, hi)), method="highs")
            res = S.SolverExecutionResult()
            if r.status != 0:
                res.status, res.terminationReason, res.iterations = S.CODE_GENERIC_ERROR, 3, 1
                res.primalObj = res.dualObj = float("inf")
                results.append(res)
                rules.append(S.NodeHeuristicResult(False, float("inf"), 0, -1, 0.0, float("inf"), 0))
                continue
            res.status, res.terminationReason, res.iterations = S.CODE_SUCCESSFUL, 0, 7
            res.primalObj = res.dualObj = float(r.fun)
            res.msStart = res.msSetup = res.msLoop = 0.0
            res.kernelsLaunched = 0
            x = r.x
            frac = np.abs(x - np.round(x))
            j = int(np.argmax(frac))
            # (no rounding heuristic here: incumbents come from integral LP points only, so that the tree is deep enough
            # to keep two windows open)
            ws.cover, ws.x = None, x
            rules.append(S.NodeHeuristicResult(False, float("inf"), 0, j if frac[j] >= 1e-6 else -1, float(frac[j]),
                                               float(c @ np.round(x)), 0))
            results.append(res)
        return results, rules

    monkeypatch.setattr(bnb, "workspace_for_nodes", lambda base, depth, device=0: FakeWs())
    monkeypatch.setattr(bnb, "set_heuristic_rules", lambda *a, **k: None)
    monkeypatch.setattr(bnb, "releaseIpmWorkspace", lambda w: None)
    monkeypatch.setattr(bnb, "window_begin", fake_window_begin)
    monkeypatch.setattr(bnb, "window_finish", fake_window_finish)
    monkeypatch.setattr(bnb, "last_window", lambda w: (0.0, 0))
    monkeypatch.setattr(bnb, "get_cover", lambda w, n: w.cover)
    monkeypatch.setattr(bnb, "get_primal", lambda w, n: np.concatenate([w.x, np.zeros(max(0, n - len(w.x)))]))

    def solve_plain(base, decisions_list, cfg, workspaces, **kw):       # pipeline = 1 goes through the blocking pair
        r, h = fake_window_finish((list(decisions_list), list(workspaces[:len(decisions_list)])))
        log["open"] += 1
        solve_plain.rules = h
        return r
    monkeypatch.setattr(bnb, "solve_batch_nodes", solve_plain)
    monkeypatch.setattr(bnb, "node_heuristics", lambda wss: solve_plain.rules)

    out = {}
    for pl in (2, 1):
        log.update(open=0, max_open=0, begun=0, finished=0, sets=[])
        drv = bnb.BatchedBnb(mdl, slots=2, share_gpu=False, pipeline=pl, heuristic_rules="plain")
        assert drv.pipeline == pl and len(drv.ws) == 2 * pl
        st = drv.run(max_nodes=5000)
        assert st.open_nodes == 0 and not drv._inflight and sorted(drv._free_sets) == list(range(pl))
        assert st.incumbent == round(opt.fun), (st.incumbent, opt.fun)
        assert np.all(cover.A @ drv.incumbent_x >= 1.0)
        if pl == 2:
            assert log["begun"] == log["finished"] and log["begun"] > 3, log
            assert log["max_open"] == 2
            assert len(set(log["sets"])) == 2                        # both sets of slots were used
        out[pl] = st.incumbent
        drv.close()
    assert out[1] == out[2]


def test_node_decisions_are_marshalled_as_flat_arrays(monkeypatch):
    """solve_batch_nodes / window_begin pass every node's (variable, fixing) pairs as three flat arrays with the deltas
    pointing into them (sb200_node_delta {n_extra_rows, var, coef, rhs}): a stand-in for the library reads them back."""
    import ctypes as C
    from sypha_b200 import _lib as L, solver as S

    seen = {}

    class FakeLib:
        def sb200_solve_batch(self, handles, k, deltas, params, res):
            seen["batch"] = [(deltas[i].n_extra_rows,
                              [deltas[i].var[r] for r in range(deltas[i].n_extra_rows)],
                              [deltas[i].coef[r] for r in range(deltas[i].n_extra_rows)],
                              [deltas[i].rhs[r] for r in range(deltas[i].n_extra_rows)]) for i in range(k)]
            for i in range(k):
                res[i].status, res[i].iterations = L.SB200_OK, 3
            return L.SB200_OK

        def sb200_window_begin(self, handles, k, deltas, params, res, with_rules):
            seen["window"] = [(deltas[i].n_extra_rows, [deltas[i].var[r] for r in range(deltas[i].n_extra_rows)],
                               [deltas[i].coef[r] for r in range(deltas[i].n_extra_rows)]) for i in range(k)]
            return L.SB200_ERR_UNSUPPORTED

        def sb200_last_error(self, h):
            return b""

        def sb200_default_params(self, p):
            pass

    monkeypatch.setattr(L, "load", lambda: FakeLib())
    mdl = gen_scp(10, 30, 0.3, 1)
    env = S.SyphaEnvironment()
    base = S.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)

    class W:
        handle = C.c_void_p(0)
    decs = [(), ((5, 1),), ((7, 0), (2, 1), (11, 0)), (), ((29, 0), (0, 1))]
    out = S.solve_batch_nodes(base, decs, S.SolverExecutionConfig(maxIterations=5), [W() for _ in decs],
                              fetch_solutions=False, fetch_trace=False)
    assert [r.iterations for r in out] == [3] * len(decs)
    want = [(len(d), [v for v, _ in d], [(-1.0 if f == 0 else 1.0) for _, f in d], [float(f) for _, f in d]) for d in decs]
    assert seen["batch"] == want
    assert S.window_begin(base, decs, S.SolverExecutionConfig(maxIterations=5), [W() for _ in decs]) is None
    assert seen["window"] == [(n, v, c) for n, v, c, _ in want]


def test_greedy_first_incumbent_against_the_references_own_function():
    """bnb.greedy_cover (the first incumbent of the B&B bench) against greedy_set_cover_heuristic itself
    (src/sypha_preprocessor.cpp:11-96, compiled in place into oracle/_ref/libref_prep.so).  The reference orders the
    columns by (cost, rows covered) with std::sort and leaves FULL ties to it (unstable, implementation-defined);
    greedy_cover breaks them by column index.  So: identical selections where no two columns tie, the same objective on
    the bench's instances (scpnre1 38, scpnrg1 266 - what `reduce_by_incumbent` then cuts by), feasible covers whose
    cost is what they report everywhere."""
    import ctypes as C
    from pathlib import Path
    import pytest
    from sypha_b200.instances import load_npz
    so = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "libref_prep.so"
    if not so.exists():
        pytest.skip("oracle/_ref/libref_prep.so not built (make -C oracle all)")
    lib = C.CDLL(str(so))
    PI, PD = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.ref_greedy_set_cover.argtypes = [C.c_int, C.c_int, PI, PI, PD, PD, PI, PI, PD]

    def reference(m):
        offs, inds = np.ascontiguousarray(m.offs, dtype=np.int32), np.ascontiguousarray(m.inds, dtype=np.int32)
        vals, c = np.ascontiguousarray(m.vals, dtype=np.float64), np.ascontiguousarray(m.c, dtype=np.float64)
        sel, ns, obj = np.zeros(m.n_orig, dtype=np.int32), C.c_int(), C.c_double()
        feas = lib.ref_greedy_set_cover(m.m, m.n_orig, offs.ctypes.data_as(PI), inds.ctypes.data_as(PI), vals.ctypes.data_as(PD),
                                        c.ctypes.data_as(PD), sel.ctypes.data_as(PI), C.byref(ns), C.byref(obj))
        return bool(feas), obj.value, set(sel[:ns.value].tolist())

    # no ties: real-valued costs -> the same columns in the same scan order
    for seed in range(4):
        mdl = gen_scp(40, 300, 0.06, seed)
        rng = np.random.default_rng(seed)
        mdl.c[:mdl.n_orig] = rng.uniform(1.0, 100.0, mdl.n_orig)
        feas, obj, sel = reference(mdl)
        o, x = bnb.greedy_cover(mdl)
        assert feas and x is not None and sel == set(np.nonzero(x)[0].tolist())
        assert abs(o - obj) <= 1e-12 * obj
    # the instances the bench runs its B&B on: same first incumbent, hence the same reduced model
    golden = Path(__file__).resolve().parent / "golden"
    for name, want in (("scpnre1", 38.0), ("scpnrg1", 266.0)):
        mdl = load_npz(golden / f"{name}.npz")
        feas, obj, _ = reference(mdl)
        o, x = bnb.greedy_cover(mdl)
        assert feas and obj == want and o == want
        red_a, _ = bnb.reduce_by_incumbent(mdl, o)
        red_b, _ = bnb.reduce_by_incumbent(mdl, obj)
        assert (red_a.n_orig, red_a.nnz) == (red_b.n_orig, red_b.nnz)
    # integer costs with ties: both are covers whose cost is what they report
    for name in ("scp41", "scpclr10", "scpnrh1"):
        mdl = load_npz(golden / f"{name}.npz")
        h = bnb.CoverHeuristic(mdl)
        feas, obj, sel = reference(mdl)
        o, x = bnb.greedy_cover(mdl)
        xr = np.zeros(mdl.n_orig)
        xr[sorted(sel)] = 1.0
        assert feas and np.all(h.A @ xr >= 1.0) and abs(float(mdl.c[:mdl.n_orig] @ xr) - obj) < 1e-9
        assert x is not None and np.all(h.A @ x >= 1.0) and abs(float(mdl.c[:mdl.n_orig] @ x) - o) < 1e-9
