"""GPU: the batched B&B node loop finds the integer optimum of small set-covering instances (checked
against SciPy's MILP solver) and batches node LPs of different depth."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from sypha_b200 import bnb  # noqa: E402
from sypha_b200.instances import gen_scp  # noqa: E402


def _milp_optimum(mdl):
    import scipy.sparse as sp
    from scipy.optimize import Bounds, LinearConstraint, milp
    A = sp.csr_matrix((mdl.vals, mdl.inds, mdl.offs), shape=(mdl.m, mdl.n))[:, :mdl.n_orig]
    r = milp(mdl.c[:mdl.n_orig], constraints=LinearConstraint(A, lb=1.0), integrality=np.ones(mdl.n_orig),
             bounds=Bounds(0, 1))
    assert r.success
    return float(round(r.fun))


@pytest.mark.parametrize("shape", [(20, 60, 0.15, 1), (30, 120, 0.1, 2), (40, 200, 0.08, 5)])
def test_batched_bnb_reaches_the_milp_optimum(shape):
    m, n, dens, seed = shape
    mdl = gen_scp(m, n, dens, seed)
    opt = _milp_optimum(mdl)
    drv = bnb.BatchedBnb(mdl, slots=4)
    try:
        st = drv.run(max_nodes=4000)
        assert st.open_nodes == 0, "search did not finish"
        assert st.incumbent == opt, (st.incumbent, opt, st)
        assert st.root_bound <= opt * (1 + 1e-4)
        assert np.all(drv.heur.A @ drv.incumbent_x >= 1.0)
        assert st.processed >= 1 and st.lp_iterations > 0
    finally:
        drv.close()


def test_slot_count_does_not_change_the_answer():
    mdl = gen_scp(30, 120, 0.1, 9)
    res = []
    for slots in (1, 3, 8):
        drv = bnb.BatchedBnb(mdl, slots=slots)
        try:
            res.append(drv.run(max_nodes=4000).incumbent)
        finally:
            drv.close()
    assert res[0] == res[1] == res[2] == _milp_optimum(mdl)


def test_device_node_delta_matches_full_upload():
    """sb200_set_node_delta (base model resident, branch rows folded on the device) must give the same LP as
    building the node CSR on the host and uploading it (build_branch_model + copyModelOnDevice)."""
    import sypha_b200 as sb
    from sypha_b200 import solver as S
    mdl = gen_scp(120, 900, 0.04, 11)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    wss = [S.workspace_for_nodes(base, 70) for _ in range(3)]
    try:
        decs = [((5, 1), (17, 0), (400, 1)), (), tuple((3 * i + 1, i % 2) for i in range(66))]
        got = S.solve_batch_nodes(base, decs, cfg, wss)
        for dec, g in zip(decs, got):
            nm = bnb.build_branch_model(mdl, dec)
            node = sb.SyphaNodeSparse.from_csr(nm.m, nm.n, nm.n_orig, nm.offs, nm.inds, nm.vals, nm.c, nm.b, env)
            ref = sb.SolverExecutionResult()
            sb.solver_sparse_mehrotra_run(node, cfg, ref)
            assert g.terminationReason == ref.terminationReason
            assert g.iterations == ref.iterations, (len(dec), g.iterations, ref.iterations)
            assert abs(g.primalObj - ref.primalObj) <= 1e-9 * max(1, abs(ref.primalObj))
            assert abs(g.dualObj - ref.dualObj) <= 1e-9 * max(1, abs(ref.dualObj))
            assert g.primalSolution.shape == ref.primalSolution.shape
            assert np.max(np.abs(g.primalSolution - ref.primalSolution)) <= 1e-7 * (1 + np.abs(ref.primalSolution).max())
        # going back to a shallower node / the base on the same workspace
        again = S.solve_batch_nodes(base, [(), decs[0], ()], cfg, wss)
        assert again[0].iterations == got[1].iterations and abs(again[0].primalObj - got[1].primalObj) < 1e-12
        assert again[1].iterations == got[0].iterations and abs(again[1].primalObj - got[0].primalObj) < 1e-12
    finally:
        for w in wss:
            sb.releaseIpmWorkspace(w)


def test_host_and_device_node_paths_search_the_same_tree():
    mdl = gen_scp(30, 120, 0.1, 2)
    out = []
    for dev_nodes in (True, False):
        drv = bnb.BatchedBnb(mdl, slots=4, device_nodes=dev_nodes)
        try:
            st = drv.run(max_nodes=4000)
            out.append((st.incumbent, st.processed, st.lp_iterations))
        finally:
            drv.close()
    assert out[0] == out[1]


def _check_node_heuristics(mdl, decs, max_iter):
    """sb200_node_heuristics (device) against bnb.CoverHeuristic + the NumPy branching rule (host) on the
    same LP points: integer work, so everything must be identical."""
    import sypha_b200 as sb
    from sypha_b200 import solver as S
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=max_iter)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    wss = [S.workspace_for_nodes(base, 70) for _ in decs]
    host = bnb.CoverHeuristic(mdl)
    n0 = mdl.n_orig
    steps = []
    try:
        for w in wss:
            S.set_heuristic_rules(w, "plain")
        S.solve_batch_nodes(base, decs, cfg, wss, fetch_solutions=False)
        got = S.node_heuristics(wss)
        for dec, ws, h in zip(decs, wss, got):
            x = S.get_primal(ws, mdl.n + len(dec))[:n0]
            obj, cover = host(x, [v for v, f in dec if f == 0])
            if np.all(np.isfinite(x)):                 # (an infeasible node's LP point is not finite)
                frac = np.abs(x - np.round(x))
                assert h.branchVar == int(np.argmax(frac))
                assert h.branchFrac == frac[h.branchVar]
                assert abs(h.roundedObj - float(mdl.c[:n0] @ np.round(x))) <= 1e-9 * max(1.0, abs(h.roundedObj))
            else:
                assert h.branchVar == -1
            if cover is None:
                assert not h.feasible
                continue
            assert h.feasible
            assert h.coverObj == obj, (h, obj)
            dev_cover = S.get_cover(ws, n0)
            assert np.array_equal(dev_cover, cover)
            assert h.nChosen == int(cover.sum())
            assert np.all(host.A @ dev_cover >= 1.0)
            steps.append(h.repairSteps)
    finally:
        for w in wss:
            sb.releaseIpmWorkspace(w)
    return steps


def test_node_heuristics_kernel_matches_the_host_rules_at_the_lp_optimum():
    mdl = gen_scp(120, 900, 0.04, 11)
    decs = [(), ((5, 1), (17, 0), (400, 1)), tuple((3 * i + 1, i % 2) for i in range(40)), ((7, 0),)]
    _check_node_heuristics(mdl, decs, 100)


@pytest.mark.parametrize("max_iter", [1, 3])
def test_node_heuristics_kernel_long_repairs(max_iter):
    """An LP point far from the optimum (1 or 3 IPM iterations) rounds to almost nothing, so the cover is built
    by the greedy repair alone: many picks, every incremental gain update exercised; unit costs give ties."""
    mdl = gen_scp(200, 3000, 0.02, 4)
    decs = [(), tuple((11 * i, 0) for i in range(30)), ((1, 1), (2, 0))]
    steps = _check_node_heuristics(mdl, decs, max_iter)
    ties = gen_scp(150, 1200, 0.03, 8)
    ties.c[:ties.n_orig] = 1.0 + (np.arange(ties.n_orig) % 3)
    steps += _check_node_heuristics(ties, [(), ((0, 0), (1, 0), (2, 0))], max_iter)
    assert max(steps) >= 10, steps


def test_node_heuristics_kernel_reports_infeasible_fixings():
    mdl = gen_scp(40, 200, 0.08, 5)
    row0 = mdl.inds[mdl.offs[0]:mdl.offs[1]]
    ban = tuple((int(j), 0) for j in row0 if j < mdl.n_orig)       # every column of row 0 fixed to 0
    assert len(ban) <= 64
    _check_node_heuristics(mdl, [ban, ()], 100)


def _check_reference_rules(mdl, decs, max_iter, branch_rule="most_fractional"):
    """sb200_node_heuristics with the reference's rules (the default) against the CPU restatement of
    sypha_solver_heuristics.cpp (oracle/heur_oracle.c, itself pinned to the reference's own translation unit by
    tests/test_heuristics_oracle.py) on the same LP points: every decision and both covers identical."""
    import sypha_b200 as sb
    from oracle import heuristics as H
    from oracle import scp_io
    from sypha_b200 import _lib as L, solver as S
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=max_iter)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    inst = scp_io.ScpInstance(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b)
    wss = [S.workspace_for_nodes(base, 70) for _ in decs]
    n0 = mdl.n_orig
    steps = []
    try:
        for w in wss:
            S.set_heuristic_rules(w, "reference", branch_rule, 1e-6)
        S.solve_batch_nodes(base, decs, cfg, wss, fetch_solutions=False)
        got = S.node_heuristics(wss)
        for dec, ws, h in zip(decs, wss, got):
            k = len(dec)
            x, y = np.empty(mdl.n + k), np.empty(mdl.m + k)
            assert L.load().sb200_get_iterates(ws.handle, x.ctypes.data, y.ctypes.data, None) == L.SB200_OK
            if not np.all(np.isfinite(x)):
                continue
            j, frac = H.select_branch(x, mdl.c, n0, 1e-6, branch_rule)
            assert h.branchVar == j and h.branchFrac == frac
            f, obj, sol = H.nearest_integer_fixing(inst, x, dec)
            assert h.nifFeasible == f
            assert np.array_equal(S.get_rounded(ws, n0), sol)
            if f:
                assert h.nifObj == obj
            f, obj, sol, st = H.dual_guided_cover_repair(inst, x, y[:mdl.m], dec)
            assert h.feasible == f, (h, f, obj)
            if f:
                assert h.coverObj == obj and h.repairSteps == st and h.nChosen == int(sol.sum())
                assert np.array_equal(S.get_cover(ws, n0), sol)
            steps.append(st)
    finally:
        for w in wss:
            sb.releaseIpmWorkspace(w)
    return steps


def test_reference_rules_kernel_at_the_lp_optimum():
    mdl = gen_scp(120, 900, 0.04, 11)
    decs = [(), ((5, 1), (17, 0), (400, 1)), tuple((3 * i + 1, i % 2) for i in range(40)), ((7, 0),)]
    _check_reference_rules(mdl, decs, 100)
    _check_reference_rules(mdl, decs[:2], 100, "highest_cost_fractional")


@pytest.mark.parametrize("max_iter", [2, 5, 12])
def test_reference_rules_kernel_long_repairs_and_dual_guidance(max_iter):
    """LP points 2, 5 and 12 iterations old: few columns at 1, so the dual-guided repair builds most of the cover
    (scores (rows + sum of duals) / cost with real duals), plus tied unit-ish costs and real instances."""
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    mdl = gen_scp(200, 3000, 0.02, 4)
    steps = _check_reference_rules(mdl, [(), tuple((11 * i, i % 2) for i in range(30)), ((1, 1), (2, 0))], max_iter)
    ties = gen_scp(150, 1200, 0.03, 8)
    ties.c[:ties.n_orig] = 1.0 + (np.arange(ties.n_orig) % 3)
    steps += _check_reference_rules(ties, [(), ((0, 0), (1, 0), (2, 0))], max_iter)
    steps += _check_reference_rules(load_npz(GOLDEN / "scp41.npz"), [(), ((10, 0), (200, 1))], max_iter)
    red, _ = bnb.reduce_by_incumbent(load_npz(GOLDEN / "scpnre1.npz"), 38.0)
    steps += _check_reference_rules(red, [(), ((3, 0), (900, 1), (1700, 0))], max_iter)
    assert max(steps) >= 8, steps


def test_reference_rules_kernel_reports_infeasible_fixings():
    mdl = gen_scp(40, 200, 0.08, 5)
    row0 = mdl.inds[mdl.offs[0]:mdl.offs[1]]
    ban = tuple((int(j), 0) for j in row0 if j < mdl.n_orig)       # every column of row 0 fixed to 0
    _check_reference_rules(mdl, [ban, ()], 100)


@pytest.mark.parametrize("shape", [(30, 120, 0.1, 2), (40, 200, 0.08, 5)])
def test_continuous_batching_reaches_the_milp_optimum(shape):
    """sb200_solve_stream: slots take the next open node as soon as they are free; the search order then depends
    on completion times, the optimum does not."""
    m, n, dens, seed = shape
    mdl = gen_scp(m, n, dens, seed)
    opt = _milp_optimum(mdl)
    drv = bnb.BatchedBnb(mdl, slots=4)
    try:
        st = drv.run(max_nodes=10 ** 6, stream_nodes=64)
        assert st.open_nodes == 0, "search did not finish"
        assert st.incumbent == opt, (st.incumbent, opt, st)
        assert np.all(drv.heur.A @ drv.incumbent_x >= 1.0)
        assert st.processed >= 1 and st.lp_iterations > 0 and st.kernels_launched > 0
    finally:
        drv.close()


def test_concurrency_hint_changes_the_launch_geometry_not_the_result():
    """sb200_set_concurrency_hint caps the CTAs of the data-flow factorisation (B&B slots share the GPU): the
    arithmetic is the same, so the LP must come out bit-identical."""
    import sypha_b200 as sb
    from sypha_b200 import _lib as L, solver as S
    mdl = gen_scp(300, 2500, 0.04, 21)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    ws = S.workspace_for_nodes(base, 8)
    try:
        out = []
        for hint in (0, 32, 0):
            assert L.load().sb200_set_concurrency_hint(ws.handle, hint) == L.SB200_OK
            S.set_solver_form(ws, "latency")            # the hint alone would also switch to one thread block per LP
            r = S.solve_batch_nodes(base, [((3, 1), (40, 0))], cfg, [ws])[0]
            out.append((r.iterations, r.primalObj, r.dualObj, r.primalSolution.copy()))
        for o in out[1:]:
            assert o[0] == out[0][0] and o[1] == out[0][1] and o[2] == out[0][2]
            assert np.array_equal(o[3], out[0][3])
    finally:
        sb.releaseIpmWorkspace(ws)


# ---- the real instances, against the integer optima the reference holds ------------------------------------
@pytest.mark.parametrize("name,node_lp", [("scp41", "reference"), ("scp48", "reference"), ("scp410", "reference"),
                                          ("scp48", "converged"), ("scp42", "reference"), ("scp46", "converged")])
def test_batched_bnb_reaches_the_reference_held_ip_optimum(name, node_lp):
    """scp41 429, scp48 492, scp410 514 ... (benchmark/results/benchmark_results_with_ip.csv:4,5,12 via
    tests/golden/ip_optima.json), with the reference's node LP configuration (gap-stagnation exit, window 5, 1 %)
    and with node LPs run to convergence."""
    import json
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    gold = json.load(open(GOLDEN / "ip_optima.json"))[name]
    mdl = load_npz(GOLDEN / f"{name}.npz")
    drv = bnb.BatchedBnb(mdl, slots=8, node_lp=node_lp)
    try:
        st = drv.run(max_nodes=20000)
        assert st.open_nodes == 0, f"search did not finish: {st}"
        assert st.incumbent == gold, (st.incumbent, gold)
        x = drv.incumbent_x
        assert np.all(drv.heur.A @ x >= 1.0) and float(mdl.c[:mdl.n_orig] @ x) == gold
        assert st.root_bound <= gold * (1 + 1e-4)          # the LP dual objective at mu <= 1e-4
    finally:
        drv.close()


def test_node_at_the_iteration_cap_is_kept_and_branched_on():
    """a node LP that stops at max_iter is not a failed LP - the reference bounds it with its parent's
    bound and branches (bnb_driver.cpp:866-877).  On scp48 (root LP 14 iterations, node LPs 8-15) a cap in between
    makes SOME node LPs stop at the limit; the search must still end at the optimum the reference holds (492): no
    subtree may be dropped."""
    import json
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    gold = json.load(open(GOLDEN / "ip_optima.json"))["scp48"]
    mdl = load_npz(GOLDEN / "scp48.npz")
    hit = finished = 0
    for cap in range(8, 16):
        drv = bnb.BatchedBnb(mdl, slots=8, max_iter=cap, node_lp="converged")
        try:
            st = drv.run(max_nodes=4000)
            if st.open_nodes:                      # too few converged LPs to bound the tree within the node budget
                continue
            finished += 1
            assert st.infeasible == 0
            assert st.incumbent == gold, (cap, st.incumbent, gold, st)
            hit += st.maxiter_nodes
        finally:
            drv.close()
    assert finished > 0 and hit > 0, (finished, hit)


# ---- windows in flight (sb200_window_begin / sb200_window_finish) ------------------------------------------------
def test_window_begin_finish_equals_the_synchronous_window():
    """The two halves of a window against sb200_solve_batch + sb200_node_heuristics on the same nodes: same kernels,
    so the same numbers bit for bit; two windows over two workspace sets in flight at once."""
    import sypha_b200 as sb
    from sypha_b200 import solver as S
    mdl = gen_scp(60, 400, 0.06, 3)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    decs_a = [(), ((5, 1),), ((5, 0), (17, 1)), ((9, 1), (30, 0), (44, 1))]
    decs_b = [((3, 0),), ((3, 1), (8, 0)), ((100, 1),), ()]
    sets = [[S.workspace_for_nodes(base, 16) for _ in range(4)] for _ in range(3)]
    try:
        for st in sets:
            for w in st:
                S.set_solver_form(w, "throughput")
                S.set_heuristic_rules(w, "reference", "most_fractional", 1e-6)
        ref = []
        for decs in (decs_a, decs_b):
            r = S.solve_batch_nodes(base, decs, cfg, sets[2], fetch_solutions=False, fetch_trace=False)
            ref.append((r, S.node_heuristics(sets[2])))
        wa = S.window_begin(base, decs_a, cfg, sets[0])
        wb = S.window_begin(base, decs_b, cfg, sets[1])            # queued behind / beside the first
        assert wa is not None and wb is not None
        for w, (r_ref, h_ref) in ((wa, ref[0]), (wb, ref[1])):
            res, rules = S.window_finish(w)
            for a, b in zip(res, r_ref):
                assert (a.status, a.terminationReason, a.iterations) == (b.status, b.terminationReason, b.iterations)
                assert a.primalObj == b.primalObj and a.dualObj == b.dualObj
            for a, b in zip(rules, h_ref):
                assert a == b
        ms, k = S.last_window(sets[1][0])
        assert k == 4 and ms > 0
    finally:
        for st in sets:
            for w in st:
                sb.releaseIpmWorkspace(w)


@pytest.mark.parametrize("name", ["scp41", "scp48", "scp410"])
def test_pipelined_windows_reach_the_reference_held_ip_optimum(name):
    """BatchedBnb(pipeline=2): two windows in flight over alternating workspace sets; the frontier order differs from
    the one-window search by the lag of one window, the optimum does not."""
    import json
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    gold = json.load(open(GOLDEN / "ip_optima.json"))[name]
    mdl = load_npz(GOLDEN / f"{name}.npz")
    drv = bnb.BatchedBnb(mdl, slots=8, pipeline=2)
    try:
        assert drv.pipeline == 2
        st = drv.run(max_nodes=20000)
        assert st.open_nodes == 0 and not drv._inflight, f"search did not finish: {st}"
        assert st.incumbent == gold, (st.incumbent, gold)
        x = drv.incumbent_x
        assert np.all(drv.heur.A @ x >= 1.0) and float(mdl.c[:mdl.n_orig] @ x) == gold
    finally:
        drv.close()
