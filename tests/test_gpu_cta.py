"""GPU: the throughput form of the solve - the whole LP by ONE thread block in ONE launch (csrc/sb200_cta.cu) -
against the golden fixtures (the reference's own outputs) and against the latency form (the multi-kernel path) on
the same inputs: same iteration counts, objectives inside the north_star contract (1e-6; in practice ~1e-10, the
two forms differ only in summation order), same termination behaviour, B&B node models included."""
import numpy as np
import pytest

from conftest import load_golden, node_from_instance

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import sypha_b200 as sb  # noqa: E402
from sypha_b200 import bnb, solver as S  # noqa: E402
from sypha_b200.instances import gen_scp  # noqa: E402

REL = 1e-6


def solve(inst, ws, form, cfg=None, **env_kw):
    S.set_solver_form(ws, form)
    node = node_from_instance(inst, linearSolverStrategy="cholesky", **env_kw)
    res = sb.SolverExecutionResult()
    st = sb.solver_sparse_mehrotra_run(node, cfg or sb.SolverExecutionConfig(maxIterations=100), res, ws)
    return st, res


@pytest.fixture()
def ws():
    w = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(w)
    yield w
    S.set_solver_form(w, "latency")
    sb.releaseIpmWorkspace(w)


@pytest.mark.parametrize("name", ["demo00", "scp_demo06", "scp41", "scp48", "scp51", "scpcyc06", "scpa1", "scpb1", "scpclr10",
                                  "scpnre1", "scpnrf1", "scpnrg1", "scpnrh1", "scpnrh4"])
def test_one_cta_solver_matches_the_reference_outputs(name, ws):
    inst, z = load_golden(name)
    st, r = solve(inst, ws, "throughput")
    assert st == sb.CODE_SUCCESSFUL and r.terminationReason == sb.SOLVER_TERM_CONVERGED
    assert r.kernelsLaunched <= 3, r.kernelsLaunched            # the LP really was one launch
    assert r.iterations == int(z["ref_iters"])
    assert abs(r.primalObj - float(z["ref_primal"])) <= REL * max(1, abs(float(z["ref_primal"])))
    assert abs(r.dualObj - float(z["ref_dual"])) <= REL * max(1, abs(float(z["ref_dual"])))
    assert r.mu <= 1e-4
    mu_tr = z["oracle_mu_trace"]
    k = min(len(mu_tr), r.trace.shape[0])
    assert np.allclose(r.trace[:k, 1], mu_tr[:k], rtol=1e-6)
    assert np.max(np.abs(r.primalSolution - z["ref_x"]) / (1 + np.abs(z["ref_x"]))) < 1e-5
    assert np.max(np.abs(r.dualSolution - z["ref_y"]) / (1 + np.abs(z["ref_y"]))) < 1e-5
    # and the latency form on the same workspace
    st2, r2 = solve(inst, ws, "latency")
    assert r2.iterations == r.iterations and r2.kernelsLaunched > 10
    assert abs(r.primalObj - r2.primalObj) <= 1e-8 * max(1, abs(r2.primalObj))
    assert abs(r.dualObj - r2.dualObj) <= 1e-8 * max(1, abs(r2.dualObj))


def test_termination_rules_are_the_same_in_both_forms(ws):
    inst, _ = load_golden("scp41")
    for cfg in (sb.SolverExecutionConfig(maxIterations=7),
                sb.SolverExecutionConfig(maxIterations=100, gapStagnation=sb.SolverGapStagnationConfig(True, 2, 60.0)),
                sb.SolverExecutionConfig(maxIterations=0)):
        a = solve(inst, ws, "throughput", cfg)[1]
        b = solve(inst, ws, "latency", cfg)[1]
        assert (a.terminationReason, a.iterations) == (b.terminationReason, b.iterations)
        assert abs(a.primalObj - b.primalObj) <= 1e-8 * max(1, abs(b.primalObj))
        assert abs(a.dualObj - b.dualObj) <= 1e-8 * max(1, abs(b.dualObj))


def test_node_models_and_infeasible_fixings_in_both_forms():
    """Base model resident, branch rows folded on the device (sb200_set_node_delta): depth 0, 3 and 40, and a node
    whose fixings leave a row uncoverable."""
    mdl = gen_scp(120, 900, 0.04, 11)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(mdl.m, mdl.n, mdl.n_orig, mdl.offs, mdl.inds, mdl.vals, mdl.c, mdl.b, env)
    row0 = mdl.inds[mdl.offs[0]:mdl.offs[1]]
    ban = tuple((int(j), 0) for j in row0 if j < mdl.n_orig)
    decs = [(), ((5, 1), (17, 0), (400, 1)), tuple((3 * i + 1, i % 2) for i in range(40)), ban]
    wss = [S.workspace_for_nodes(base, 70) for _ in decs]
    try:
        out = {}
        for form in ("latency", "throughput"):
            for w in wss:
                S.set_solver_form(w, form)
            out[form] = S.solve_batch_nodes(base, decs, cfg, wss)
        for a, b in zip(out["throughput"], out["latency"]):
            assert a.status == b.status and a.terminationReason == b.terminationReason
            if a.status == sb.CODE_SUCCESSFUL:
                assert a.iterations == b.iterations
                assert abs(a.primalObj - b.primalObj) <= 1e-8 * max(1, abs(b.primalObj))
                assert abs(a.dualObj - b.dualObj) <= 1e-8 * max(1, abs(b.dualObj))
                assert np.max(np.abs(a.primalSolution - b.primalSolution)) <= 1e-6 * (1 + np.abs(b.primalSolution).max())
        assert out["throughput"][3].status != sb.CODE_SUCCESSFUL or out["throughput"][3].terminationReason != sb.SOLVER_TERM_CONVERGED
    finally:
        for w in wss:
            sb.releaseIpmWorkspace(w)


def test_pattern_only_products_match_the_value_carrying_products():
    """The one-block solver's A v / A' v over the 2-byte pattern lists of the base model (CompactLists, the default where
    every column is all +1 or all -1) against the same solver reading the 12-byte CSR / CSC entries
    (SB200_COMPACT_PRODUCTS=0): root LP and node LPs with branch rows, including two rows on one variable."""
    import os
    inst, _ = load_golden("scpnre1")
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    decs = [(), ((5, 1), (17, 0), (400, 1)), tuple((7 * i + 2, (i + 1) % 2) for i in range(24)), ((9, 1), (9, 1), (30, 0))]
    out = {}
    for flag in ("1", "0"):
        os.environ["SB200_COMPACT_PRODUCTS"] = flag
        try:
            base = sb.SyphaNodeSparse.from_csr(inst.m, inst.n, inst.n_orig, inst.offs, inst.inds, inst.vals, inst.c, inst.b, env)
            wss = [S.workspace_for_nodes(base, 32) for _ in decs]
            for w in wss:
                S.set_solver_form(w, "throughput")
            out[flag] = S.solve_batch_nodes(base, decs, cfg, wss)
            for w in wss:
                sb.releaseIpmWorkspace(w)
        finally:
            os.environ.pop("SB200_COMPACT_PRODUCTS", None)
    for dec, a, b in zip(decs, out["1"], out["0"]):
        assert a.status == b.status == sb.CODE_SUCCESSFUL and a.terminationReason == b.terminationReason
        assert a.iterations == b.iterations
        assert abs(a.primalObj - b.primalObj) <= 1e-9 * max(1, abs(b.primalObj))
        assert abs(a.dualObj - b.dualObj) <= 1e-9 * max(1, abs(b.dualObj))
        assert np.max(np.abs(a.primalSolution - b.primalSolution)) <= 1e-7 * (1 + np.abs(b.primalSolution).max())
        # duals: the base rows only.  A "fix to 0" row (-x_j - t = 0, x_j, t >= 0) has no interior, its dual grows like
        # 1 / mu and is not a number two summation orders agree on; the same bound twice is a degenerate pair
        m0 = inst.m
        assert np.max(np.abs(a.dualSolution[:m0] - b.dualSolution[:m0])) <= 1e-6 * (1 + np.abs(b.dualSolution[:m0]).max())


def test_forms_search_the_same_tree():
    """BatchedBnb over one-CTA node LPs (the default with several slots) and over the multi-kernel form."""
    import os
    mdl = gen_scp(30, 120, 0.1, 2)
    out = []
    for cta in ("1", "0"):
        os.environ["SB200_CTA_SOLVER"] = cta
        try:
            drv = bnb.BatchedBnb(mdl, slots=4)
            st = drv.run(max_nodes=4000)
            out.append((st.incumbent, st.open_nodes))
            drv.close()
        finally:
            os.environ.pop("SB200_CTA_SOLVER", None)
    assert out[0] == out[1] and out[0][1] == 0


def test_warm_started_children_need_fewer_iterations():
    """a child LP started from its parent's final iterate, floored at 0.1 (sb200_node_delta.warm_start),
    against the same LP from the Mehrotra starting point - on the model the reference's B&B really solves for scpnre1
    (500 x 1775 after its reductions).  Both runs converge (mu <= 1e-4, dual <= primal); exit objectives of two different
    trajectories agree to the duality gap the mu-only stop test leaves (n mu / |obj|), not to 1e-6."""
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    red, _ = bnb.reduce_by_incumbent(load_npz(GOLDEN / "scpnre1.npz"), 38.0)
    env = sb.SyphaEnvironment()
    cfg = sb.SolverExecutionConfig(maxIterations=100)
    base = sb.SyphaNodeSparse.from_csr(red.m, red.n, red.n_orig, red.offs, red.inds, red.vals, red.c, red.b, env)
    wss = [S.workspace_for_nodes(base, 16) for _ in range(4)]
    try:
        for w in wss:
            S.set_solver_form(w, "throughput")
        n, m = red.n, red.m
        exp0 = torch.empty(2 * n + m, dtype=torch.float64, device="cuda")
        root = S.solve_batch_nodes(base, [()], cfg, wss[:1], export=[exp0.data_ptr()])[0]
        assert root.terminationReason == sb.SOLVER_TERM_CONVERGED
        assert np.array_equal(exp0[:n].cpu().numpy(), root.primalSolution)
        x0 = root.primalSolution[:red.n_orig]
        j = int(np.argmax(np.abs(x0 - np.round(x0))))
        decs = [((j, 0),), ((j, 1),)]
        cold = S.solve_batch_nodes(base, decs, cfg, wss[:2])
        exps = [torch.empty(2 * (n + 1) + m + 1, dtype=torch.float64, device="cuda") for _ in decs]
        warm = S.solve_batch_nodes(base, decs, cfg, wss[2:4], warm=[(exp0.data_ptr(), n, m)] * 2,
                                   export=[e.data_ptr() for e in exps])
        total_c = total_w = 0
        for c, w_ in zip(cold, warm):
            assert c.terminationReason == w_.terminationReason == sb.SOLVER_TERM_CONVERGED
            assert w_.mu <= 1e-4 and w_.dualObj <= w_.primalObj + 1e-9
            gap = 2.0 * (n + 1) * 1e-4
            assert abs(w_.primalObj - c.primalObj) <= gap and abs(w_.dualObj - c.dualObj) <= gap
            total_c += c.iterations
            total_w += w_.iterations
        assert total_w <= 0.8 * total_c, (total_w, total_c)
        # second level: children of the x_j = 1 child, from ITS exported iterate
        x1 = warm[1].primalSolution[:red.n_orig]
        j2 = int(np.argmax(np.abs(x1 - np.round(x1))))
        decs2 = [((j, 1), (j2, 0)), ((j, 1), (j2, 1))]
        cold2 = S.solve_batch_nodes(base, decs2, cfg, wss[:2])
        warm2 = S.solve_batch_nodes(base, decs2, cfg, wss[2:4], warm=[(exps[1].data_ptr(), n + 1, m + 1)] * 2)
        assert sum(r.iterations for r in warm2) <= 0.85 * sum(r.iterations for r in cold2)
        assert all(r.terminationReason == sb.SOLVER_TERM_CONVERGED for r in warm2)
    finally:
        for w in wss:
            sb.releaseIpmWorkspace(w)


@pytest.mark.parametrize("name", ["scp41", "scp48", "scp410"])
def test_warm_started_bnb_reaches_the_ip_optimum(name):
    import json
    from conftest import GOLDEN
    from sypha_b200.instances import load_npz
    gold = json.load(open(GOLDEN / "ip_optima.json"))[name]
    drv = bnb.BatchedBnb(load_npz(GOLDEN / f"{name}.npz"), slots=8, warm_start=True)
    try:
        st = drv.run(max_nodes=20000)
        assert st.open_nodes == 0 and st.incumbent == gold, st
    finally:
        drv.close()
