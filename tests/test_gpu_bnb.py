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
