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
