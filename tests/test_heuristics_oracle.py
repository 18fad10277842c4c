"""CPU: oracle/heur_oracle.c (the restatement the device kernel is checked against) pinned to the reference's OWN
integer heuristics and branching selectors - src/sypha_solver_heuristics.cpp compiled as it lies into
oracle/_ref/libref_heur.so - on random set-covering instances with LP-like points, branching decisions, dual
guidance, distinct and tied costs, infeasible fixings."""
import numpy as np
import pytest

from oracle import heuristics as H
from oracle import scp_io

REF = H.reference_lib()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/libref_heur.so not built (make -C oracle)")


def lp_like_point(inst, rng, p_one=0.08, p_frac=0.15):
    """x with a few ones, some fractional entries, the rest near zero; y >= 0 mostly, some negative."""
    n0 = inst.n_orig
    x = np.abs(rng.normal(0, 1e-7, inst.n))
    u = rng.random(n0)
    x[:n0][u < p_one] = 1.0 - np.abs(rng.normal(0, 1e-8, (u < p_one).sum()))
    fr = (u >= p_one) & (u < p_one + p_frac)
    x[:n0][fr] = rng.random(fr.sum())
    y = rng.gamma(1.0, 2.0, inst.m) * (rng.random(inst.m) < 0.8) - 0.1 * (rng.random(inst.m) < 0.1)
    return x, y


def distinct_costs(inst, rng):
    inst.c[:inst.n_orig] = rng.permutation(inst.n_orig) + 1.0 + rng.random(inst.n_orig) * 0.5
    return inst


def decisions_for(inst, rng, k):
    vars_ = rng.choice(inst.n_orig, size=k, replace=False)
    return [(int(v), int(rng.integers(0, 2))) for v in vars_]


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_dual_guided_cover_repair_equals_the_reference(seed):
    rng = np.random.default_rng(seed)
    inst = distinct_costs(scp_io.gen_scp(int(rng.integers(15, 60)), int(rng.integers(60, 300)), 0.06, seed), rng)
    x, y = lp_like_point(inst, rng)
    dec = decisions_for(inst, rng, int(rng.integers(0, 6)))
    f0, o0, s0 = H.reference_heuristic("dual_guided_cover_repair", inst, x, y, dec)
    f1, o1, s1, steps = H.dual_guided_cover_repair(inst, x, y, dec)
    assert f0 == f1
    if f0:
        assert o0 == o1 and np.array_equal(s0, s1)
        A0 = inst.csr()[:, :inst.n_orig]
        assert np.all(A0 @ s1 >= 1.0)
        assert all(s1[v] == f for v, f in dec)


@needs_ref
@pytest.mark.parametrize("seed", range(8))
def test_nearest_integer_fixing_equals_the_reference(seed):
    rng = np.random.default_rng(100 + seed)
    inst = scp_io.gen_scp(20, 80, 0.15, seed)
    x, y = lp_like_point(inst, rng, p_one=0.25 if seed % 2 else 0.5, p_frac=0.2)
    dec = decisions_for(inst, rng, seed % 4)
    f0, o0, s0 = H.reference_heuristic("nearest_integer_fixing", inst, x, y, dec)
    f1, o1, s1 = H.nearest_integer_fixing(inst, x, dec)
    assert f0 == f1 and np.array_equal(s0, s1)
    if f0:
        assert o0 == o1


@needs_ref
def test_infeasible_fixings_and_tied_costs():
    inst = scp_io.gen_scp(12, 30, 0.2, 3)
    A0 = inst.csr()[:, :inst.n_orig].toarray()
    row = 4
    dec = [(int(j), 0) for j in np.nonzero(A0[row])[0]]          # every column of one row fixed to 0
    x, y = lp_like_point(inst, np.random.default_rng(0))
    assert H.reference_heuristic("dual_guided_cover_repair", inst, x, y, dec)[0] is False
    assert H.dual_guided_cover_repair(inst, x, y, dec)[0] is False
    # tied (integer 1..3) costs: the reference's std::sort leaves the order of equal costs open, so only the
    # properties are compared - both covers feasible and irredundant among their non-fixed columns
    rng = np.random.default_rng(7)
    inst = scp_io.gen_scp(25, 120, 0.08, 11)
    inst.c[:inst.n_orig] = rng.integers(1, 4, inst.n_orig)
    x, y = lp_like_point(inst, rng)
    A0 = inst.csr()[:, :inst.n_orig]
    for f, o, s in (H.reference_heuristic("dual_guided_cover_repair", inst, x, y)[:3], H.dual_guided_cover_repair(inst, x, y)[:3]):
        assert f and np.all(A0 @ s >= 1.0) and o == inst.c[:inst.n_orig] @ s
        for j in np.nonzero(s)[0]:
            t = s.copy()
            t[j] = 0
            assert np.any(A0 @ t < 1.0)


@needs_ref
@pytest.mark.parametrize("rule", ["most_fractional", "highest_cost_fractional"])
def test_branch_selectors_equal_the_reference(rule):
    rng = np.random.default_rng(5)
    for trial in range(30):
        n0 = int(rng.integers(5, 200))
        x = np.round(rng.random(n0), 1 if trial % 3 == 0 else 6)       # ties in the fractional part
        x[rng.random(n0) < 0.5] = np.round(x[rng.random(n0) < 0.5].shape[0] and 1.0)
        if trial % 7 == 0:
            x = np.round(x)                                           # integral point: no candidate
        obj = rng.integers(1, 10, n0).astype(float)
        j0 = REF.ref_select_branch(0 if rule == "most_fractional" else 1, np.ascontiguousarray(x), obj, n0, 1e-6)
        j1, frac = H.select_branch(x, obj, n0, 1e-6, rule)
        assert j0 == j1
        if j1 >= 0:
            assert frac == abs(x[j1] - np.floor(x[j1] + 0.5))


def test_oracle_heuristics_properties_without_the_reference():
    """Runs anywhere (the GPU box has no reference tree): the restatement's covers are feasible, honour the
    decisions, and cost what they say."""
    rng = np.random.default_rng(1)
    inst = scp_io.gen_scp(30, 150, 0.07, 2)
    x, y = lp_like_point(inst, rng)
    dec = [(3, 1), (10, 0)]
    f, o, s, steps = H.dual_guided_cover_repair(inst, x, y, dec)
    A0 = inst.csr()[:, :inst.n_orig]
    assert f and np.all(A0 @ s >= 1.0) and s[3] == 1 and s[10] == 0 and o == inst.c[:inst.n_orig] @ s and steps > 0
    f, o, s = H.nearest_integer_fixing(inst, np.ones(inst.n), dec)
    assert f and s[10] == 0 and o == inst.c[:inst.n_orig].sum() - inst.c[10]
