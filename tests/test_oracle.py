"""CPU: the oracle against the reference's golden vectors (tests/golden/*.npz were produced by the
reference's own python/interior_point.py, see tests/golden/make_golden.py) and known-answer tables."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import mehrotra as mo
from oracle import scp_io

SMALL = ["demo00", "scp_demo06", "scp_demo_tiny03", "scp41", "scp42", "scp46", "scp48", "scp49", "scp410",
         "scp51", "scpclr10", "scpcyc06", "scpa1", "scpb1"]
LARGE = ["scpnre1", "scpnrg1"]          # the other large fixtures are exercised on the GPU side
# iteration counts of the reference at the C++ parameters
EXPECTED_ITERS = {"scp41": 13, "scp42": 13, "scp46": 13, "scp48": 14, "scp49": 12, "scp410": 12, "scp51": 17,
                  "scpa1": 21, "scpclr10": 5, "scpcyc06": 2, "scpnre1": 39, "scpnrg1": 30, "scpnrh1": 34,
                  "scpnrf1": 42, "scpclr13": 3}


@pytest.mark.parametrize("name", SMALL + LARGE)
def test_oracle_matches_reference_iterates(name):
    inst, z = load_golden(name)
    r = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
    assert r.reason == mo.TERM_CONVERGED
    assert r.iterations == int(z["ref_iters"])
    if name in EXPECTED_ITERS:
        assert r.iterations == EXPECTED_ITERS[name]
    # objectives of the reference's own run, 1e-6 relative is the contract; we are far inside it
    assert abs(r.primal - float(z["ref_primal"])) <= 1e-9 * max(1.0, abs(float(z["ref_primal"])))
    assert abs(r.dual - float(z["ref_dual"])) <= 1e-9 * max(1.0, abs(float(z["ref_dual"])))
    # iterates
    for a, b in ((r.x, z["ref_x"]), (r.y, z["ref_y"]), (r.s, z["ref_s"])):
        assert np.max(np.abs(a - b) / (1.0 + np.abs(b))) < 1e-7


@pytest.mark.parametrize("name", ["demo00", "scp_demo06", "scp_demo_tiny03", "scpcyc06"])
def test_kkt_form_equals_normal_equations_form(name):
    """The CUDA reference solves the full KKT system by dense LU (sypha_solver_dense_linear.cpp);
    the normal-equations form must give the same trajectory."""
    inst, _ = load_golden(name)
    a = mo.solve_instance(inst, mo.Params(max_iter=100), "kkt")
    b = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
    assert a.iterations == b.iterations
    assert abs(a.primal - b.primal) < 1e-10 * max(1, abs(b.primal))
    assert abs(a.dual - b.dual) < 1e-10 * max(1, abs(b.dual))
    for ta, tb in zip(a.trace, b.trace):
        assert abs(ta["mu"] - tb["mu"]) <= 1e-8 * tb["mu"]


def test_known_lp_optima_band():
    """Exit objectives bracket the known LP optimum (python/sypha_unit_tests.py:21-77 and the GLOP
    column of benchmark_results_with_ip.csv); mu-only stop => ~1e-4 relative."""
    gold = json.load(open(GOLDEN / "lp_optima.json"))
    for name in ["scp41", "scp48", "scp49", "scp51", "scpa1", "scpclr10", "scpcyc06", "demo00"]:
        inst, z = load_golden(name)
        r = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
        g = gold[name]
        assert abs(float(z["lp_gold"]) - g) < 1e-12
        assert abs(r.dual - g) <= 2e-3 * max(1.0, abs(g))
        assert abs(r.primal - g) <= 2e-3 * max(1.0, abs(g))


def test_default_cap_binds_on_nre():
    """maxIter 25 is not enough on scpnre1."""
    inst, _ = load_golden("scpnre1")
    r = mo.solve_instance(inst, mo.Params(), "ne")
    assert r.iterations == 25 and r.reason == mo.TERM_MAX_ITER


def test_gap_stagnation_exit():
    inst, _ = load_golden("scp41")
    P = mo.Params(max_iter=100, gap_stagnation=True, gap_window=2, gap_min_improv_pct=60.0)
    r = mo.solve_instance(inst, P, "ne")
    assert r.reason == mo.TERM_GAP_STALLED and 0 < r.iterations < 13


def test_generator_check_values():
    """SURVEY.md Appendix C check values for gen_scp(1000, 20000, 0.005, 0)."""
    inst = scp_io.gen_scp(1000, 20000, 0.005, 0)
    assert inst.nnz == 101147
    r = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
    assert r.iterations == 24
    assert abs(r.primal - 554.166452726) < 1e-6 and abs(r.dual - 552.956531973) < 1e-6


def test_pcg_reference_schedule_fails_late():
    """Jacobi-PCG with the reference's cap (500, sypha_environment_defaults.h:21) hits the cap before
    mu <= 1e-4 on the synthetic instances -> the reference reports a failed LP.  With a
    large cap and a tight fixed tolerance the PCG trajectory matches the direct solve (the parity
    contract of config 4: iterations +-1, objectives 1e-6)."""
    inst = scp_io.gen_scp(300, 6000, 0.005, 1)
    d = mo.solve_instance(inst, mo.Params(max_iter=100), "ne")
    ref = mo.solve_instance(inst, mo.Params(max_iter=100), "pcg")
    assert ref.reason == mo.TERM_NUMERICAL and ref.status == 1
    p = mo.solve_instance(inst, mo.Params(max_iter=100, cg_max_iter=50000, cg_tol_initial=1e-8,
                                          cg_tol_final=1e-8, cg_tol_decay=1.0), "pcg")
    assert p.reason == mo.TERM_CONVERGED
    assert abs(p.iterations - d.iterations) <= 1
    assert abs(p.primal - d.primal) <= 1e-6 * abs(d.primal)
    assert abs(p.dual - d.dual) <= 1e-6 * abs(d.dual)


def test_branch_rows_follow_reference():
    """append_branch_rows == build_branch_model (sypha_solver_bnb.cpp:453-468)."""
    inst, _ = load_golden("scp_demo06")
    node = scp_io.append_branch_rows(inst, [(3, 1), (7, 0)])
    A = node.dense()
    assert node.m == inst.m + 2 and node.n == inst.n + 2
    assert A[inst.m, 3] == 1.0 and A[inst.m, inst.n] == -1.0 and node.b[inst.m] == 1.0
    assert A[inst.m + 1, 7] == -1.0 and A[inst.m + 1, inst.n + 1] == -1.0 and node.b[inst.m + 1] == 0.0
    r = mo.solve_instance(node, mo.Params(max_iter=100), "ne")
    assert r.reason == mo.TERM_CONVERGED and r.dual >= 4.33
