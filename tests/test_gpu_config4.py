"""Config 4 (BASELINE.json configs[3]: synthetic SCP up to 50k x 1M, PCG normal-equations path) at its own sizes.

The reference's Krylov branch cannot pin these (its right-hand side has the wrong sign, SURVEY.md F3, and its
Jacobi-PCG runs out of iterations, F12), so the parity target is the DIRECT solve on the ladder rungs where both
strategies run (iterations +-1, objectives 1e-6 - the north_star contract), SURVEY.md Appendix C's check values
for the generator + direct path, and - where no direct solve or oracle fits (50k x 1M) - size-independent
properties of the returned point: positivity, complementarity below the tolerance, the objectives recomputed
from the iterates, and the KKT residuals implied by the reference's update rule."""
import numpy as np
import pytest

from conftest import node_from_instance

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import sypha_b200 as sb  # noqa: E402
from oracle import scp_io  # noqa: E402

REL = 1e-6
PCG = dict(krylovMaxCgIter=200000, krylovCgTolInitial=1e-8, krylovCgTolFinal=1e-8, krylovCgTolDecayRate=1.0)


@pytest.fixture(scope="module")
def cuda_ws():
    """A workspace of this module's own: the ladder models grow it to 10k x 200k, and a grown workspace selects other
    kernel variants for small models (e.g. the assembly without shared-memory staging once the pad id no longer fits) -
    same results to rounding, but the session-wide workspace is also used by bit-equality tests."""
    w = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(w)
    yield w
    sb.releaseIpmWorkspace(w)


def solve(inst, ws, strategy, **kw):
    node = node_from_instance(inst, linearSolverStrategy=strategy, **kw)
    res = sb.SolverExecutionResult()
    st = sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=100), res, ws)
    assert st == sb.CODE_SUCCESSFUL and res.terminationReason == sb.SOLVER_TERM_CONVERGED, (st, res.terminationReason)
    return res


def close(a, b):
    return abs(a - b) <= REL * max(1.0, abs(b))


def test_appendix_c_check_values_1000x20000(cuda_ws):
    """gen_scp(1000, 20000, 0.005, 0): nnz 101147, 24 iterations, 554.166452726 / 552.956531973."""
    inst = scp_io.gen_scp(1000, 20000, 0.005, 0)
    assert inst.nnz == 101147            # nnz(A) of SURVEY App. C counts the surplus column of each row
    d = solve(inst, cuda_ws, "cholesky")
    assert d.iterations == 24 and close(d.primalObj, 554.166452726) and close(d.dualObj, 552.956531973)
    p = solve(inst, cuda_ws, "pcg", **PCG)
    assert abs(p.iterations - 24) <= 1 and close(p.primalObj, d.primalObj) and close(p.dualObj, d.dualObj)
    assert p.cgIterations > 0


def test_pcg_tracks_direct_on_the_2kx40k_rung(cuda_ws):
    inst = scp_io.gen_scp(2000, 40000, 0.001, 0)
    d = solve(inst, cuda_ws, "cholesky")
    p = solve(inst, cuda_ws, "pcg", **PCG)
    assert abs(p.iterations - d.iterations) <= 1
    assert close(p.primalObj, d.primalObj) and close(p.dualObj, d.dualObj)
    assert np.max(np.abs(p.primalSolution - d.primalSolution) / (1 + np.abs(d.primalSolution))) < 1e-4


def test_appendix_c_check_values_5kx100k_direct_and_pcg(cuda_ws):
    """gen_scp(5000, 100000, 0.001, 0): nnz 505691, 27 iterations, 2719.944694956 / 2710.607473538
    from the direct path, and the PCG path within the contract of it."""
    inst = scp_io.gen_scp(5000, 100000, 0.001, 0)
    assert inst.nnz == 505691
    d = solve(inst, cuda_ws, "cholesky")
    assert d.iterations == 27 and close(d.primalObj, 2719.944694956) and close(d.dualObj, 2710.607473538)
    p = solve(inst, cuda_ws, "pcg", **PCG)
    assert abs(p.iterations - 27) <= 1 and close(p.primalObj, d.primalObj) and close(p.dualObj, d.dualObj)


def check_point(inst, res, rb_tol, rc_tol):
    """Properties of an exit point that hold at any size (Appendix A of SURVEY.md)."""
    A = inst.csr()
    x, y = res.primalSolution, res.dualSolution
    s = res.slackSolution
    n0 = inst.n_orig
    assert x.min() > 0 and s.min() > 0
    mu = float(x @ s) / inst.n
    assert mu <= 1e-4 and abs(mu - res.mu) <= 1e-9 * max(1.0, mu) + 1e-12
    assert close(float(x[:n0] @ inst.c[:n0]), res.primalObj) and close(float(y @ inst.b), res.dualObj)
    rb = np.abs(inst.b - A @ x).max()
    rc = np.abs(inst.c - s - A.T @ y).max()
    # the reference never recomputes its residuals, it scales them by (1 - alpha) (sypha_solver.cpp:714-720):
    # exact solves would leave exactly prod(1 - alpha) of the initial residual; what is measured here also holds
    # the accumulated error of the inexact (CG) solves
    assert rb <= rb_tol and rc <= rc_tol, (rb, rc)
    # weak duality up to the residuals: c'x - b'y = x's + rc'x - rb'y
    gap = float(inst.c @ x - inst.b @ y)
    assert abs(gap - (float(x @ s) + float((inst.c - s - A.T @ y) @ x) - float((inst.b - A @ x) @ y))) <= 1e-6 * max(1.0, abs(gap))
    return mu, rb, rc


def test_exit_point_properties_10kx200k_pcg(cuda_ws):
    inst = scp_io.gen_scp_fast(10000, 200000, 0.001, 3)
    node = node_from_instance(inst, linearSolverStrategy="pcg", **PCG)
    res = sb.SolverExecutionResult()
    st = sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=100), res, cuda_ws)
    assert st == sb.CODE_SUCCESSFUL and res.terminationReason == sb.SOLVER_TERM_CONVERGED
    d = solve(inst, cuda_ws, "cholesky")
    assert abs(res.iterations - d.iterations) <= 1 and close(res.primalObj, d.primalObj) and close(res.dualObj, d.dualObj)
    check_point(inst, res, 1e-3, 1e-3)


def test_exit_point_properties_50kx1M_pcg():
    """configs[3] at full size: no oracle and no direct solve fits, so the returned point itself is checked."""
    inst = scp_io.gen_scp_fast(50000, 1000000, 0.001, 1)
    ws = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(ws)
    try:
        node = node_from_instance(inst, linearSolverStrategy="pcg", **PCG)
        res = sb.SolverExecutionResult()
        st = sb.solver_sparse_mehrotra_run(node, sb.SolverExecutionConfig(maxIterations=100), res, ws)
        assert st == sb.CODE_SUCCESSFUL and res.terminationReason == sb.SOLVER_TERM_CONVERGED
        assert 40 <= res.iterations <= 60
        mu, rb, rc = check_point(inst, res, 1e-3, 1e-3)
        print(f"50k x 1M: {res.iterations} iterations, {res.cgIterations} CG iterations, primal {res.primalObj:.9f} "
              f"dual {res.dualObj:.9f} mu {mu:.3e} |rb| {rb:.3e} |rc| {rc:.3e}")
    finally:
        sb.releaseIpmWorkspace(ws)
