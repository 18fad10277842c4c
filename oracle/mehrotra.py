"""NumPy restatement of the reference's Mehrotra predictor-corrector LP solve.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Never imported by sypha_b200/.

What is restated, and from where (all paths under /root/reference):

* starting point ........ src/sypha_solver_init.cpp:543-652  (== python/interior_point.py:13-57)
* initial residuals ..... src/sypha_solver.cpp:375-459
* main loop ............. src/sypha_solver.cpp:496-772
* exit / status ......... src/sypha_solver.cpp:774-797
* ratio-test predicate .. src/sypha_solver_utils.cu:68-79  (strict ``< 0``, empty set -> DBL_MAX)
* parameters ............ src/sypha_environment_defaults.h:14-24
* Krylov schedule ....... src/sypha_solver.cpp:552-553, src/sypha_solver_krylov.cu:243-392
                          (the RHS sign of krylov.cu:177-219 is NOT followed; it contradicts the
                          KKT system of sypha_solver.cpp:84-92 - SURVEY.md F3)

Sign convention is the C++ one: resC = c - s - A'y, resB = b - Ax, resXS = -x.*s (the Python
prototype uses the negatives).

Three linear-solve back ends produce the same Newton directions:
  ``kkt``  dense LU of the full (2n+m) KKT matrix - literally what the CUDA reference executes
           (src/sypha_solver_dense_linear.cpp:150-203); O((2n+m)^3), small instances only.
  ``ne``   normal equations (A D A') dy = resB + A((x.*resC - resXS)./s), D = x./s, Cholesky;
           ds = resC - A'dy; dx = (resXS - x.*ds)./s   (python/interior_point.py:112-121).
  ``pcg``  the same normal equations solved by Jacobi-preconditioned CG with the reference's
           tolerance schedule.
"""
from __future__ import annotations

import dataclasses
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

DBL_MAX = np.finfo(np.float64).max

# src/sypha_solver_sparse.h:13-20
TERM_CONVERGED, TERM_MAX_ITER, TERM_GAP_STALLED, TERM_NUMERICAL, TERM_TIME_LIMIT = range(5)


@dataclasses.dataclass
class Params:
    """src/sypha_environment_defaults.h:14-24 and src/sypha_solver_sparse.h:22-36."""
    max_iter: int = 25
    eta: float = 0.95
    mu_tol: float = 1e-4
    gap_stagnation: bool = False
    gap_window: int = 0
    gap_min_improv_pct: float = 0.0
    cg_max_iter: int = 500
    cg_tol_initial: float = 1e-2
    cg_tol_final: float = 1e-8
    cg_tol_decay: float = 0.5


@dataclasses.dataclass
class Result:
    status: int
    reason: int
    iterations: int
    primal: float
    dual: float
    rel_gap: float
    x: np.ndarray
    y: np.ndarray
    s: np.ndarray
    x0: np.ndarray
    y0: np.ndarray
    s0: np.ndarray
    trace: list
    cg_iters: list
    loop_seconds: float = 0.0
    start_seconds: float = 0.0


def _chol_solve_factory(M):
    cf = sla.cho_factor(M, lower=True, check_finite=False)
    return lambda r: sla.cho_solve(cf, r, check_finite=False)


def form_normal_matrix(A: sp.csr_matrix, d: np.ndarray) -> np.ndarray:
    """Dense M = A diag(d) A'.  (python/interior_point.py:114-115)"""
    AD = A.multiply(d[None, :]).tocsr()
    M = (AD @ A.T)
    return np.asarray(M.todense()) if sp.issparse(M) else np.asarray(M)


def start_point(A: sp.csr_matrix, b, c, explicit_inverse=False):
    """Mehrotra's starting-point heuristic, src/sypha_solver_init.cpp:543-652.

    x~ = A'(AA')^-1 b ; y~ = (AA')^-1 A c ; s~ = c - A'y~ ;
    dx = max(-1.5 min x~, 0), ds likewise ; x^ = x~+dx, s^ = s~+ds ;
    x0 = x^ + .5 (x^.s^)/sum(s^) ; s0 = s^ + .5 (x^.s^)/sum(x^).
    ``explicit_inverse`` follows the reference literally (LU inverse of AA', init.cpp:586-591).
    """
    AAT = form_normal_matrix(A, np.ones(A.shape[1]))
    if explicit_inverse:
        inv = np.linalg.inv(AAT)
        x = A.T @ (inv @ b)
        y = inv @ (A @ c)
    else:
        solve = _chol_solve_factory(AAT)
        x = A.T @ solve(b)
        y = solve(A @ c)
    s = c - A.T @ y
    dx = max(-1.5 * x.min(), 0.0)
    ds = max(-1.5 * s.min(), 0.0)
    x = x + dx
    s = s + ds
    prod = 0.5 * float(x @ s)
    dxh = prod / s.sum()
    dsh = prod / x.sum()
    return x + dxh, y, s + dsh


def ratio_test(v, dv):
    """min over dv<0 of -v/dv, DBL_MAX if the set is empty (src/sypha_solver_utils.cu:68-79)."""
    neg = dv < 0.0
    if not neg.any():
        return DBL_MAX
    return float(np.min(-v[neg] / dv[neg]))


def _kkt_matrix(A_dense, x, s):
    """[0 A' I; A 0 0; S 0 X]  (src/sypha_solver.cpp:84-92,113-186)."""
    m, n = A_dense.shape
    N = 2 * n + m
    K = np.zeros((N, N))
    K[:n, n:n + m] = A_dense.T
    K[:n, n + m:] = np.eye(n)
    K[n:n + m, :n] = A_dense
    K[n + m:, :n] = np.diag(s)
    K[n + m:, n + m:] = np.diag(x)
    return K


def pcg_jacobi(matvec, rhs, diag, tol, max_iter):
    """Jacobi-PCG restating src/sypha_solver_krylov.cu:243-392.  Returns (dy, iters) with
    iters = -1 on failure (pAp <= 0, non-finite, |rz| < 1e-30, or the cap is reached)."""
    m = rhs.shape[0]
    rhs_norm = float(np.linalg.norm(rhs))
    dy = np.zeros(m)
    if rhs_norm < 1e-30:
        return dy, 0
    pd = np.maximum(diag, 1e-30)
    r = rhs.copy()
    z = r / pd
    p = z.copy()
    rz = float(r @ z)
    for it in range(max_iter):
        Ap = matvec(p)
        pAp = float(p @ Ap)
        if pAp <= 0.0 or not np.isfinite(pAp):
            return dy, -1
        a = rz / pAp
        dy += a * p
        r -= a * Ap
        if float(np.linalg.norm(r)) / rhs_norm < tol:
            return dy, it + 1
        z = r / pd
        rz_new = float(r @ z)
        if abs(rz) < 1e-30:
            return dy, -1
        beta = rz_new / rz
        rz = rz_new
        p = z + beta * p
    return dy, -1


def mehrotra(A: sp.csr_matrix, b, c, n_orig, params: Params = None, solver="ne",
             start=None, stop_after=None) -> Result:
    """The reference's LP solve (see module docstring for the line-by-line map).

    ``start`` overrides the starting point (x, y, s); ``stop_after`` truncates the loop after that
    many iterations without touching the termination logic (used to line the oracle up with the
    Python prototype, whose own stop test is hard-coded to mu > 1e-10).
    """
    P = params or Params()
    A = sp.csr_matrix(A)
    AT = A.T.tocsr()
    m, n = A.shape
    b = np.asarray(b, dtype=np.float64)
    c = np.asarray(c, dtype=np.float64)

    t0 = time.perf_counter()
    if start is None:
        x, y, s = start_point(A, b, c, explicit_inverse=(solver == "kkt"))
    else:
        x, y, s = (np.array(v, dtype=np.float64) for v in start)
    x0, y0, s0 = x.copy(), y.copy(), s.copy()
    t1 = time.perf_counter()

    # src/sypha_solver.cpp:375-459
    resC = c - s - AT @ y
    resB = b - A @ x
    mu = float(x @ s) / n

    A_dense = A.toarray() if solver == "kkt" else None
    A_sq = A.multiply(A).tocsr() if solver == "pcg" else None

    it = 0
    best_gap = np.inf
    stall = 0
    reason = TERM_MAX_ITER
    numerical = False
    gap_enabled = P.gap_stagnation and P.gap_window > 0 and P.gap_min_improv_pct >= 0.0
    min_ratio = P.gap_min_improv_pct / 100.0
    trace, cg_iters = [], []
    limit = P.max_iter if stop_after is None else min(P.max_iter, stop_after)

    while it < limit and mu > P.mu_tol:
        resXS = -x * s                                           # :505
        d = x / s

        if solver == "kkt":
            lu = sla.lu_factor(_kkt_matrix(A_dense, x, s), check_finite=False)

            def solve(rxs):
                sol = sla.lu_solve(lu, np.concatenate([resC, resB, rxs]), check_finite=False)
                return sol[:n], sol[n:n + m], sol[n + m:]
        else:
            if solver == "ne":
                try:
                    msolve = _chol_solve_factory(form_normal_matrix(A, d))
                except (np.linalg.LinAlgError, ValueError):
                    numerical, reason = True, TERM_NUMERICAL
                    break
            else:
                diag = np.asarray(A_sq @ d).ravel()              # krylov.cu:26-43
                tol = max(P.cg_tol_final, P.cg_tol_initial * P.cg_tol_decay ** it)   # solver.cpp:552
                mv = lambda p: A @ (d * (AT @ p))

                def msolve(r):
                    dy, k = pcg_jacobi(mv, r, diag, tol, P.cg_max_iter)
                    cg_iters.append(k)
                    if k < 0:
                        raise FloatingPointError("cg failed")
                    return dy

            def solve(rxs):
                dy = msolve(resB + A @ ((x * resC - rxs) / s))
                ds = resC - AT @ dy
                dx = (rxs - x * ds) / s
                return dx, dy, ds

        try:
            dxa, dya, dsa = solve(resXS)                         # :514-593
        except FloatingPointError:
            numerical, reason = True, TERM_NUMERICAL
            break
        ap_aff = min(1.0, ratio_test(x, dxa))                    # :596-601
        ad_aff = min(1.0, ratio_test(s, dsa))
        mu_aff = float((x + ap_aff * dxa) @ (s + ad_aff * dsa)) / n   # :609-619
        sigma = (mu_aff / mu) ** 3                               # :622
        resXS = resXS + (-dxa * dsa + sigma * mu)                # :625-629
        try:
            dx, dy, ds = solve(resXS)                            # :633-689
        except FloatingPointError:
            numerical, reason = True, TERM_NUMERICAL
            break
        ap = min(1.0, P.eta * ratio_test(x, dx))                 # :693-698
        ad = min(1.0, P.eta * ratio_test(s, ds))
        x = x + ap * dx                                          # :703-710
        y = y + ad * dy
        s = s + ad * ds
        resC = resC * (-(ad - 1.0))                              # :714-720 (scaled, not recomputed)
        resB = resB * (-(ap - 1.0))
        mu_prev = mu
        mu = float(x @ s) / n                                    # :722
        if not np.isfinite(mu) or mu < 0.0:
            numerical, reason = True, TERM_NUMERICAL
            break
        primal = float(x[:n_orig] @ c[:n_orig])                  # :740-745
        dual = float(y @ b)
        gap = abs(primal - dual) / max(1.0, abs(primal))
        trace.append(dict(it=it, mu_in=mu_prev, mu=mu, mu_aff=mu_aff, sigma=sigma,
                          alpha_p=ap, alpha_d=ad, primal=primal, dual=dual, gap=gap))
        if not (np.isfinite(primal) and np.isfinite(dual) and np.isfinite(gap)):
            numerical, reason = True, TERM_NUMERICAL
            break
        if gap < best_gap * (1.0 - min_ratio):                   # :755-769
            best_gap = gap
            stall = 0
        elif gap_enabled:
            stall += 1
            if stall >= P.gap_window:
                reason = TERM_GAP_STALLED
                it += 1
                break
        it += 1
    t2 = time.perf_counter()

    if not numerical and reason != TERM_GAP_STALLED:             # :775-778
        reason = TERM_CONVERGED if mu <= P.mu_tol else TERM_MAX_ITER
    primal = float(x[:n_orig] @ c[:n_orig])                      # :781-785
    dual = float(y @ b)
    rel_gap = abs(primal - dual) / max(1.0, abs(primal))
    viol = dual - primal
    if (not numerical and reason != TERM_CONVERGED and np.isfinite(viol)
            and viol > 1e6 * max(1.0, abs(primal))):             # :788-796
        numerical, reason = True, TERM_NUMERICAL
    return Result(status=1 if numerical else 0, reason=reason, iterations=it, primal=primal,
                  dual=dual, rel_gap=rel_gap, x=x, y=y, s=s, x0=x0, y0=y0, s0=s0, trace=trace,
                  cg_iters=cg_iters, loop_seconds=t2 - t1, start_seconds=t1 - t0)


def solve_instance(inst, params: Params = None, solver="ne", **kw) -> Result:
    return mehrotra(inst.csr(), inst.b, inst.c, inst.n_orig, params, solver, **kw)
