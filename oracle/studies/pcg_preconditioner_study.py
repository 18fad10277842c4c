"""TEST INFRASTRUCTURE / study (CPU, NumPy): see DESIGN.md.  Not imported by the product, the tests or the bench."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, scipy.sparse as sp, scipy.linalg as sl
from oracle import scp_io, mehrotra as mo

def pcg(matvec, prec, rhs, tol, cap, x0=None):
    x = np.zeros_like(rhs) if x0 is None else x0.copy(); r = rhs.copy() if x0 is None else rhs - matvec(x0); z = prec(r); p = z.copy(); rz = r @ z; n0 = np.linalg.norm(rhs)
    for it in range(cap):
        Ap = matvec(p); a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) / n0 < tol: return it + 1
        z = prec(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return cap

m, n = int(sys.argv[1]), int(sys.argv[2])
inst = scp_io.gen_scp(m, n, float(sys.argv[3]), 0)
A = inst.csr(); At = A.T.tocsr()
# run direct IPM, capture d at chosen iterations
caps = {}
orig = mo.form_normal_matrix
res = mo.solve_instance(inst, mo.Params(max_iter=100), "ne", keep_d=True) if False else None
# manual loop: reuse oracle pieces
x, y, s = mo.start_point(A, inst.b, inst.c)
resC = inst.c - s - A.T @ y; resB = inst.b - A @ x
it = 0
mu = x @ s / inst.n
rows = []
while mu > 1e-4 and it < 100:
    d = x / s
    M = (A.multiply(d)) @ A.T
    Md = M.toarray()
    resXS = -x * s
    cf = sl.cho_factor(Md)
    def solve(rxs):
        t = (x * resC - rxs) / s
        rhs = resB + A @ t
        dy = sl.cho_solve(cf, rhs)
        ds = resC - A.T @ dy
        dx = (rxs - x * ds) / s
        return dx, dy, ds, rhs
    dxa, dya, dsa, rhs = solve(resXS)
    if it in (7, 13, 17, 20, 22, 23):
        diag = Md.diagonal().copy()
        mv = lambda p: A @ (d * (A.T @ p))
        r0 = {}
        r0['jacobi'] = pcg(mv, lambda r: r / diag, rhs, 1e-8, 5000)
        for k in ():
            if k >= n // 4: continue
            K = np.argsort(-d)[:k]
            U = A[:, K].multiply(np.sqrt(d[K])).tocsc()          # m x k
            dl = diag - np.asarray(U.multiply(U).sum(axis=1)).ravel()
            dl = np.maximum(dl, 1e-300)
            G = np.eye(k) + (U.T @ sp.diags(1 / dl) @ U).toarray()
            cg = sl.cho_factor(G)
            def prec(r, U=U, dl=dl, cg=cg):
                t = r / dl
                return t - (U @ sl.cho_solve(cg, U.T @ t)) / dl
            r0[f'lowrank{k}'] = pcg(mv, prec, rhs, 1e-8, 5000)
        for bs in ():
            blocks = [sl.cho_factor(Md[i:i + bs, i:i + bs]) for i in range(0, m, bs)]
            def precb(r, blocks=blocks, bs=bs):
                z = np.empty_like(r)
                for q, i in enumerate(range(0, m, bs)): z[i:i + bs] = sl.cho_solve(blocks[q], r[i:i + bs])
                return z
            r0[f'blockjac{bs}'] = pcg(mv, precb, rhs, 1e-8, 5000)
        print(f"it {it:2d} mu {mu:.2e} d range {d.min():.1e}..{d.max():.1e} cond {np.linalg.cond(Md):.1e} :", r0, flush=True)
    apa = min(1, mo.ratio_test(x, dxa)); ada = min(1, mo.ratio_test(s, dsa))
    mua = (x + apa * dxa) @ (s + ada * dsa) / inst.n
    sig = (mua / mu) ** 3
    rxs2 = resXS - dxa * dsa + sig * mu
    dx, dy, ds, rhs2 = solve(rxs2)
    if it in (7, 13, 17, 20, 22, 23):
        diag = Md.diagonal().copy(); mv = lambda p: A @ (d * (A.T @ p))
        print('   corrector CG: cold', pcg(mv, lambda r: r / diag, rhs2, 1e-8, 5000), 'warm from dy_aff', pcg(mv, lambda r: r / diag, rhs2, 1e-8, 5000, dya), ' |dy-dya|/|dy| %.2e' % (np.linalg.norm(dy-dya)/np.linalg.norm(dy)), flush=True)
    ap = min(1, 0.95 * mo.ratio_test(x, dx)); ad = min(1, 0.95 * mo.ratio_test(s, ds))
    x += ap * dx; y += ad * dy; s += ad * ds
    resC *= (1 - ad); resB *= (1 - ap)
    mu = x @ s / inst.n; it += 1
print("iterations", it)
