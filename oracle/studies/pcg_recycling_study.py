"""TEST INFRASTRUCTURE / study (CPU, NumPy): see DESIGN.md section 8, item 1.  Not imported by the product, the tests or the bench."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np, scipy.linalg as sl
from oracle import scp_io, mehrotra as mo

def pcg(matvec, diag, rhs, tol, cap, x0=None, store=False):
    x = np.zeros_like(rhs) if x0 is None else x0.copy()
    r = rhs.copy() if x0 is None else rhs - matvec(x0)
    z = r / diag; p = z.copy(); rz = r @ z; n0 = np.linalg.norm(rhs)
    P, PAP = [], []
    if np.linalg.norm(r) / n0 < tol: return x, 0, P, PAP
    for it in range(cap):
        Ap = matvec(p); pap = p @ Ap
        if store: P.append(p.copy()); PAP.append(pap)
        a = rz / pap; x += a * p; r -= a * Ap
        if np.linalg.norm(r) / n0 < tol: return x, it + 1, P, PAP
        z = r / diag; rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return x, cap, P, PAP

# python oracle/studies/pcg_recycling_study.py [m n density]   (default: the 1000 x 20000 rung of the ladder)
m, n, dens = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (1000, 20000, 0.005)
inst = scp_io.gen_scp(m, n, dens, 0)
A = inst.csr()
x, y, s = mo.start_point(A, inst.b, inst.c)
resC = inst.c - s - A.T @ y; resB = inst.b - A @ x
it = 0; mu = x @ s / inst.n
tot_cold = tot_rec = 0
while mu > 1e-4 and it < 100:
    d = x / s
    Md = ((A.multiply(d)) @ A.T).toarray(); diag = Md.diagonal().copy()
    cf = sl.cho_factor(Md)
    mv = lambda p: A @ (d * (A.T @ p))
    resXS = -x * s
    def rhs_of(rxs): return resB + A @ ((x * resC - rxs) / s)
    def rest(dy, rxs):
        ds = resC - A.T @ dy; dx = (rxs - x * ds) / s; return dx, ds
    b1 = rhs_of(resXS)
    dya, n1, P, PAP = pcg(mv, diag, b1, 1e-8, 20000, store=True)
    dxa, dsa = rest(dya, resXS)
    apa = min(1, mo.ratio_test(x, dxa)); ada = min(1, mo.ratio_test(s, dsa))
    mua = (x + apa * dxa) @ (s + ada * dsa) / inst.n; sig = (mua / mu) ** 3
    rxs2 = resXS - dxa * dsa + sig * mu
    b2 = rhs_of(rxs2)
    _, n2c, _, _ = pcg(mv, diag, b2, 1e-8, 20000)
    Pm = np.array(P).T if P else np.zeros((m, 0))
    coef = (Pm.T @ b2) / np.array(PAP) if P else np.zeros(0)
    x0 = Pm @ coef
    r0 = np.linalg.norm(b2 - mv(x0)) / np.linalg.norm(b2)
    dy, n2r, _, _ = pcg(mv, diag, b2, 1e-8, 20000, x0=x0)
    tot_cold += n1 + n2c; tot_rec += n1 + n2r
    print(f"it {it:2d} mu {mu:.2e}: affine {n1}, corrector cold {n2c}, recycled {n2r} (residual after projection {r0:.1e})", flush=True)
    dx, ds = rest(dy, rxs2)
    ap = min(1, 0.95 * mo.ratio_test(x, dx)); ad = min(1, 0.95 * mo.ratio_test(s, ds))
    x += ap * dx; y += ad * dy; s += ad * ds; resC *= (1 - ad); resB *= (1 - ap)
    mu = x @ s / inst.n; it += 1
print("iterations", it, "total CG cold", tot_cold, "with recycling", tot_rec)
